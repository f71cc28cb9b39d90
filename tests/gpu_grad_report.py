"""Standalone gradient parity report (GPU box): lgar_columns autograd vs the reference-autograd goldens."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import golden_names, load_golden
from gpu_common import ensemble_from_golden
import lgar_b200
from lgar_b200 import lgar_columns

names = sys.argv[1:] or golden_names(prefix="grad_")
for name in names:
    g = load_golden(name)
    ens, a, n, k = ensemble_from_golden(g, copies=2)
    line = [f"{name:26s}"]
    for loss_name in ("AET", "infiltration", "runoff", "final_volume"):
        A = torch.tensor(a, device="cuda", requires_grad=True)
        N = torch.tensor(n, device="cuda", requires_grad=True)
        K = torch.tensor(k, device="cuda", requires_grad=True)
        out = lgar_columns(A, N, K, ens, outputs=("runoff", "AET", "infiltration", "ending_volume"))
        if loss_name == "final_volume":
            loss = out["ending_volume"][-1, 0]
        else:
            loss = out[loss_name][:, 0].sum()
        t0 = time.time()
        loss.backward()
        torch.cuda.synchronize()
        dt = time.time() - t0
        mine = np.stack([A.grad[:, 0].cpu().numpy(), N.grad[:, 0].cpu().numpy(), K.grad[:, 0].cpu().numpy()])
        ref = g[f"grad_{loss_name}"]
        scale = max(np.abs(ref).max(), 1e-300)
        err = np.abs(mine - ref) / (1e-9 * np.abs(ref) + 1e-11 * scale)
        line.append(f"{loss_name}: excess {err.max():.3g} (|g|max {scale:.3g}, {dt*1e3:.0f} ms)")
        if not np.isfinite(err.max()) or err.max() > 1:
            print("   ref ", ref.ravel()); print("   mine", mine.ravel())
    print("  ".join(line), flush=True)
