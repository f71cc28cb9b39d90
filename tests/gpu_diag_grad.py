"""Time-boxed timing of forward(keep checkpoints) + backward for a (columns x steps) shape."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, lgar_columns
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(2048, 256)]
for B, T in shapes:
    we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=max(1, min(128, B // 32)), rank=0)
    ens = ColumnEnsemble(theta_r=we.theta_r, theta_e=we.theta_e, thickness=we.thickness, forcing=we.forcing,
                         site_index=we.site_index, chunk_steps=int(os.environ.get("LGAR_DIAG_CHUNK", "64")), reverse_counters=True)
    if os.environ.get("LGAR_DIAG_BALANCE"):
        ens.balance(we.ksat)
    for rep in range(2):
        A = torch.tensor(we.alpha, device="cuda", requires_grad=True)
        N = torch.tensor(we.n, device="cuda", requires_grad=True)
        K = torch.tensor(we.ksat, device="cuda", requires_grad=True)
        torch.cuda.synchronize(); e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = lgar_columns(A, N, K, ens, outputs=("runoff", "AET"))
        e[1].record()
        ok = out["status"] == 0
        loss = torch.nan_to_num(out["runoff"] + out["AET"]).sum(dim=0)[ok].mean()
        loss.backward()
        e[2].record(); torch.cuda.synchronize()
    st = out["status"].cpu().numpy(); cr = out["crash_step"].cpu().numpy()
    alive = int(np.where(st == 0, T, np.maximum(cr, 0)).sum())
    f, b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    gn = A.grad[:, ok]
    rc = ens.last_reverse_counters.cpu().numpy().astype(np.float64)
    print(f"  reverse kernel: taped recompute {100 * rc[0] / max(rc[0] + rc[1], 1):.1f} % of warp cycles, reverse sweeps "
          f"{100 * rc[1] / max(rc[0] + rc[1], 1):.1f} %; {rc[2] / max(rc[3], 1):.1f} tape entries per sub-step; overflowed columns {int(rc[4])}; of the recompute: move sweep {100 * rc[5] / max(rc[0], 1):.1f} %, calc_dzdt Geff {100 * rc[6] / max(rc[0], 1):.1f} %, other Geff {100 * rc[7] / max(rc[0], 1):.1f} %", flush=True)
    print(f"B={B} T={T}: fwd {f:.0f} ms, bwd {b:.0f} ms -> fwd+grad {alive/(f+b)*1e3:.4g} col-steps/s (fwd alone {alive/f*1e3:.4g}); "
          f"grad finite frac {float(torch.isfinite(gn).float().mean()):.4f}  |dL/dalpha0| mean {float(gn[0].abs().nanmean()):.3g}", flush=True)
