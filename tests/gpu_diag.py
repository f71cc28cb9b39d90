"""Time-boxed scaling diagnostic (run under `timeout`): elapsed time of one forward pass for a
series of (columns, steps) shapes of the bench ensemble."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(2048, 256)]
for B, T in shapes:
    we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=max(1, min(128, B // 32)), rank=0)
    ens = ColumnEnsemble(theta_r=we.theta_r, theta_e=we.theta_e, thickness=we.thickness, forcing=we.forcing,
                         site_index=we.site_index)
    torch.cuda.synchronize(); t0 = time.time()
    res, ws = forward_raw(ens, we.alpha, we.n, we.ksat, outputs=("runoff", "AET"), counters=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = res.status.cpu().numpy(); cr = res.crash_step.cpu().numpy()
    alive = int(np.where(st == 0, T, np.maximum(cr, 0)).sum())
    print(f"B={B} T={T}: {dt:.3f}s  alive col-steps={alive}  rate={alive/dt:.3g}/s  status hist={np.bincount(st, minlength=9).tolist()} counters={res.counters.cpu().numpy().tolist()}", flush=True)
