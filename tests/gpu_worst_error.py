"""Diagnostic: where is the largest CUDA-vs-oracle deviation of the full-year random-column parity test?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw, OUT_NAMES
from oracle import lgar_oracle as O
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, T = 64, 8760
big = workloads.synthetic_sites_ensemble(B=125_000, T=T, sites=128, rank=0) if which == "c4" else workloads.bushland_ensemble(B=100_000, T=T, seed=0)
cols = np.random.default_rng(5).choice(big.num_columns, B, replace=False)
sl = lambda x: np.ascontiguousarray(x[:, cols])
ens = ColumnEnsemble(theta_r=sl(big.theta_r), theta_e=sl(big.theta_e), thickness=sl(big.thickness), forcing=big.forcing, site_index=big.site_index[cols])
res, _ = forward_raw(ens, sl(big.alpha), sl(big.n), sl(big.ksat), outputs=OUT_NAMES, num_fronts=True)
torch.cuda.synchronize()
series = res.per_step.cpu().numpy()
rows = []
for j, b in enumerate(cols):
    cfg = O.make_cfg(big.alpha[:, b], big.n[:, b], big.ksat[:, b], big.theta_r[:, b], big.theta_e[:, b], thickness=big.thickness[:, b], iter_cap=1_000_000)
    r = O.forward(cfg, big.forcing[big.site_index[b]], fronts=False)
    n = T if r["status"] == 0 else r["crash_step"]
    for k in range(10):
        a, o = series[k, :n, j], r["out"][:n, k]
        if n == 0: continue
        ex = np.abs(a - o) / (1e-12 + 1e-9 * np.abs(o))
        t = int(np.argmax(ex))
        rows.append((float(ex[t]), int(b), OUT_NAMES[k], t, float(a[t]), float(o[t]), float(np.abs(a - o).max())))
rows.sort(reverse=True)
for r in rows[:12]:
    print("excess %.3f col %d %s step %d gpu %.17g oracle %.17g (max abs diff of the series %.3g)" % r)
