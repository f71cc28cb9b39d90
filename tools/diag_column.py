"""Diagnostic: one column of the bench ensemble on the GPU (front dumps) against the CPU oracle, step by step.
usage: python tools/diag_column.py COLUMN [T] [c4|c3]   -> first step where the front lists or the fluxes differ."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw, OUT_NAMES
from oracle import lgar_oracle as O
col = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 8760
which = sys.argv[3] if len(sys.argv) > 3 else "c4"
big = workloads.synthetic_sites_ensemble(B=125_000, T=8760, sites=128, rank=0) if which == "c4" else workloads.bushland_ensemble(B=100_000, T=8760, seed=0)
cols = np.array([col])
sl = lambda x: np.ascontiguousarray(x[:, cols])
forcing = big.forcing[:, :T]
ens = ColumnEnsemble(theta_r=sl(big.theta_r), theta_e=sl(big.theta_e), thickness=sl(big.thickness), forcing=forcing, site_index=big.site_index[cols])
print("params alpha", big.alpha[:, col], "n", big.n[:, col], "ksat", big.ksat[:, col], "site", big.site_index[col])
for dump in (True, False):
    res, _ = forward_raw(ens, sl(big.alpha), sl(big.n), sl(big.ksat), outputs=OUT_NAMES, num_fronts=True, dump_fronts=dump)
    torch.cuda.synchronize()
    print("dump" if dump else "production", "kernel: status", int(res.status[0]), "crash", int(res.crash_step[0]))
    if dump:
        keep = res
res = keep
cfg = O.make_cfg(big.alpha[:, col], big.n[:, col], big.ksat[:, col], big.theta_r[:, col], big.theta_e[:, col], thickness=big.thickness[:, col], iter_cap=1_000_000)
for name, ctx in (("glibc", None), ("devpow", O.device_pow())):
    if ctx: ctx.__enter__()
    r = O.forward(cfg, forcing[big.site_index[col]], fronts=True)
    if ctx: ctx.__exit__()
    print("oracle", name, "status", r["status"], "crash", r["crash_step"])
    nf = res.num_fronts.cpu().numpy()[:, 0]
    fr = res.fronts.cpu().numpy()[..., 0]   # [T,16,5]
    series = res.per_step.cpu().numpy()[:, :, 0]  # [10,T]
    n = min(T if r["status"] == 0 else r["crash_step"], T if int(res.status[0]) == 0 else int(res.crash_step[0]))
    first = None
    for t in range(n):
        if nf[t] != r["nfronts"][t] or not np.array_equal(fr[t].view(np.int64), r["fronts"][t].view(np.int64)) \
                or not np.array_equal(series[:, t].view(np.int64), r["out"][t].view(np.int64)):
            first = t
            break
    print("  first step with ANY bit difference:", first, "of", n)
    if first is not None:
        for t in range(max(0, first - 1), min(n, first + 3)):
            print("  step", t, "nfronts gpu/oracle", nf[t], r["nfronts"][t], "forcing", forcing[big.site_index[col], t])
            for i in range(max(nf[t], r["nfronts"][t])):
                print("     front", i, "gpu", [float(x) for x in fr[t, i]], "lay", int(res.front_layer[t, i, 0]))
                print("     front", i, "ora", [float(x) for x in r["fronts"][t, i]], "lay", int(r["front_layer"][t, i]))
            print("     out gpu", [float(x) for x in series[:, t]])
            print("     out ora", [float(x) for x in r["out"][t]])
    # first step where the discrete state differs
    fd = next((t for t in range(n) if nf[t] != r["nfronts"][t]), None)
    print("  first step with a different front COUNT:", fd)
