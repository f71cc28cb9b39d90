"""Diagnostic: what does ONE tile of the balanced C4 shard cost over the year?  (the tail of the forward pass)
  python tests/gpu_diag_tile.py RANK TILE
1. the tile's 32 columns as one warp, the year in 20 windows, each timed;
2. 32 tiles of 32 copies of one column each: the sequential cost of every column on its own (tile busy time)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw

rank, tile = int(sys.argv[1]), int(sys.argv[2])
B, T = 125_000, 8760
we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=128, rank=rank, shared_sites=True)
k0 = torch.as_tensor(we.ksat[0])
order = torch.argsort(k0, stable=True)
order = order[torch.argsort(torch.as_tensor(we.site_index)[order].to(torch.int64), stable=True)].numpy()
cols = order[tile * 32:(tile + 1) * 32]
print("tile", tile, "site(s)", sorted(set(we.site_index[cols].tolist())), "ksat0 range", we.ksat[0, cols].min(), we.ksat[0, cols].max())


def ens_of(idx):
    sub = lambda x: np.ascontiguousarray(x[:, idx])
    e = ColumnEnsemble(theta_r=sub(we.theta_r), theta_e=sub(we.theta_e), thickness=sub(we.thickness), forcing=we.forcing,
                       site_index=we.site_index[idx])
    return e, sub(we.alpha), sub(we.n), sub(we.ksat)


# 1. the tile as it runs in the bench, window by window
ens, a, n, k = ens_of(cols)
ws = res = None
nseg = 20
seg = -(-T // nseg)
times = []
for i in range(nseg):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res, ws = forward_raw(ens, a, n, k, outputs=(), per_step=False, workspace=ws, window=(i * seg, min(T, (i + 1) * seg)), into=res)
    e1.record(); torch.cuda.synchronize()
    times.append(round(e0.elapsed_time(e1), 1))
st = res.status.cpu().numpy(); cr = res.crash_step.cpu().numpy()
print("window ms", times, "total s", round(sum(times) / 1e3, 2))
print("status", st.tolist())
print("crash ", cr.tolist())

# 2. every column on its own (32 identical lanes per tile)
idx = np.repeat(cols, 32)
ens2, a2, n2, k2 = ens_of(idx)
r2, _ = forward_raw(ens2, a2, n2, k2, outputs=(), per_step=False, counters=True, tile_cycles=True)
torch.cuda.synchronize()
tc = r2.tile_cycles.cpu().numpy()
busy = (tc[0] if tc.ndim == 2 else tc) / 1.965e9
print("per-column sequential time s", [round(float(x), 2) for x in busy])
j = int(np.argmax(busy))
print("slowest column", int(cols[j]), "status", int(st[j]), "crash", int(cr[j]), "alpha", we.alpha[:, cols[j]].tolist(), "n", we.n[:, cols[j]].tolist(),
      "ksat", we.ksat[:, cols[j]].tolist())
# where in the year is the slowest column slow?  (one tile of 32 copies, 20 windows)
e3, a3, n3, k3 = ens_of(np.repeat(cols[j:j + 1], 32))
ws = res = None
t3 = []
for i in range(nseg):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    res, ws = forward_raw(e3, a3, n3, k3, outputs=(), per_step=False, workspace=ws, window=(i * seg, min(T, (i + 1) * seg)), into=res, counters=True)
    e1.record(); torch.cuda.synchronize()
    t3.append(round(e0.elapsed_time(e1), 1))
    if i == nseg - 1 or t3[-1] > 300:
        print("  window", i, "ms", t3[-1], "counters", res.counters.cpu().numpy()[:8].tolist())
print("slowest column alone, window ms", t3)
