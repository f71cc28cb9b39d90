"""Time-boxed scaling diagnostic (run under `timeout`): elapsed time of one forward pass for a
series of (columns, steps) shapes of the bench ensemble."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(2048, 256)]
for B, T in shapes:
    we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=max(1, min(128, B // 32)), rank=int(os.environ.get("LGAR_DIAG_RANK", "0")),
                                             shared_sites=bool(os.environ.get("LGAR_DIAG_SHARED")))
    sl = os.environ.get("LGAR_DIAG_SLICE")
    if sl:
        a, cnt = (int(x) for x in sl.split(":"))
        for k in ("alpha", "n", "ksat", "theta_r", "theta_e", "thickness"):
            setattr(we, k, np.ascontiguousarray(getattr(we, k)[:, a:a + cnt]))
        we.site_index = np.ascontiguousarray(we.site_index[a:a + cnt])
        B = cnt
    rep = os.environ.get("LGAR_DIAG_REPLICATE")
    if rep:  # experiment: every warp works on a copy of the same 32 columns (tile `rep`): the warps of an SM then run in phase
        k = int(rep)
        idx = np.tile(np.arange(32 * k, 32 * k + 32), (B + 31) // 32)[:B]
        for key in ("alpha", "n", "ksat", "theta_r", "theta_e", "thickness"):
            setattr(we, key, np.ascontiguousarray(getattr(we, key)[:, idx]))
        we.site_index = np.ascontiguousarray(we.site_index[idx])
    ens = ColumnEnsemble(theta_r=we.theta_r, theta_e=we.theta_e, thickness=we.thickness, forcing=we.forcing,
                         site_index=we.site_index, max_fronts=int(os.environ.get("LGAR_DIAG_FM", "16")),
                         chunk_steps=int(os.environ.get("LGAR_DIAG_CHUNK", "64")))
    if os.environ.get("LGAR_DIAG_BALANCE"):
        ens.balance(we.ksat)
    torch.cuda.synchronize(); t0 = time.time()
    res, ws = forward_raw(ens, we.alpha, we.n, we.ksat, outputs=("runoff", "AET"), counters=True, tile_cycles=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    times = []
    for _ in range(int(os.environ.get("LGAR_DIAG_REPS", "3"))):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); forward_raw(ens, we.alpha, we.n, we.ksat, outputs=("runoff", "AET"), workspace=ws); e1.record()
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    kms = min(times)
    cn = res.counters.cpu().numpy().astype(np.float64)
    if cn.shape[0] >= 13 and cn[12] > 0:
        names = ("insert-water Geff", "move sweep + merge/cross/fix/update_psi", "dry-depth Geff + surficial front", "calc_dzdt Geff")
        print("  warp-cycle shares (counting pass): " + ", ".join(f"{nm} {100 * cn[8 + i] / cn[12]:.1f} %" for i, nm in enumerate(names))
              + f", rest {100 * (1 - cn[8:12].sum() / cn[12]):.1f} %", flush=True)
    st = res.status.cpu().numpy(); cr = res.crash_step.cpu().numpy()
    alive = int(np.where(st == 0, T, np.maximum(cr, 0)).sum())
    if os.environ.get("LGAR_DIAG_SAVE"):
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez_compressed(f"gpurun_out/diag_{B}x{T}{os.environ.get('LGAR_DIAG_TAG', '')}.npz", sums=res.sums.cpu().numpy(), status=st, crash=cr)
    tc = res.tile_cycles.cpu().numpy()
    if tc.ndim == 2:  # counting kernel: busy cycles, wait cycles, finish time (ns) per tile
        wait, fin = tc[1].astype(np.float64), tc[2].astype(np.float64)
        tc = tc[0]
        rel = (fin.max() - fin) * 1e-9
        wtop = np.argsort(-wait)[:4]
        print(f"  scheduler: warps blocked on a predecessor chunk {wait.sum() / 1.965e9:.1f} warp-s = {wait.sum() / 1.965e9 / 1184:.3f} s per resident warp; "
              f"most waited-for tiles {[(int(i), round(float(wait[i]) / 1.965e9, 2)) for i in wtop]}", flush=True)
        print(f"  tail: last tile finishes {np.percentile(rel, 50):.3f} s after the median tile, {np.percentile(rel, 10):.3f} s after the 90th percentile, "
              f"{np.sort(rel)[min(len(rel) - 1, 1184)]:.3f} s after all but 1184 tiles, {np.sort(rel)[min(len(rel) - 1, 148)]:.3f} s after all but 148; "
              f"last tiles {[(int(i), round(float(rel[i]), 3)) for i in np.argsort(rel)[:6]]}", flush=True)
    top = np.argsort(-tc)[:4]
    print(f"  tile busy time: sum {tc.sum() / 1.965e9:.1f} s = {tc.sum() / 1.965e9 / 1184:.3f} s per resident warp (1184), "
          f"max {tc.max() / 1.965e9:.3f} s, mean {tc.mean() / 1.965e9:.4f} s, p99 {np.percentile(tc, 99) / 1.965e9:.3f} s", flush=True)
    if os.environ.get("LGAR_DIAG_TILES"):
        print("  tiles", {int(k): round(float(tc[int(k)]) / 1.965e9, 4) for k in os.environ["LGAR_DIAG_TILES"].split(",")}, flush=True)
    print("  slowest tiles:", [(int(i), round(float(tc[i]) / 1.9e9, 2)) for i in top], "median tile s", round(float(np.median(tc)) / 1.9e9, 3), flush=True)
    print(f"B={B} T={T}: first {dt:.3f}s, best of {len(times)} passes {kms:.1f} ms (all: {[round(x) for x in times]}) -> {alive/kms*1e3:.4g} col-steps/s  alive col-steps={alive}  status hist={np.bincount(st, minlength=9).tolist()} counters={res.counters.cpu().numpy().tolist()}", flush=True)
