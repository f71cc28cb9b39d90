"""Standalone parity report (run on the GPU box): prints, per golden case, the worst
tolerance-normalised error of the CUDA path and the first mismatch.  Not a pytest file."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import golden_names, load_golden, max_excess
from gpu_common import run_golden_on_gpu

FLUX = ("runoff", "percolation", "AET", "infiltration", "ending_volume", "ponded_water", "giuh_runoff")
names = sys.argv[1:] or golden_names()
for name in names:
    g = load_golden(name)
    t0 = time.time()
    dump = "fronts" in g.files
    r = run_golden_on_gpu(g, copies=2, dump=dump)
    dt = time.time() - t0
    T = g["forcing"].shape[0]
    cs = int(g["crash_step"]); n_ok = T if cs < 0 else cs
    nf_eq = np.array_equal(r["nfronts"][:n_ok, 0], g["nfronts"][:n_ok])
    worst = {k: max_excess(r[k][:n_ok, 0], g[k][:n_ok]) for k in FLUX}
    wk = max(worst, key=worst.get)
    line = f"{name:26s} T={T:5d} {dt:6.2f}s status={r['status'][0]} crash={r['crash_step'][0]} (ref {cs}) nf_eq={nf_eq} worst={wk}:{worst[wk]:.3g}"
    if dump and nf_eq:
        line += f" fronts:{max_excess(r['fronts'][:n_ok,:,:,0], g['fronts'][:n_ok]):.3g}"
        line += f" lay_eq={np.array_equal(r['front_layer'][:n_ok,:,0], g['front_layer'][:n_ok])}"
        line += f" cnt={r['counters'].tolist()}"
    if not nf_eq:
        bad = np.nonzero(r["nfronts"][:n_ok, 0] != g["nfronts"][:n_ok])[0]
        line += f" first_nf_mismatch@{bad[0]}: {r['nfronts'][bad[0],0]} vs {g['nfronts'][bad[0]]}"
    print(line, flush=True)
