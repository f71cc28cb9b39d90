# after the flat-run jump in check_column_mass: GPU tests, the tail tile, the two shards, the default bench
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_tests.log 2>&1
tail -n 4 gpurun_out/r2b_tests.log
timeout 300 python tests/gpu_diag_tile.py 4 2633 > gpurun_out/r2b_tile_2633.log 2>&1
grep "window ms\|slowest column alone" gpurun_out/r2b_tile_2633.log | cut -c1-400
export LGAR_DIAG_BALANCE=1 LGAR_DIAG_REPS=1 LGAR_DIAG_SHARED=1
LGAR_DIAG_RANK=4 timeout 300 python tests/gpu_diag.py 125000x8760 > gpurun_out/r2b_tail_rank4.log 2>&1
grep "scheduler\|tail\|best of" gpurun_out/r2b_tail_rank4.log | cut -c1-600
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/r2b_bench_n1.json") if x.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_s")}, d["roofline"]["frac"], d["e2e"]["value"], d["fwd_grad"]["value"], d.get("parity"), d["config"]["status_histogram"])
PY
