# config C3 (Bushland record, 100,000-member parameter ensemble, forward only) as a full bench line on one GPU
set -x
mkdir -p gpurun_out
( time python bench.py --workload c3 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
tail -c 300 gpurun_out/r2_bench_c3.err
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/r2_bench_c3.json") if x.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_s")}, d["roofline"]["frac"], d["e2e"]["value"], d.get("parity"), d["config"]["status_histogram"])
PY
