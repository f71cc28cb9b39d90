# dealt Geff batches: full GPU suite + forward-only full-year line (A/B against profiles/bench_r2b_n1.json: 49.28 M, 19.27 s)
set -x
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-e2e --grad-columns 0 > gpurun_out/r2c_ab_deal.json 2> gpurun_out/r2c_ab_deal.err
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/r2c_ab_deal.json") if x.startswith("{")][-1])
print("DEAL", d["value"], d["ms_per_step"] * d["steps"], d["roofline"]["frac"])
PY
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c_tests.log 2>&1
tail -n 5 gpurun_out/r2c_tests.log
