# single-GPU A/B lines of the forward pass (full C4 shard, full year): shard of rank 1, no pipelining, identity placement
set -x
mkdir -p gpurun_out
F="--no-cpu-baseline --no-e2e --grad-columns 0"
python bench.py $F --shard-rank 1 > gpurun_out/r2_ab_shard1.json 2> gpurun_out/r2_ab.err
python bench.py $F --no-pipeline  > gpurun_out/r2_ab_nopipeline.json 2>> gpurun_out/r2_ab.err
python bench.py $F --no-balance   > gpurun_out/r2_ab_nobalance.json 2>> gpurun_out/r2_ab.err
python - <<PY
import json
for n in ("shard1", "nopipeline", "nobalance"):
    d = json.loads([x for x in open(f"gpurun_out/r2_ab_{n}.json") if x.startswith("{")][-1])
    print(n, d["value"], d["ms_per_step"] * d["steps"], d["roofline"]["achieved"], d["config"]["alive_column_steps_per_gpu"])
PY
