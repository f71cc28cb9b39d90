# weak-scaling bench exactly as the driver launches it: bash tools/gpu_round2_scale.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/gpu_n$N.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 400 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
l = [x for x in open("gpurun_out/r2_bench_n$N.json") if x.startswith("{")]
d = json.loads(l[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "wall_s")}, d["per_rank"], d["fwd_grad"]["value"], d["e2e"]["value"])
PY
