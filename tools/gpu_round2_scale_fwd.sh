# forward-only weak-scaling line (no e2e / forward+gradient legs: a third of the GPU time of the full bench)
set -x
N=${1:-8}
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --grad-columns 0 ) > gpurun_out/r2b_bench_fwd_n$N.json 2> gpurun_out/r2b_bench_fwd_n$N.err
tail -c 300 gpurun_out/r2b_bench_fwd_n$N.err
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/r2b_bench_fwd_n$N.json") if x.startswith("{")][-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "wall_s")}, d["per_rank"]["pass_ms"], d["per_rank"]["achieved_tflops"])
PY
