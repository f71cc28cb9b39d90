set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -3 gpurun_out/r2_tests.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 300 gpurun_out/r2_bench_ref.err
# launch list of a reduced command (same code path, 1280 rows)
PC="python bench.py --nsteps 1280 --steps 10 --warmup 3 --no-cpu-baseline"
$PC > gpurun_out/r2_profiled_command.json 2> gpurun_out/r2_profiled_command.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $PC > gpurun_out/r2_ncu_launch.log 2>&1
# full capture forward kernel (4th forward launch = a timed-ish segment of 128 rows) and the reverse kernel
ncu --set full --clock-control none --import-source on -k regex:lgar_forward_kernel -s 6 -c 1 -o gpurun_out/r2_fwd $PC --no-e2e --grad-columns 0 > gpurun_out/r2_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lgar_backward_kernel -c 1 -o gpurun_out/r2_bwd python bench.py --nsteps 256 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_bwd.log 2>&1
ls -la gpurun_out
