# final build: smoke, C3 line, launch list + full ncu capture of the forward kernel (same recipe as tools/gpu_round2_run1.sh)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -n 2 gpurun_out/r2b_smoke.log
( time python bench.py --workload c3 --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
PC="python bench.py --nsteps 1280 --steps 10 --warmup 3 --no-cpu-baseline"
$PC > gpurun_out/r2b_profiled_command.json 2> gpurun_out/r2b_profiled_command.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv $PC > gpurun_out/r2b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lgar_forward_kernel -s 6 -c 1 -o gpurun_out/r2b_fwd $PC --no-e2e --grad-columns 0 > gpurun_out/r2b_ncu_fwd.log 2>&1
ls -la gpurun_out | tail -n 12
