# is the forward pass bound by the longest per-tile chain?  tile busy times of rank 0's and rank 4's shard (full year)
set -x
mkdir -p gpurun_out
export LGAR_DIAG_BALANCE=1 LGAR_DIAG_REPS=1 LGAR_DIAG_SHARED=1
python tests/gpu_diag.py 125000x8760 > gpurun_out/r2_tail_rank0.log 2>&1
LGAR_DIAG_RANK=4 python tests/gpu_diag.py 125000x8760 > gpurun_out/r2_tail_rank4.log 2>&1
tail -n 6 gpurun_out/r2_tail_rank0.log; tail -n 6 gpurun_out/r2_tail_rank4.log
