#!/bin/bash
# Run on the GPU box (gpurun): launch list of one C4 step + `ncu --set full` captures of the dominant kernels.
# Every program is first run WITHOUT ncu (must exit 0); numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python tools/fwd_once.py bf16 > gpurun_out/plain_fwd.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_fwd.log; exit 1; }
python bench.py --steps 2 --warmup 1 --no-side-legs --no-cuda-graph > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
# launch list: every kernel of the timed steps with its device time (cold cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 2 --warmup 1 --no-side-legs --no-cuda-graph > gpurun_out/ncu_launch.log 2>&1
# full captures: split-operand forward / dX / dW at the C4 layer shape (third iteration of tools/fwd_once.py)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sampled_gemm_tc|dw_tc' -s 6 -c 3 -f -o gpurun_out/r02_tc \
    python tools/fwd_once.py bf16 > gpurun_out/ncu_tc.log 2>&1
# HBM-bound kernels of the step: multi-tensor Adam + KL, fused CE, KL reduction
timeout 900 ncu --set full --clock-control none -k regex:'shard_adam|ce_kernel|kl_fwd' -s 4 -c 4 -f -o gpurun_out/r02_hbm \
    python bench.py --steps 2 --warmup 1 --no-side-legs --no-cuda-graph > gpurun_out/ncu_hbm.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c4.csv
