#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of one bench step pair and a --set full capture of
# the top kernels.  Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
ROUND=${1:-r01}
# 1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${ROUND}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${ROUND}.log 2>&1
# 2) full capture of the heaviest kernels: conv (eval + train), wgrad, knn_sv
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'conv_kernel|wgrad_kernel|knn_sv_kernel' -s 400 -c 12 -f -o gpurun_out/prof_${ROUND} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${ROUND}.log 2>&1
ls -la gpurun_out | tail -8
