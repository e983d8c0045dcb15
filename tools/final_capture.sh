#!/bin/bash
# Round-end evidence run on the GPU box (under gpurun): tests, smoke, bench (both arms), ncu launch list of two
# steady-state steps and --set full captures of the top kernels.  Numbers printed under ncu are never bench values.
set -u
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_${R}.log 2>&1; tail -2 gpurun_out/pytest_gpu_${R}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${R}.log 2>&1; tail -1 gpurun_out/smoke_${R}.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_${R}_1gpu.json 2> gpurun_out/bench_${R}_1gpu.err; cat gpurun_out/bench_${R}_1gpu.json | cut -c1-400
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err; cat gpurun_out/bench_${R}_reference.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1100 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tcp_kernel|wgrad_tc_kernel|knn_sv_kernel' \
    -s 400 -c 8 -f -o gpurun_out/prof_${R} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${R}.log 2>&1
ls -la gpurun_out | tail -6
