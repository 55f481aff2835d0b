#!/bin/bash
# cv-typed surface test + ncu evidence for the C2 step: launch list (shares) and --set full captures of the kernels that dominate it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_hal.py -q -x -p no:cacheprovider > gpurun_out/hal_tests.log 2>&1; tail -5 gpurun_out/hal_tests.log
B="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph --no-extra"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $B > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"filter2d_tc_f32_kernel|ff_pad_split" -c 4 -o gpurun_out/prof_f2d_tc_f32 $B > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gauss_u8_stream_kernel<3|gauss_u8_stream_kernel<5" -c 2 -o gpurun_out/prof_gauss_stream $B > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
