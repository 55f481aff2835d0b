#!/bin/bash
mkdir -p gpurun_out
B200CV_GFTT_TRACE=1 timeout 300 python tools/time_ops.py gftt_4k_noise > gpurun_out/time_gftt_trace.log 2>&1; tail -22 gpurun_out/time_gftt_trace.log
B200CV_GFTT_STREAMS=1 B200CV_GFTT_TRACE=1 timeout 300 python tools/time_ops.py gftt_4k_noise > gpurun_out/time_gftt_trace1.log 2>&1; tail -12 gpurun_out/time_gftt_trace1.log
for op in resize_lin_up resize_lin5k resize_cub5k hsv2bgr; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resize_sep|cvt_kernel" -c 1 -o gpurun_out/prof_$op python tools/prof_ops.py $op > gpurun_out/ncu_$op.log 2>&1; tail -1 gpurun_out/ncu_$op.log | cut -c1-160
done
