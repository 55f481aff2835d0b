#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/full_gpu_tests.log 2>&1; tail -6 gpurun_out/full_gpu_tests.log
timeout 300 python tools/time_ops.py cvt_f32_bgr2gray_4k cvt_f32_bgr2yuv_4k cvt_f32_yuv2bgr_4k cvt_u16_bgr2gray_4k cvt_u16_bgr2ycrcb_4k cvt_u16_bgr2rgb_4k > gpurun_out/time_cvt_depth.log 2>&1; tail -7 gpurun_out/time_cvt_depth.log
