#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_color_depths.py tests/test_gpu_color.py -q -x -p no:cacheprovider > gpurun_out/depth_tests.log 2>&1; tail -12 gpurun_out/depth_tests.log
timeout 300 python tools/time_ops.py cvt_f32_bgr2gray_4k cvt_f32_bgr2yuv_4k cvt_f32_yuv2bgr_4k cvt_u16_bgr2gray_4k cvt_u16_bgr2ycrcb_4k cvt_u16_bgr2rgb_4k > gpurun_out/time_cvt_depth.log 2>&1; tail -7 gpurun_out/time_cvt_depth.log
