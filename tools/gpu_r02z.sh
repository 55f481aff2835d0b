#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_color_depths.py tests/test_gpu_hal.py -q -p no:cacheprovider > gpurun_out/hsv_tests.log 2>&1; tail -12 gpurun_out/hsv_tests.log
