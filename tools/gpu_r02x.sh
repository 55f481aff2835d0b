#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bayer.py tests/test_gpu_color_depths.py tests/test_gpu_color.py -q -p no:cacheprovider > gpurun_out/bayer_tests.log 2>&1; tail -8 gpurun_out/bayer_tests.log
