#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resize_lanczos.py -q -p no:cacheprovider > gpurun_out/lanczos_tests.log 2>&1; tail -6 gpurun_out/lanczos_tests.log
timeout 300 python tools/time_ops.py lanczos4_8k_to_5k > gpurun_out/time_lanczos.log 2>&1; tail -1 gpurun_out/time_lanczos.log
B200CV_RESIZE_LANCZOS_PATH=v1 timeout 300 python tools/time_ops.py lanczos4_8k_to_5k > gpurun_out/time_lanczos_v1.log 2>&1; tail -1 gpurun_out/time_lanczos_v1.log
