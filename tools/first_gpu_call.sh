#!/bin/bash
# First GPU call after round 1: run the parity tests of everything that was written without GPU time (tests/test_gpu_zz_*.py, marked
# xfail-until-first-run), then time those ops.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# Outputs (merged back under gpurun_out/): zz_tests.log (-rxX lists every XPASS / XFAIL with its message), time_ops.log
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k zz -rxX -p no:cacheprovider > gpurun_out/zz_tests.log 2>&1
tail -5 gpurun_out/zz_tests.log
timeout 240 python tools/time_ops.py > gpurun_out/time_ops.log 2>&1
tail -30 gpurun_out/time_ops.log
