#!/bin/bash
# compute-sanitizer memcheck / racecheck over the small-image GPU parity tests (the 4K / 8K cases are left out: the tool slows kernels 10-100x)
mkdir -p gpurun_out
SEL='not 4k and not 8k and not 4K and not 8K and not preselection and not batch and not surface and not hal'
for mod in filters geometry color features yuv_wire_formats lab integral gaussian_u16 resize_area resize_exact resize_lanczos bayer match_template_mask; do
  timeout 500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck_$mod.log python -m pytest tests/test_gpu_$mod.py -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/memcheck_$mod.out 2>&1
  echo "memcheck $mod rc=$? $(tail -1 gpurun_out/memcheck_$mod.out | cut -c1-80) | $(grep -c 'Invalid\|out of bounds\|misaligned' gpurun_out/memcheck_$mod.log) findings | $(grep 'ERROR SUMMARY' gpurun_out/memcheck_$mod.log | tail -1)"
done
for mod in features filters geometry; do
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/racecheck_$mod.log python -m pytest tests/test_gpu_$mod.py -q -x -p no:cacheprovider -k "$SEL and not sift" > gpurun_out/racecheck_$mod.out 2>&1
  echo "racecheck $mod rc=$? $(tail -1 gpurun_out/racecheck_$mod.out | cut -c1-80) | $(grep 'RACECHECK SUMMARY' gpurun_out/racecheck_$mod.log | tail -1)"
done
