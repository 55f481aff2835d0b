#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filters.py tests/test_gpu_hal.py -q -x -p no:cacheprovider -k "tensor_core or tma_path or filter2d or cv_typed" > gpurun_out/f2d_tests.log 2>&1; tail -8 gpurun_out/f2d_tests.log
timeout 200 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c2.json"))
    print("c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "filter2D" in k: print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("failed", e); print(open("gpurun_out/bench_c2.err").read()[-1500:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"filter2d_tc_f32_kernel" -c 4 -o gpurun_out/prof_f2d_tc_f32_v2 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph --no-extra > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log | cut -c1-200
