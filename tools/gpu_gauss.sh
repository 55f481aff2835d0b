#!/bin/bash
# u8 Gaussian / 8.8 sepFilter2D marching kernel: parity, timing, one ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_filters.py tests/test_gpu_hal.py -q -x -p no:cacheprovider > gpurun_out/gauss_tests.log 2>&1; tail -3 gpurun_out/gauss_tests.log
timeout 200 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2.json"))
print("c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
for k,v in d["per_op"].items():
    if "u8" in k and ("Gauss" in k or "sep" in k): print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gauss_u8_march -c 4 -o gpurun_out/prof_gauss_march python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_gauss.log 2>&1; tail -2 gpurun_out/ncu_gauss.log
