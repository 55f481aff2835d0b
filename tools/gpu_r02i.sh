#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_features.py tests/test_gpu_geometry.py tests/test_gpu_hal.py tests/test_gpu_batch.py -q -p no:cacheprovider > gpurun_out/geo_tests.log 2>&1; tail -12 gpurun_out/geo_tests.log
for w in c3 c4 c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items(): print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cvt_kernel|warp_tile4" -c 10 -o gpurun_out/prof_c3_cvt python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log | cut -c1-200
