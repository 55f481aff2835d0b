#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_features.py tests/test_gpu_geometry.py tests/test_gpu_filters.py tests/test_gpu_gftt_detector.py -q -p no:cacheprovider > gpurun_out/all_tests.log 2>&1; tail -8 gpurun_out/all_tests.log
for w in c2 c3 c4 c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "$w"!="c2" or "_k3" in k or "_k5" in k or "_k9" in k: print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"harris_fast" -c 1 -o gpurun_out/prof_harris_fast python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_h.log 2>&1; tail -1 gpurun_out/ncu_h.log | cut -c1-200
