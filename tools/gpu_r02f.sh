#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filters.py tests/test_gpu_features.py tests/test_gpu_sift_front_end.py tests/test_gpu_hal.py -q -x -p no:cacheprovider > gpurun_out/sep_tests.log 2>&1; tail -6 gpurun_out/sep_tests.log
for w in c2 c4 c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "$w"!="c2" or "k11" in k or "k15" in k or "k21" in k or "k31" in k or "_k5" in k: print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
