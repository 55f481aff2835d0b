#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
for w in c3 c4 c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "$w"!="c3" or "cvt" in k or "LINEAR" in k: print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
timeout 300 python tools/time_ops.py integral_4k integral_sq_4k gauss_u16_k5 gauss_u16_k15 sift_detect_1080p > gpurun_out/time_ops2.log 2>&1; cat gpurun_out/time_ops2.log | tail -8
B200CV_INTEGRAL_PATH=v1 B200CV_GAUSS_U16_PATH=v1 timeout 300 python tools/time_ops.py integral_4k integral_sq_4k gauss_u16_k5 gauss_u16_k15 > gpurun_out/time_ops2_v1.log 2>&1; cat gpurun_out/time_ops2_v1.log | tail -5
