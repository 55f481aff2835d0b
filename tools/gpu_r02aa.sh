#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filters.py -q -p no:cacheprovider -k "gaussian_u8" > gpurun_out/gb_tests.log 2>&1; tail -8 gpurun_out/gb_tests.log
B200CV_GAUSS_U8_PATH=binomial timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e > gpurun_out/bench_c2_n.json 2> gpurun_out/bench_c2_n.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c2_n.json"))
    print("c2", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "GaussianBlur_u8" in k and any(k.endswith(x) for x in ("_k3","_k5","_k7")): print("   %-40s %8.4f ms  frac %.3f"%(k, v["ms"], v.get("frac_hbm",0)))
except Exception as e: print("failed",e); print(open("gpurun_out/bench_c2_n.err").read()[-1500:])
PY
