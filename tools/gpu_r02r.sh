#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_features.py -q -p no:cacheprovider -k "not sift" > gpurun_out/resize_tests.log 2>&1; tail -4 gpurun_out/resize_tests.log
timeout 300 python tools/time_ops.py gftt_4k_noise gftt_4k_smooth > gpurun_out/time_gftt.log 2>&1; tail -2 gpurun_out/time_gftt.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e > gpurun_out/bench_c3_n.json 2> gpurun_out/bench_c3_n.err
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e > gpurun_out/bench_c4_n.json 2> gpurun_out/bench_c4_n.err
python - <<PY
import json
for w in ("c3","c4"):
    try:
        d=json.load(open("gpurun_out/bench_%s_n.json"%w))
        print(w, round(d["ms_per_step"],3), "ms/step")
        for k,v in d["per_op"].items():
            if "resize" in k or w=="c4": print("   %-40s %8.4f ms  frac %.3f"%(k, v["ms"], v.get("frac_hbm",0)))
    except Exception as e: print(w,"failed",e); print(open("gpurun_out/bench_%s_n.err"%w).read()[-1500:])
PY
