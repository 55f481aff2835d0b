#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_hal.py -q -p no:cacheprovider > gpurun_out/geo_tests.log 2>&1; tail -8 gpurun_out/geo_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default.json"))
    print("default", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step; e2e", {k:v for k,v in d.get("e2e",{}).items() if k!="api"})
    print("  roofline", d["roofline"]); print("  cpu", d.get("cpu_baseline")); print("  clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
    for w,x in d["config"].get("extra_workloads",{}).items():
        print("  extra", w, {k:v for k,v in x.items() if k not in ("per_op","workload")})
except Exception as e: print("failed", e); print(open("gpurun_out/bench_default.err").read()[-2000:])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
