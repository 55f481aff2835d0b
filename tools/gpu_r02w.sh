#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default.json"))
    print("c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "clocks", d["clocks"], "roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","bound","achieved","peak","frac","traffic","share_of_step")})
    print("cpu_baseline", d.get("cpu_baseline"))
    for w,x in d["config"].get("extra_workloads",{}).items():
        print("  extra", w, round(x["value"]), round(x["ms_per_step"],3), "ms; e2e", round(x["e2e"]["value"]))
except Exception as e: print("bench failed", e); print(open("gpurun_out/bench_default.err").read()[-2500:])
PY
echo "bench wall seconds: $SECONDS"
