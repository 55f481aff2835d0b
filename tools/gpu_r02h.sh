#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filters.py tests/test_gpu_features.py -q -x -p no:cacheprovider  > gpurun_out/sep_tests.log 2>&1; tail -3 gpurun_out/sep_tests.log
for mode in walk tile; do
if [ $mode = tile ]; then export B200CV_SEP_DOG_TILE=1; else unset B200CV_SEP_DOG_TILE; fi
timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_$mode.json 2> gpurun_out/bench_c5_$mode.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c5_$mode.json"))
    print("$mode", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step", {k:v["ms"] for k,v in d["per_op"].items()})
except Exception as e: print("$mode failed", e); print(open("gpurun_out/bench_c5_$mode.err").read()[-1500:])
PY
done
unset B200CV_SEP_DOG_TILE
timeout 200 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2.json"))
print("c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
for k,v in d["per_op"].items():
    if ("Gaussian" in k or "sep" in k) and ("k11" in k or "k15" in k or "k21" in k or "k31" in k or "_k9" in k): print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
PY

