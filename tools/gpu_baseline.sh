#!/bin/bash
# Round baseline on one B200: the whole -m gpu suite, then one bench line per workload and the timings of the ops outside the workloads.
#   gpurun --timeout 900 -- 'bash tools/gpu_baseline.sh'
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
for w in c2 c3 c4 c5; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", d.get("e2e",{}).get("value"))
    for k,v in d["per_op"].items(): print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e)
PY
done
timeout 240 python tools/time_ops.py > gpurun_out/time_ops.log 2>&1
cat gpurun_out/time_ops.log
