#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -q -x -p no:cacheprovider -k "sift" > gpurun_out/sift_tests.log 2>&1; tail -3 gpurun_out/sift_tests.log
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sep_f32_tma_kernel -s 10 -c 3 -o gpurun_out/prof_sift_sep_walk python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_s.log 2>&1; tail -1 gpurun_out/ncu_s.log | cut -c1-200
