#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_gftt_detector.py tests/test_gpu_batch.py tests/test_gpu_sift_front_end.py -q -x -p no:cacheprovider > gpurun_out/feat_tests.log 2>&1; tail -6 gpurun_out/feat_tests.log
for w in c4 c5; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items(): print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:harris_kernel -c 1 -o gpurun_out/prof_harris python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_h.log 2>&1; tail -1 gpurun_out/ncu_h.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gauss_u8_stream -c 2 -o gpurun_out/prof_gauss_stream python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph --no-extra > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_l5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sep_f32_tma_kernel -s 10 -c 3 -o gpurun_out/prof_sift_sep python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_s.log 2>&1; tail -1 gpurun_out/ncu_s.log | cut -c1-200
