#!/bin/bash
# iteration run: parity of the kernels under work, per-op tables, ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filters.py tests/test_gpu_geometry.py tests/test_gpu_features.py tests/test_gpu_hal.py -q -x -p no:cacheprovider > gpurun_out/iter_tests.log 2>&1; tail -15 gpurun_out/iter_tests.log
for w in c2 c3 c4; do
timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step")
    for k,v in d["per_op"].items():
        if "$w"!="c2" or ("u8" in k and ("Gauss" in k or "sep" in k) and "k1" not in k and "k2" not in k) or "filter2D_f32" in k: print("   %-40s %8.4f ms  frac %.3f"%(k,v["ms"],v["frac_hbm"]))
except Exception as e: print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gauss_u8_stream_kernel<3|gauss_u8_stream_kernel<5|filter2d_tc_f32" -c 10 -o gpurun_out/prof_c2 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resize_sep -c 4 -o gpurun_out/prof_resize_sep python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_resize.log 2>&1; tail -2 gpurun_out/ncu_resize.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mt_fused -c 1 -o gpurun_out/prof_mt python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-graph > gpurun_out/ncu_mt.log 2>&1; tail -2 gpurun_out/ncu_mt.log
