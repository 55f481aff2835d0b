#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py -q -p no:cacheprovider -k "resize" > gpurun_out/resize_tests.log 2>&1; tail -4 gpurun_out/resize_tests.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e > gpurun_out/bench_c3_n.json 2> gpurun_out/bench_c3_n.err
python - <<PY
import json
for w in ("c3",):
    try:
        d=json.load(open("gpurun_out/bench_%s_n.json"%w))
        print(w, round(d["ms_per_step"],3), "ms/step")
        for k,v in d["per_op"].items():
            if "resize" in k: print("   %-40s %8.4f ms  frac %.3f"%(k, v["ms"], v.get("frac_hbm",0)))
    except Exception as e: print(w,"failed",e); print(open("gpurun_out/bench_%s_n.err"%w).read()[-1500:])
PY
for op in warp_cub warp_lin; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"warp_tile4" -c 1 -o gpurun_out/prof_$op python tools/prof_ops.py $op > gpurun_out/ncu_$op.log 2>&1; tail -1 gpurun_out/ncu_$op.log | cut -c1-160
done
