#!/bin/bash
# batch driver bring-up on one GPU: its tests, then the default bench line (C2 + extra workloads + e2e through the driver) and the C5 line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_hal.py -q -x -p no:cacheprovider > gpurun_out/batch_tests.log 2>&1; tail -15 gpurun_out/batch_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -3 gpurun_out/bench_c5.err
python - <<PY
import json
for f in ("bench_default","bench_c5"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step; e2e", d.get("e2e"))
        print("  roofline", d["roofline"]); print("  cpu", d.get("cpu_baseline"))
        for w,x in d["config"].get("extra_workloads",{}).items():
            print("  extra", w, {k:v for k,v in x.items() if k!="per_op"})
    except Exception as e: print(f, "failed", e)
PY
