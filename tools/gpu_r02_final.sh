#!/bin/bash
# end-of-round verification: the full GPU suite, smoke, the reference arm, the default bench line, and the ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/full_gpu_tests.log 2>&1; tail -4 gpurun_out/full_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default.json"))
    print("c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    for w,x in d["config"].get("extra_workloads",{}).items():
        print("  extra", w, round(x["value"]), round(x["ms_per_step"],3), "ms; e2e", round(x["e2e"]["value"]))
    r=json.load(open("gpurun_out/bench_reference.json")); print("reference arm", round(r["value"],1), r["unit"])
except Exception as e: print("bench failed", e); print(open("gpurun_out/bench_default.err").read()[-2500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-e2e --no-graph > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_bench_c2.csv
