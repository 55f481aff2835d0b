#!/bin/bash
# the 8-GPU evidence: batch driver tests over all devices, the in-process driver (C5 scaling, PCIe sharing probe, NCCL template broadcast),
# then the driver's own launch line (torchrun, one process per GPU) with the default workload
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_multi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_batch.py -q -x -p no:cacheprovider > gpurun_out/batch_tests_multi.log 2>&1; tail -4 gpurun_out/batch_tests_multi.log
timeout 900 python tools/batch_multi.py --quick > gpurun_out/batch_multi.txt 2>&1; cat gpurun_out/batch_multi.txt | grep -v "^GPU\|^NIC\|^$\|Legend\|^  " | head -40
N=$(python -c "import torch;print(torch.cuda.device_count())")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_default_${N}gpu.json 2> gpurun_out/bench_default_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default_${N}gpu.json"))
    print("N=$N c2", round(d["value"]), "Mpix/s", round(d["ms_per_step"],3), "ms/step; e2e", d.get("e2e"))
    for w,x in d["config"].get("extra_workloads",{}).items():
        print("  extra", w, {k:v for k,v in x.items() if k not in ("per_op","workload")})
except Exception as e: print("bench failed", e); print(open("gpurun_out/bench_default_${N}gpu.err").read()[-2500:])
PY
