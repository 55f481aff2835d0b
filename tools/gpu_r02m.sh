#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sift_front_end.py tests/test_gpu_features.py -q -p no:cacheprovider -k "sift" > gpurun_out/sift_tests.log 2>&1; tail -6 gpurun_out/sift_tests.log
timeout 300 python tools/time_ops.py sift_detect_1080p > gpurun_out/time_sift.log 2>&1; tail -2 gpurun_out/time_sift.log
B200CV_SIFT_DESC_PATH=v1 timeout 300 python tools/time_ops.py sift_detect_1080p > gpurun_out/time_sift_v1.log 2>&1; tail -1 gpurun_out/time_sift_v1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sift_detect.csv python tools/time_ops.py sift_detect_1080p > gpurun_out/ncu_sd.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_sift_detect.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    n=r[ki].split("(")[0][:60]; agg[n]+=v; cnt[n]+=1
for k,v in agg.most_common(10): print("%-62s %9.1f us x%d"%(k,v,cnt[k]))
PY
