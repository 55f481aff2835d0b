#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py gftt_4k_noise gftt_4k_smooth > gpurun_out/time_gftt.log 2>&1; tail -2 gpurun_out/time_gftt.log
B200CV_GFTT_PRESELECT=0 timeout 300 python tools/time_ops.py gftt_4k_noise gftt_4k_smooth > gpurun_out/time_gftt_v1.log 2>&1; tail -2 gpurun_out/time_gftt_v1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gftt.csv python tools/time_ops.py gftt_4k_noise > gpurun_out/ncu_gftt.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_gftt.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    n=r[ki].split("(")[0][:70]; agg[n]+=v; cnt[n]+=1
for k,v in agg.most_common(14): print("%-72s %9.1f us x%d"%(k,v,cnt[k]))
PY
