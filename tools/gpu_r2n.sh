#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/r2n_train_launches.csv python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/r2n_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2n_train_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split("(")[0][-60:]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[vi].replace(",",""))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:45]: print(f"{k:62s} x{v[0]:4d} {v[1]/1e3:9.1f} us")
print("total", tot/1e3, "us over", sum(v[0] for v in agg.values()), "launches")
PY
