#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload cfg4 --steps 10 --warmup 3 --dump-ops gpurun_out/r2c_train_ops_b32.txt > gpurun_out/r2c_train_b32.json 2> gpurun_out/r2c_train_b32.err; tail -c 400 gpurun_out/r2c_train_b32.err
python - <<PY
import json,re,collections
d=json.loads(open("gpurun_out/r2c_train_b32.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline_parts"))
agg=collections.OrderedDict()
for l in open('gpurun_out/r2c_train_ops_b32.txt'):
    m=re.match(r'^((?:fwd|bwd):.*?)\s+([0-9.]+) us\s+([0-9.]+) GFLOP',l)
    if not m: continue
    k=m.group(1); us=float(m.group(2)); agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=us
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:34s} x{v[0]:3d} {v[1]/1e3:9.2f} ms")
print("sum", sum(v[1] for v in agg.values())/1e3)
PY
grep "wgrad_tc" gpurun_out/r2c_train_ops_b32.txt
