#!/bin/bash
# final-build record (round 2, second pass): every one-GPU bench workload, ncu launch list of the C3 Newton loop
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r3z_bench_c4_1gpu.json 2> gpurun_out/r3z_bench_c4_1gpu.err
python bench.py --dtype f32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3z_bench_c4_f32.json 2>/dev/null
python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r3z_bench_c2.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r3z_bench_c3.json 2>/dev/null
python bench.py --workload c1 --steps 3 --warmup 3 > gpurun_out/r3z_bench_c1.json 2>/dev/null
python bench.py --workload c5mb --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3z_bench_c5mb.json 2>/dev/null
for f in c4_1gpu c4_f32 c2 c5mb; do python - gpurun_out/r3z_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python - <<'PY'
import json
for f in ("c3","c1"):
    try:
        d=json.loads(open(f'gpurun_out/r3z_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d.get("roofline",{}).get("ms_per_newton_iteration"), d.get("cpu_baseline"), {k:v["us_per_evaluation"] for k,v in (d.get("us_per_evaluation") or {}).items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r3z_ncu_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[]
with open('gpurun_out/r3z_ncu_launches_c3.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
for row in r:
    try:
        v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    unit=row.get('Metric Unit','')
    if unit in ('ns','nsecond'): v/=1e3
    elif unit in ('ms','msecond'): v*=1e3
    k=row['Kernel Name'][:70]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print("%-70s n=%5d  %9.1f us  %5.1f %%"%(k,v[0],v[1],100*v[1]/tot))
PY
