#!/bin/bash
# ncu evidence for the C3 Newton loop of the final build (the SM partition is switched off under the profiler: ncu cannot
# prepare kernels launched into green-context streams)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export AGP_CHOL_PARTITION=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r3y_ncu_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
with open('gpurun_out/r3y_ncu_launches_c3.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    unit=row.get('Metric Unit','')
    v = v/1e3 if unit.startswith('n') else (v*1e3 if unit.startswith('m') else v)
    k=row['Kernel Name'][:64]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
out=["kernel | launches | total us | share (serialised ncu launch list of bench.py --workload c3 --steps 1 --warmup 0)"]
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    out.append("%-64s n=%5d  %10.1f us  %5.1f %%  (%.1f us each)"%(k,v[0],v[1],100*v[1]/tot,v[1]/v[0]))
open('gpurun_out/r3y_ncu_launches_c3_summary.txt','w').write("\n".join(out)+"\n")
print("\n".join(out))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain_gemm32|potrf_trinv128' -s 40 -c 4 -o gpurun_out/r3y_ncu_full_chain python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r3y_ncu_full_chain.ncu-rep
