#!/bin/bash
# regression check of the headline workloads after the shared-memory carve-out / Cholesky changes
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3l_bench_c4_1gpu.json 2> gpurun_out/r3l_bench_c4_1gpu.err
AGP_CARVEOUT=0 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3l_bench_c4_nocarve.json 2>/dev/null
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3l_bench_c2.json 2>/dev/null
python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3l_bench_c1.json 2>/dev/null
for f in c4_1gpu c4_nocarve c2; do python - gpurun_out/r3l_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3l_bench_c1.json').read().strip().splitlines()[-1])
print("c1", d["value"], d["unit"], d.get("us_per_evaluation"))
PY
