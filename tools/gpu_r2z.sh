#!/bin/bash
# final-build record: GPU suite, smoke, every bench workload of one GPU (FP64 headline, Float32 mode, C2, C3, C1), ncu launch list
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2z_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2z_tests.log
grep -E "passed|failed|FAILED|EXIT" gpurun_out/r2z_tests.log | tail -6
cp gpurun_out/parity_errors.json gpurun_out/r2z_parity_errors.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench_c4_1gpu.json 2> gpurun_out/r2z_bench_c4_1gpu.err
python bench.py --dtype f32 --steps 5 --warmup 3 > gpurun_out/r2z_bench_c4_f32.json 2>/dev/null
python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r2z_bench_c2.json 2>/dev/null
python bench.py --workload c2 --dtype f32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_c2_f32.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2z_bench_c3.json 2>/dev/null
python bench.py --workload c1 --steps 3 --warmup 3 > gpurun_out/r2z_bench_c1.json 2>/dev/null
python bench.py --workload c5mb --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_c5mb.json 2>/dev/null
for f in c4_1gpu c4_f32 c2 c2_f32 c5mb; do python - gpurun_out/r2z_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], "exec=%.3f"%d["roofline"]["whole_step"]["frac_of_fp64_peak_executed"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python - <<'PY'
import json
for f in ("c3","c1"):
    try:
        d=json.loads(open(f'gpurun_out/r2z_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d.get("yardstick"), d.get("us_per_evaluation"))
    except Exception as e:
        print(f, "FAILED", e)
PY
