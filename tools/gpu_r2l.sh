#!/bin/bash
# kgrad from stored kernel values (S1 out of place) vs the recomputing kernel; full-size C4 line, C2, C3 (with the cuSOLVER yard-stick)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2l_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2l_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2l_tests.log | tail -15
cp gpurun_out/parity_errors.json gpurun_out/r2l_parity_errors.json 2>/dev/null
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2l_bench_kfast.json 2> gpurun_out/r2l_bench_kfast.err
AGP_KGRAD_FAST=0 $B > gpurun_out/r2l_bench_kslow.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench_c4_1gpu.json 2> gpurun_out/r2l_bench_c4_1gpu.err
python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r2l_bench_c2.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2l_bench_c3.json 2> gpurun_out/r2l_bench_c3.err
python bench.py --dtype f32 --steps 3 --warmup 3 > gpurun_out/r2l_bench_c4_f32.json 2>/dev/null
for f in gpurun_out/r2l_bench_k*.json gpurun_out/r2l_bench_c4_1gpu.json gpurun_out/r2l_bench_c2.json gpurun_out/r2l_bench_c4_f32.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], d["roofline"]["kernel"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2l_bench_c3.json').read().strip().splitlines()[-1])
    print("c3", d["value"], d["roofline"]["ms_per_newton_iteration"], d["roofline"]["achieved"], d["yardstick"])
except Exception as e:
    print("c3 FAILED", e)
PY
