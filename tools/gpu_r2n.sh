#!/bin/bash
# streaming kgrad (Kf read back, SE needs no second matrix), 16-wide register Cholesky in the diagonal-block kernel: tests, phase timing A/B, C3, C4, C2
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 60 ./build/pt_timing > gpurun_out/r2n_pt_timing.txt 2>&1; echo "--- 32-wide (round 1) ---" >> gpurun_out/r2n_pt_timing.txt; timeout 60 ./build/pt_timing_b32 >> gpurun_out/r2n_pt_timing.txt 2>&1
cat gpurun_out/r2n_pt_timing.txt
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2n_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2n_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2n_tests.log | tail -15
cp gpurun_out/parity_errors.json gpurun_out/r2n_parity_errors.json 2>/dev/null
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2n_bench_new.json 2> gpurun_out/r2n_bench_new.err
AGP_KGRAD_FAST=0 $B > gpurun_out/r2n_bench_kslow.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c3.json 2> gpurun_out/r2n_bench_c3.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_c4_1gpu.json 2> gpurun_out/r2n_bench_c4_1gpu.err
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c2.json 2>/dev/null
for f in gpurun_out/r2n_bench_new.json gpurun_out/r2n_bench_kslow.json gpurun_out/r2n_bench_c4_1gpu.json gpurun_out/r2n_bench_c2.json; do python - "$f" <<'PY'
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
    d=json.loads(open('gpurun_out/r2n_bench_c3.json').read().strip().splitlines()[-1])
    print("c3", d["value"], d["roofline"]["ms_per_newton_iteration"], d["roofline"]["achieved"], d["yardstick"], d["lml"])
except Exception as e:
    print("c3 FAILED", e)
PY
