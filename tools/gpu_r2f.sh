#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r2f_f32tests.log 2>&1; grep -E "^\[f32|passed|failed|Error|error" gpurun_out/r2f_f32tests.log | head -30
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_f32.py > gpurun_out/r2f_tests.log 2>&1; tail -4 gpurun_out/r2f_tests.log
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B --dtype f32 > gpurun_out/r2f_bench_f32.json 2> gpurun_out/r2f_bench_f32.err; tail -2 gpurun_out/r2f_bench_f32.err
$B > gpurun_out/r2f_bench_f64.json 2>/dev/null
for f in gpurun_out/r2f_bench_f32.json gpurun_out/r2f_bench_f64.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "ms/step=%.1f"%d["ms_per_step"], "elbo", d["elbo"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
AGP_B200_LIB=$PWD/build/libagp_timing.so python tools/c1_phases.py 2>&1 | tail -2 | tee gpurun_out/r2f_c1_phases.txt
