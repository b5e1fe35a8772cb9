#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r2g_f32tests.log 2>&1; grep -E "^\[f32|passed|failed|Error|error" gpurun_out/r2g_f32tests.log | head -30
AGP_F32_S5=tf32 timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s 2>&1 | grep -E "^\[f32|passed|failed" | head
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_f32.py > gpurun_out/r2g_tests.log 2>&1; tail -4 gpurun_out/r2g_tests.log
python bench.py --dtype f32 --steps 3 --warmup 3 > gpurun_out/r2g_bench_c4_f32.json 2> gpurun_out/r2g_bench_c4_f32.err; tail -2 gpurun_out/r2g_bench_c4_f32.err
python - gpurun_out/r2g_bench_c4_f32.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["dtype"], "value %.3e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d.get("correctness"))
PY
