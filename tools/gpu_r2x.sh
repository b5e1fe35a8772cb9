#!/bin/bash
# experiment: the 7-slice INT8 forward solve inside the Float64 mode -- parity of the M >= 768 cases and speed
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
AGP_F64_S1=i8 timeout 900 python -m pytest tests/test_gpu_svgp.py -m gpu -q -s -k "twin or m4096 or full_size or multi_block" > gpurun_out/r2x_tests.log 2>&1; grep -E "^\[|passed|failed|Error|assert" gpurun_out/r2x_tests.log | cut -c1-330 | head -20
AGP_F64_S1=i8 python bench.py --points 3031040 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2x_bench_s1i8.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_s1i8.json').read().strip().splitlines()[-1])
print("ms/step=%.1f"%d["ms_per_step"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d["correctness"])
PY
