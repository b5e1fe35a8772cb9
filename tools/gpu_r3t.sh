#!/bin/bash
# experiment: S6 of the Float64 sweep on the INT8 tensor path (AGP_F64_S6=i8)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export AGP_F64_S6=i8
timeout 600 python -m pytest tests/test_gpu_svgp.py -m gpu -q -s > gpurun_out/r3t_tests_s6i8.log 2>&1; echo "EXIT $?" >> gpurun_out/r3t_tests_s6i8.log
grep -E "passed|failed|FAILED|EXIT|Error|error" gpurun_out/r3t_tests_s6i8.log | tail -12
grep -E "M=1024|M=512|M=2048|M=4096" gpurun_out/r3t_tests_s6i8.log | head -12
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r3t_bench_c4_s6i8.json 2> gpurun_out/r3t_bench_c4_s6i8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r3t_bench_c4_s6i8.json').read().strip().splitlines()[-1])
    print("c4 s6=i8 value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
    print((d.get("cpu_baseline") or {}).get("parity") or d.get("parity") or {k:v for k,v in d.items() if 'check' in k or 'parity' in k})
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/r3t_bench_c4_s6i8.err').read()[-2000:])
PY
