#!/bin/bash
# AGP_COMPUTE_F64_EMU with S2 on the INT8 engine as well: parity tests, C4 line with and without the S2 part
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f64emu.py -m gpu -q -s > gpurun_out/r3w_tests_f64emu.log 2>&1; echo "EXIT $?" >> gpurun_out/r3w_tests_f64emu.log
grep -E "passed|failed|FAILED|EXIT|Error|assert" gpurun_out/r3w_tests_f64emu.log | tail -8
grep -E "^\[f64emu" gpurun_out/r3w_tests_f64emu.log | cut -c1-230
python bench.py --dtype f64emu --steps 4 --warmup 3 > gpurun_out/r3w_bench_c4_f64emu.json 2> gpurun_out/r3w_bench_c4_f64emu.err
python - gpurun_out/r3w_bench_c4_f64emu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, json.dumps(d.get("correctness"))[:600])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open('gpurun_out/r3w_bench_c4_f64emu.err').read()[-1500:])
PY
