#!/bin/bash
# tf32 engine: elected-lane issue + n-fast tile order; Float32-mode tests and bench; INT8 emulation engine test
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 120 ./build/tf32x3_test timing > gpurun_out/r2r_tf32x3.jsonl 2>&1; echo "exit $?" >> gpurun_out/r2r_tf32x3.jsonl; tail -3 gpurun_out/r2r_tf32x3.jsonl | cut -c1-400
timeout 120 ./build/i8emu_test timing > gpurun_out/r2r_i8emu.jsonl 2>&1; echo "exit $?" >> gpurun_out/r2r_i8emu.jsonl; grep -E "sweep|exit" gpurun_out/r2r_i8emu.jsonl | cut -c1-90,330-600
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r2r_f32tests.log 2>&1; grep -E "^\[f32|passed|failed|Error|error" gpurun_out/r2r_f32tests.log | cut -c1-300 | head -12
python bench.py --dtype f32 --steps 3 --warmup 3 > gpurun_out/r2r_bench_c4_f32.json 2> gpurun_out/r2r_bench_c4_f32.err
python bench.py --dtype f32 --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_c2_f32.json 2>/dev/null
for f in gpurun_out/r2r_bench_c4_f32.json gpurun_out/r2r_bench_c2_f32.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and "%.4e"%d["e2e"]["value"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
