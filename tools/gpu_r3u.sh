#!/bin/bash
# AGP_COMPUTE_F64_EMU: parity tests, Float32-mode tests (shared engine), C4 / C2 / C5-minibatch lines of the mode
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f64emu.py tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r3u_tests_f64emu.log 2>&1; echo "EXIT $?" >> gpurun_out/r3u_tests_f64emu.log
grep -E "passed|failed|FAILED|EXIT|Error" gpurun_out/r3u_tests_f64emu.log | tail -8
grep -E "^\[f64emu" gpurun_out/r3u_tests_f64emu.log
python bench.py --dtype f64emu --steps 5 --warmup 3 > gpurun_out/r3u_bench_c4_f64emu.json 2> gpurun_out/r3u_bench_c4_f64emu.err
python bench.py --dtype f64emu --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3u_bench_c2_f64emu.json 2>/dev/null
python bench.py --dtype f64emu --workload c5mb --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3u_bench_c5mb_f64emu.json 2>/dev/null
for f in c4 c2 c5mb; do python - gpurun_out/r3u_bench_${f}_f64emu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, json.dumps(d.get("correctness"))[:400])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
