#!/bin/bash
# Float32 mode with S5 on the INT8 tensor path: tests, A/B against the FP64 S5, C4 / C2 lines
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f32.py tests/test_gpu_engines.py -m gpu -q -s > gpurun_out/r2v_f32tests.log 2>&1; grep -E "^\[f32|passed|failed|Error|error" gpurun_out/r2v_f32tests.log | cut -c1-330 | head -14
python bench.py --dtype f32 --steps 3 --warmup 3 > gpurun_out/r2v_bench_c4_f32.json 2> gpurun_out/r2v_bench_c4_f32.err
AGP_F32_S1=fp64 python bench.py --dtype f32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_c4_f32_s1fp64.json 2>/dev/null
python bench.py --dtype f32 --workload c2 --steps 5 --warmup 3 > gpurun_out/r2v_bench_c2_f32.json 2>/dev/null
for f in gpurun_out/r2v_bench_c4_f32.json gpurun_out/r2v_bench_c4_f32_s1fp64.json gpurun_out/r2v_bench_c2_f32.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get("correctness") or {}
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and "%.4e"%d["e2e"]["value"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, c.get("elbo_rel"), c.get("grad_rel_to_max"), c.get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
