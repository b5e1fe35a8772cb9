#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
./build/tf32x3_test timing > gpurun_out/r2h_tf32x3.jsonl 2>&1; echo "exit $?" >> gpurun_out/r2h_tf32x3.jsonl; cat gpurun_out/r2h_tf32x3.jsonl
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r2h_f32tests.log 2>&1; grep -E "^\[f32|passed|failed|Error|error" gpurun_out/r2h_f32tests.log | head -30
AGP_F32_S5=tf32 timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s 2>&1 | grep -E "^\[f32|passed|failed" | head
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-e2e"
$B --dtype f32 > gpurun_out/r2h_bench_f32.json 2> gpurun_out/r2h_bench_f32.err; tail -2 gpurun_out/r2h_bench_f32.err
AGP_F32_S5=tf32 $B --dtype f32 > gpurun_out/r2h_bench_f32_s5t.json 2>/dev/null
for f in gpurun_out/r2h_bench_f32.json gpurun_out/r2h_bench_f32_s5t.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "ms/step=%.1f"%d["ms_per_step"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d["correctness"]["elbo_rel"], d["correctness"]["grad_rel_to_max"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
