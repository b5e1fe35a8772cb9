#!/bin/bash
# S1 split (kuf_gen_kernel + in-place forward solve with column sums) vs the fused round-1 kernel: GPU tests on the new default, then A/B on one box
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2k_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2k_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2k_tests.log | tail -15
cp gpurun_out/parity_errors.json gpurun_out/r2k_parity_errors.json 2>/dev/null
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2k_bench_split.json 2> gpurun_out/r2k_bench_split.err
AGP_S1_FUSED=1 $B > gpurun_out/r2k_bench_fused.json 2>/dev/null
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c2_split.json 2>/dev/null
AGP_S1_FUSED=1 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c2_fused.json 2>/dev/null
for f in gpurun_out/r2k_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms/step=%.1f"%d["ms_per_step"], "frac=%.3f"%d["roofline"]["frac"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
