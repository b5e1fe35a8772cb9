#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3j_bench_c3_$label.json 2>gpurun_out/r3j_bench_c3_$label.err
  python - $label <<'PY'
import json,sys
s=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r3j_bench_c3_{s}.json').read().strip().splitlines()[-1])
    print(s, "it/s %.2f"%d["value"], "ms/newton %.3f"%d["roofline"]["ms_per_newton_iteration"], "lml %r"%d["lml"], "lml+grad ms %.1f"%d["lml_and_gradient_ms"], "TF %.2f"%d["roofline"]["achieved"])
except Exception as e:
    print(s, "FAILED", e); print(open(f'gpurun_out/r3j_bench_c3_{s}.err').read()[-1500:])
PY
}
run part
run nopart AGP_CHOL_PARTITION=0
AGP_CHOL_TRACE=1 timeout 300 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/r3j_chol_trace_v2.txt >/dev/null
grep "chol partition" gpurun_out/r3j_chol_trace_v2.txt | head -2
grep "chol trace" gpurun_out/r3j_chol_trace_v2.txt | awk '{print $4, $7, $10, $13}' | tr -d 'J=,' | awk '{printf "%s:%s/%s/%s  ", $1,$2,$3,$4; if (NR%4==0) print ""}'
timeout 600 python -m pytest tests/test_gpu_laplace.py -m gpu -q -x 2>&1 | tail -2
