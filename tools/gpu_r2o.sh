#!/bin/bash
# diagonal-block kernel v2 (LDL^T-style register Cholesky + DMMA trailing update): phase timing, the Adam end-to-end test on both builds, tests, C3
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 60 ./build/pt_timing > gpurun_out/r2o_pt_timing.txt 2>&1; cat gpurun_out/r2o_pt_timing.txt
timeout 600 python -m pytest tests/test_gpu_svgp.py -m gpu -q -s -k "optimised_posterior" 2>&1 | grep -E "adam|passed|failed|assert" | head -5
AGP_B200_LIB=$PWD/build/libagp_b32.so timeout 600 python -m pytest tests/test_gpu_svgp.py -m gpu -q -s -k "optimised_posterior" 2>&1 | grep -E "adam|passed|failed|assert" | head -5
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2o_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2o_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2o_tests.log | tail -15
cp gpurun_out/parity_errors.json gpurun_out/r2o_parity_errors.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_c3.json 2> gpurun_out/r2o_bench_c3.err
AGP_B200_LIB=$PWD/build/libagp_b32.so python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_c3_b32.json 2>/dev/null
for f in r2o_bench_c3 r2o_bench_c3_b32; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["roofline"]["ms_per_newton_iteration"], d["roofline"]["achieved"], d["lml"])
except Exception as e:
    print("c3 FAILED", e)
PY
done
