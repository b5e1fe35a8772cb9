#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --maxfail=10 > gpurun_out/r2c_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2c_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_tests.log | tail -15
python bench.py --workload c1 --steps 3 --warmup 3 > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_c1.json')); print(d['us_per_evaluation']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/r2c_bench_c1.err
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline'])"
