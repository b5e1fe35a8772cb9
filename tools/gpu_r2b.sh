#!/bin/bash
# Round-2 GPU call B: full GPU suite (incl. the one-launch stepper path and the C3 full-size case), then the new bench workloads.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --maxfail=10 > gpurun_out/r2b_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2b_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2b_tests.log | tail -15
python bench.py --workload c1 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c1.json 2> gpurun_out/r2b_bench_c1.err; tail -c 1500 gpurun_out/r2b_bench_c1.json; tail -3 gpurun_out/r2b_bench_c1.err
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err; cut -c1-1500 gpurun_out/r2b_bench_c3.json; tail -3 gpurun_out/r2b_bench_c3.err
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; cut -c1-600 gpurun_out/r2b_bench_c2.json; tail -3 gpurun_out/r2b_bench_c2.err
python bench.py --points 3031040 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c4short.json 2> gpurun_out/r2b_bench_c4short.err; cut -c1-600 gpurun_out/r2b_bench_c4short.json; tail -3 gpurun_out/r2b_bench_c4short.err
