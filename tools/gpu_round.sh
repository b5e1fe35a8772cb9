#!/bin/bash
# One GPU-box visit: parity tests, the headline bench line, and an ncu launch list of a short run.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [bench args...]
tag=${1:-r01}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/${tag}_tests.log
grep -E "passed|failed|EXIT" gpurun_out/${tag}_tests.log | tail -5
timeout 1200 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --points 600000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/${tag}_launches.csv
