#!/bin/bash
# ncu --set full captures of the sweep kernels on a short run (2 chunks of 151 552 points per step).  Usage: bash tools/ncu_full.sh <tag> <regex> [<regex> ...]
tag=$1; shift
mkdir -p gpurun_out
for rx in "$@"; do
  name=$(echo "$rx" | tr -c 'A-Za-z0-9' '_')
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$rx" -s 3 -c 2 -f -o gpurun_out/${tag}_${name} \
    python bench.py --points 303104 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${name}.log 2>&1
  echo "$rx rc=$?"
done
ls -la gpurun_out/*.ncu-rep
