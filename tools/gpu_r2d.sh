#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
./build/tf32x3_test > gpurun_out/r2d_tf32x3.jsonl 2>&1; echo "exit $?" >> gpurun_out/r2d_tf32x3.jsonl; cat gpurun_out/r2d_tf32x3.jsonl
AGP_B200_LIB=$PWD/build/libagp_timing.so python tools/c1_phases.py 2>&1 | tail -4 | tee gpurun_out/r2d_c1_phases.txt
