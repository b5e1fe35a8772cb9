#!/bin/bash
# Timing experiments on the Kuf-forward TRSM (S1): builds variant libraries with parts of the generator disabled
# (results are wrong on purpose) and prints the S1 time of a short run for each.  Development aid.
set -e
mkdir -p build gpurun_out
SRC=approximategps.jl_b200/csrc/agp.cu
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC"
for v in BASE NOGEN NOEXP WAIT2; do
  ( [ -f build/libagp_$v.so ] || nvcc $FLAGS $([ $v = BASE ] || echo -DAGP_EXP_$v) -o build/libagp_$v.so $SRC -ldl ) &
done
wait
if [ "$1" = "run" ]; then
  for v in BASE NOGEN NOEXP WAIT2; do
    AGP_B200_LIB=$PWD/build/libagp_$v.so python bench.py --points 3000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']
print('$v', 'S1 ms/step', round(k['trsm_kuf_fwd']['ms_per_step'],1), 'S5', round(k['trsm_bwd']['ms_per_step'],1), 'step', round(d['ms_per_step'],1))"
  done
fi
