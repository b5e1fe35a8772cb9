#!/bin/bash
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
./build/tf32x3_test timing > gpurun_out/r2e_tf32x3.jsonl 2>&1; echo "exit $?" >> gpurun_out/r2e_tf32x3.jsonl; cat gpurun_out/r2e_tf32x3.jsonl
AGP_B200_LIB=$PWD/build/libagp_timing.so python tools/c1_phases.py 2>&1 | tail -4 | tee gpurun_out/r2e_c1_phases.txt
timeout 600 python -m pytest tests/test_gpu_stepper.py -m gpu -q 2>&1 | tail -5
python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:{a:round(b,1) for a,b in v.items() if a.startswith('us_')} for k,v in d['us_per_evaluation'].items()})"
