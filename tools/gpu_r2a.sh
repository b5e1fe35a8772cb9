#!/bin/bash
# Round-2 GPU call A: GPU test-suite on the new build, S1 variant A/B (same box, back to back), ncu capture of S1.
#   build/libagp_noexp.so = the same sources with -DAGP_NO_FAST_EXP (library exp() in the Kuf generator)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2a_bench_new.json 2> gpurun_out/r2a_bench_new.err
AGP_S1_DEPHASE=0 $B > gpurun_out/r2a_bench_nodephase.json 2>/dev/null
AGP_B200_LIB=$PWD/build/libagp_noexp.so $B > gpurun_out/r2a_bench_noexp.json 2>/dev/null
AGP_S1_DEPHASE=0 AGP_B200_LIB=$PWD/build/libagp_noexp.so $B > gpurun_out/r2a_bench_noexp_nodephase.json 2>/dev/null
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_c2_new.json 2>/dev/null
AGP_S1_DEPHASE=0 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_c2_nodephase.json 2>/dev/null
for f in gpurun_out/r2a_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms/step=%.1f"%d["ms_per_step"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:trsm_kernelILi0 -s 4 -c 1 -o gpurun_out/r2a_s1 \
  python bench.py --points 303104 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu.log 2>&1
ls -la gpurun_out/r2a_s1.ncu-rep
