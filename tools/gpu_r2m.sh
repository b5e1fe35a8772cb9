#!/bin/bash
# SE-specialised fast kgrad (no Kf), optimised generator: tests, short A/B, full C4 line, ncu --set full of kuf_gen / kgrad / S4
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2m_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r2m_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2m_tests.log | tail -15
cp gpurun_out/parity_errors.json gpurun_out/r2m_parity_errors.json 2>/dev/null
B="python bench.py --points 3031040 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2m_bench_new.json 2> gpurun_out/r2m_bench_new.err
AGP_KGRAD_FAST=0 $B > gpurun_out/r2m_bench_kslow.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_c4_1gpu.json 2> gpurun_out/r2m_bench_c4_1gpu.err
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_c2.json 2>/dev/null
for f in gpurun_out/r2m_bench_new.json gpurun_out/r2m_bench_kslow.json gpurun_out/r2m_bench_c4_1gpu.json gpurun_out/r2m_bench_c2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], d["roofline"]["kernel"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
for rx in kuf_gen_kernel kgrad_kernel "gemm_kernel.*EpiS4"; do
  name=$(echo "$rx" | tr -c 'A-Za-z0-9' '_')
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$rx" -s 3 -c 1 -f -o gpurun_out/r2m_${name} \
    python bench.py --points 303104 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2m_ncu_${name}.log 2>&1
  echo "$rx rc=$?"
done
ls -la gpurun_out/r2m_*.ncu-rep
