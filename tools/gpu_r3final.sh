#!/bin/bash
# final build of round 2: GPU suite, smoke, headline line (default Float64 mode) and the opt-in emulation mode's line
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r3final_tests.log 2>&1; echo "EXIT $?" >> gpurun_out/r3final_tests.log
grep -E "passed|failed|FAILED|EXIT" gpurun_out/r3final_tests.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r3final_bench_c4_1gpu.json 2> gpurun_out/r3final_bench_c4_1gpu.err
python bench.py --dtype f64emu --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3final_bench_c4_f64emu.json 2>/dev/null
python bench.py --dtype f64emu --workload c5mb --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3final_bench_c5mb_f64emu.json 2>/dev/null
for f in c4_1gpu c4_f64emu c5mb_f64emu; do python - gpurun_out/r3final_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["dtype"], "value=%.4e"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "e2e=", d.get("e2e") and d["e2e"].get("value") and "%.4e"%d["e2e"]["value"], "frac=%.3f"%d["roofline"]["frac"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, (d.get("correctness") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
