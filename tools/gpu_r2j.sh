#!/bin/bash
# N-GPU call (N = $1): C5 full sweep (N = 1e8, D = 16, M = 2048; strong scaling) and its minibatch mode (weak scaling) the way the driver launches bench.py
G=${1:-8}
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
run() { # workload, steps, warmup, extra
  if [ "$G" = "1" ]; then timeout 1200 python bench.py --gpus 1 --workload $1 --steps $2 --warmup $3 $4
  else timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --workload $1 --steps $2 --warmup $3 $4; fi
}
run c5 1 1 "$2" > gpurun_out/r2_bench_c5_${G}gpu.json 2> gpurun_out/r2_bench_c5_${G}gpu.err
run c5mb 3 3 "$2" > gpurun_out/r2_bench_c5mb_${G}gpu.json 2> gpurun_out/r2_bench_c5mb_${G}gpu.err
for f in c5 c5mb; do python - gpurun_out/r2_bench_${f}_${G}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4e ms %.1f"%(d["value"], d["ms_per_step"]), "e2e", d.get("e2e",{}) and d["e2e"].get("value"), d.get("correctness",{}) and d["correctness"].get("ok"), d["roofline"]["frac"], d["roofline"]["whole_step"]["frac_of_fp64_peak_executed"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -n 3 gpurun_out/r2_bench_c5_${G}gpu.err | cut -c1-300
