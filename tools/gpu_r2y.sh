#!/bin/bash
# N-GPU C4 line of the final build the way the driver launches it
G=${1:-8}
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/r2y_bench_c4_${G}gpu.json 2> gpurun_out/r2y_bench_c4_${G}gpu.err
python - gpurun_out/r2y_bench_c4_${G}gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4e ms %.1f e2e %.4e"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["correctness"]["ok"], d["correctness"]["elbo_rel"], max(d["correctness"]["grad_rel_to_max"].values()), {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()}, d["clocks"])
PY
grep -c "NCCL INFO" gpurun_out/r2y_bench_c4_${G}gpu.err; grep -E "nranks" gpurun_out/r2y_bench_c4_${G}gpu.err | head -1 | cut -c1-200
grep -E "NCCL INFO.*(nranks|Init COMPLETE|NVLS|Connected all)" gpurun_out/r2y_bench_c4_${G}gpu.err | head -20 | cut -c1-220 > gpurun_out/r2y_nccl_excerpt_${G}gpu.txt
