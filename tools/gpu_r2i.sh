#!/bin/bash
# 2-GPU call: multi-GPU parity test, bench at N=2 the way the driver launches it (correctness block + NCCL log), and N=1 on the same box
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_c4_2gpu.json 2> gpurun_out/r2i_bench_c4_2gpu.err
grep -c "NCCL INFO" gpurun_out/r2i_bench_c4_2gpu.err; grep -E "nranks|Init COMPLETE" gpurun_out/r2i_bench_c4_2gpu.err | head -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_c4_2gpu.json').read().strip().splitlines()[-1])
print("2gpu value %.4e e2e %.4e ms %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["correctness"], d["roofline"]["frac"], d["roofline"]["peak"])
PY
python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_bench_c4_1gpu.json 2> gpurun_out/r2i_bench_c4_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_c4_1gpu.json').read().strip().splitlines()[-1])
print("1gpu value %.4e e2e %.4e ms %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["correctness"], d["roofline"]["frac"], d["roofline"]["whole_step"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d["cpu_baseline"])
PY
