"""cuBLAS DGEMM yard-stick (torch.float64 matmul): the FP64 roofline denominator.
Measuring stick only; nothing in the product path uses torch or cuBLAS."""
import json, torch, time
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(2): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
    if n == 8192:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = max(3, int(3000 / best))
        e0.record()
        for _ in range(reps): torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        res["dgemm_8192_sustained_tflops"] = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
