import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops_burst"] = 2 * n**3 / best * 1e-9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 20 if n == 8192 else 100
    e0.record()
    for _ in range(reps): c = a @ b
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained"] = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
# batched small-ish gemm like the stack contraction shapes
for (m, k, n) in ((192, 64, 4096), (4096, 192, 192), (192, 4096, 64)):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    for _ in range(5): c = a @ b
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): c = a @ b
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 200
    res[f"cublas_{m}x{k}x{n}_us"] = t * 1e3
    res[f"cublas_{m}x{k}x{n}_tflops"] = 2 * m * n * k / t * 1e-9
print(json.dumps(res, indent=1))
