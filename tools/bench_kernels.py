"""Kernel-level timing probe (GPU box): the three contractions of one C3/C4 stack update / matvec in
each GEMM mode, plus the dense factorisations.  Prints one line per measurement; not the bench."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scikit_tt_b200._device import get_device

dev = get_device()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def stack_flops(r, R, n, r2, R2):
    return 2 * r * R * r * n * r2 + 2 * r * r2 * R * R2 * n * n + 2 * r2 * R2 * r * n * r2


out = []
for name, (r, R, n) in {"C1": (4, 4, 64), "C2r32": (32, 21, 3), "C3": (64, 3, 64), "C3r32": (32, 3, 64),
                        "C4r128": (128, 8, 16), "C4r256": (256, 8, 16)}.items():
    g = torch.Generator(device="cuda").manual_seed(0)
    L = torch.randn(r, R, r, dtype=torch.float64, device="cuda", generator=g)
    Rt = torch.randn(r, R, r, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(r, n, r, dtype=torch.float64, device="cuda", generator=g)
    A = torch.randn(R, n, n, R, dtype=torch.float64, device="cuda", generator=g)
    F = stack_flops(r, R, n, r, R)
    for mode in (0, 1, 2):
        dev.set_gemm_mode(mode)
        l0 = dev.launches()
        dev.stack_left_op(L, x, A)
        nl = dev.launches() - l0
        t1 = timeit(lambda: dev.stack_left_op(L, x, A))
        t2 = timeit(lambda: dev.stack_right_op(Rt, x, A))
        t3 = timeit(lambda: dev.micro_matvec_als(L, A, Rt, x))
        rec = dict(cfg=name, mode=mode, launches=nl, left_us=round(t1, 2), right_us=round(t2, 2), matvec_us=round(t3, 2),
                   gflop=F / 1e9, left_tflops=round(F / t1 / 1e6, 3), right_tflops=round(F / t2 / 1e6, 3),
                   matvec_tflops=round(F / t3 / 1e6, 3))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    dev.set_gemm_mode(0)

# plain GEMM sweep (DMMA vs SIMT) for tuning
big = 1 << 40
for (M, N, K) in [(192, 4096, 64), (4096, 192, 192), (4096, 64, 192), (2048, 2048, 2048), (4096, 4096, 32), (4096, 4096, 256)]:
    a = torch.randn(M, K, dtype=torch.float64, device="cuda")
    b = torch.randn(K, N, dtype=torch.float64, device="cuda")
    c = torch.empty(M, N, dtype=torch.float64, device="cuda")
    for mode in (1, 2):
        dev.set_gemm_mode(mode)
        t = timeit(lambda: dev.gemm2(M, N, K, a, (big, 0, K), (big, 0, 1), b, (big, 0, N), (big, 0, 1), c, (big, 0, N), (big, 0, 1)))
        print(json.dumps(dict(gemm=[M, N, K], mode=mode, us=round(t, 2), tflops=round(2 * M * N * K / t / 1e6, 3))), flush=True)
    dev.set_gemm_mode(0)
    t = timeit(lambda: torch.matmul(a, b, out=c))
    print(json.dumps(dict(gemm=[M, N, K], mode="cublas", us=round(t, 2), tflops=round(2 * M * N * K / t / 1e6, 3))), flush=True)

# dense factorisations
for N in (192, 1024, 3072):
    for dt in (torch.float64, torch.complex128):
        M = torch.randn(N, N, dtype=dt, device="cuda") + N ** 0.5 * torch.eye(N, dtype=dt, device="cuda")
        f = torch.randn(N, dtype=dt, device="cuda")
        def run():
            m = M.clone()
            ipiv, info = dev.lu_factor(m)
            dev.lu_solve(m, ipiv, f.clone())
        t = timeit(run, iters=3, warm=1)
        print(json.dumps(dict(lu_solve=N, dtype=str(dt), us=round(t, 1))), flush=True)
for (m, n) in [(256, 4), (4096, 64), (4096, 256)]:
    A = torch.randn(m, n, dtype=torch.float64, device="cuda")
    t = timeit(lambda: dev.qr(A), iters=5, warm=1)
    print(json.dumps(dict(qr=[m, n], us=round(t, 1))), flush=True)
    t = timeit(lambda: dev.svd(A), iters=3, warm=1)
    print(json.dumps(dict(svd=[m, n], us=round(t, 1), sweeps=dev.last_svd_sweeps)), flush=True)
