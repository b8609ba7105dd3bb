"""Kernel-level timing of the batched small-system path (CUDA events): eig / svd / stacks / micro matrix at the C5 shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
rng = np.random.default_rng(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for N in (81, 192, 768):
    lam = 1.0 + 0.05 * (np.arange(N) + 1)
    mats = []
    for b in range(B):
        Q = rng.standard_normal((N, N)) / np.sqrt(N) + np.eye(N)
        mats.append(Q @ np.diag(lam) @ np.linalg.inv(Q))
    M = dev.to_device(np.stack(mats))
    us = timed(lambda: dev.batch_eig_shift_invert(M, 1.0, 1), reps=5)
    _, _, st = dev.batch_eig_shift_invert(M, 1.0, 1)
    print(json.dumps(dict(kernel="batch_eig", B=B, N=N, us=us, nconv_min=float(st[:, 0].min()))), flush=True)
big = 1 << 40
for (r, n, r2) in ((8, 3, 8), (16, 3, 16)):
    k = 1
    vec = torch.randn((B, r * n * r2, k), dtype=torch.complex128, device="cuda")
    out = dev.empty((B, r, n, r2), torch.complex128)
    P, Q = r * n, r2 * k
    us = timed(lambda: dev.batch_svd_left(vec, P, Q, min(r2, P, Q), (big, 0, Q), (big, 0, 1), 0, out, min(r2, P, Q), 1, 0))
    print(json.dumps(dict(kernel="batch_svd_left", B=B, P=P, Q=Q, us=us)), flush=True)
    R = 21
    L = torch.randn((B, r, R, r), dtype=torch.complex128, device="cuda")
    x = torch.randn((B, r, n, r2), dtype=torch.complex128, device="cuda")
    A = torch.randn((B, R, n, n, R), dtype=torch.complex128, device="cuda")
    Rt = torch.randn((B, r2, R, r2), dtype=torch.complex128, device="cuda")
    print(json.dumps(dict(kernel="batch_stack_left", B=B, r=r, us=timed(lambda: dev.batch_stack_left_op(L, x, A, 1)))), flush=True)
    print(json.dumps(dict(kernel="batch_stack_right", B=B, r=r, us=timed(lambda: dev.batch_stack_right_op(Rt, x, A)))), flush=True)
    print(json.dumps(dict(kernel="batch_micro_matrix", B=B, r=r, us=timed(lambda: dev.batch_micro_matrix_als(L, A, Rt)))), flush=True)
# where the time of the eig kernel goes: Krylov dimension sweep at N = 192 (LU is the intercept, the slope is one Arnoldi step)
N = 192
lam = 1.0 + 0.05 * (np.arange(N) + 1)
mats = []
for b in range(B):
    Q = rng.standard_normal((N, N)) / np.sqrt(N) + np.eye(N)
    mats.append(Q @ np.diag(lam) @ np.linalg.inv(Q))
M = dev.to_device(np.stack(mats))
for ncv in (2, 6, 12, 20, 32):
    us = timed(lambda: dev.batch_eig_shift_invert(M, 1.0, 1, ncv=ncv, max_restarts=0), reps=5)
    print(json.dumps(dict(kernel="batch_eig", N=N, ncv=ncv, restarts=0, us=us)), flush=True)
for Bs in (1, 8, 32):
    us = timed(lambda: dev.batch_eig_shift_invert(M[:Bs].contiguous(), 1.0, 1, ncv=20, max_restarts=0), reps=5)
    print(json.dumps(dict(kernel="batch_eig", N=N, B=Bs, ncv=20, us=us)), flush=True)
