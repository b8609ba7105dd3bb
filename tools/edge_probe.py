"""Diagnostic (GPU box): the padded edge-core solve (r = 1, R = 1) in the persistent kernel vs the host-driven path."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
from oracle import kernels as K
dev = get_device()
rng = np.random.default_rng(3)
n, r2 = 64, 64
S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1); D = np.sqrt(1e-3) * 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1)); I = np.eye(n)
A = np.zeros((1, n, n, 3)); A[0, :, :, 0], A[0, :, :, 1], A[0, :, :, 2] = S, D, I
def spd(lo, hi):
    q, _ = np.linalg.qr(rng.standard_normal((r2, r2))); return (q * np.geomspace(lo, hi, r2)) @ q.T
x = rng.standard_normal((r2, r2)); skew = 1e-3 * (x - x.T)
L = np.ones((1, 1, 1))
Rt = np.stack([np.eye(r2), skew, spd(0.5, 120.0)], axis=1)
f = rng.standard_normal((1, n, r2))
M = K.micro_matrix_als(L, A, Rt)
uref = np.linalg.solve(M, f.reshape(-1)).reshape(f.shape)
print("cond", np.linalg.cond(M))
dL, dA, dR, df = (dev.to_device(a) for a in (L, A, Rt, f))
for guess in ("zero", "warm"):
    for dbg in (16, 0):
        dev.set_debug(dbg)
        op = dev.local_op(dL, dA, dR, prepare=True)
        u = torch.zeros(f.size, dtype=torch.float64, device="cuda") if guess == "zero" else dev.to_device((uref * (1 + 1e-3 * rng.standard_normal(uref.shape))).reshape(-1))
        st, iters, relres, cycles = dev.krylov_solve_refined(op, df, u, tol=1e-14, max_iters=20000, max_cycles=5)
        pk = dev.scratch_peek(65536 + 4 * 256 * 8, 4, ctype=ctypes.c_double)
        err = np.linalg.norm(u.cpu().numpy().reshape(f.shape) - uref) / np.linalg.norm(uref)
        print(json.dumps(dict(guess=guess, debug=dbg, status=st, iters=iters, relres=relres, cycles=cycles, err=err,
                              persistent_out=[float(v) for v in pk])))
dev.set_debug(0)
