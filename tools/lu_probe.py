"""Diagnostic (GPU box): where a dense micro solve of N unknowns spends its time -- assembly, LU factor, triangular solves --
and the launch list of one factorisation."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
rng = np.random.default_rng(0)
for N in (1024, 192, 3072):
    A = rng.standard_normal((N, N)) + N ** 0.5 * np.eye(N)
    dA = dev.to_device(A); f = dev.to_device(rng.standard_normal(N))
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    t_clone = timed(lambda: dA.clone())
    l0 = dev.launches()
    t_fac = timed(lambda: dev.lu_factor(dA.clone())) - t_clone
    nl = (dev.launches() - l0) // 6
    m = dA.clone(); ipiv, info = dev.lu_factor(m)
    t_sol = timed(lambda: dev.lu_solve(m, ipiv, f.clone()))
    x = dev.lu_solve(m, ipiv, f.clone()).cpu().numpy()
    err = np.linalg.norm(A @ x - f.cpu().numpy()) / np.linalg.norm(f.cpu().numpy())
    print(json.dumps(dict(N=N, factor_us=round(t_fac, 1), solve_us=round(t_sol, 1), launches_factor=nl, relres=err)), flush=True)
