"""Timing of the dense solve routes at N = 1024 (CUDA events): one-launch fused LU vs the host-driven factorisation."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
rng = np.random.default_rng(0)
for N in (256, 1024, 1536):
    M = dev.to_device(rng.standard_normal((N, N)) + N * 0.01 * np.eye(N)); f = dev.to_device(rng.standard_normal(N))
    def t(fn, reps=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(reps):
            Mc = M.clone(); torch.cuda.synchronize()
            e0.record(); fn(Mc); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / reps * 1e3
    fused = t(lambda Mc=None: dev.solve_fused(M.clone() if Mc is None else Mc, f))
    def old(Mc=None):
        Mc = M.clone() if Mc is None else Mc
        ipiv, info = dev.lu_factor(Mc); dev.lu_solve(Mc, ipiv, f.clone())
    print(json.dumps(dict(N=N, fused_us=fused, host_driven_us=t(old))), flush=True)
