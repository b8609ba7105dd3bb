"""Diagnostic (GPU box): where the end-to-end step spends its time outside the sweep (upload, download, TT construction)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle
opc, rhsc, x0c = workload_cores(32, 64, 64)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out
ms_state, st = t(lambda: sle._State(op, x0, rhs))
ms_run, _ = t(lambda: (st.reset(list(sle._State(op, x0, rhs).x)), sle._run_als(st, 1, 'solve')), n=3)
ms_res, sol = t(lambda: st.result())
ms_all, _ = t(lambda: sle.als(op, x0, rhs, repeats=1), n=3)
print(f"upload (_State) {ms_state:.1f} ms, upload+sweep {ms_run:.1f} ms, download (result) {ms_res:.1f} ms, sle.als total {ms_all:.1f} ms")
x0p = TT([np.array(c) for c in x0.cores])            # pageable copy of the same guess
ms_pageable, _ = t(lambda: sle.als(op, x0p, rhs, repeats=1), n=3)
print(f"pinned-guess sle.als {ms_all:.1f} ms, pageable-guess sle.als {ms_pageable:.1f} ms; guess core pinned: "
      f"{torch.from_numpy(x0.cores[5].reshape(-1)).is_pinned()}, result core pinned: {torch.from_numpy(sol.cores[5].reshape(-1)).is_pinned()}")
