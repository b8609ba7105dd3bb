"""Diagnostics (GPU box): full-depth bench problem (d=32) at ranks 8..64: residual after 1, 2 sweeps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, _local
from oracle import sle as osle, tt as ott

for r in (2, 8, 16, 64):
    opc, rhsc, x0c = workload_cores(32, 64, r)
    x0 = TT([c.copy() for c in x0c]).ortho_right()
    print(f"=== r={r} x0 norm {x0.norm():.3e} rhs norm {TT(rhsc).norm():.3e} res(x0) {osle.residual(opc, x0.cores, rhsc):.3e}", flush=True)
    for reps in (1, 2):
        _local._TRACE = (r == 16 and reps == 1)
        t = time.time()
        sol = sle.als(TT(opc), x0, TT(rhsc), repeats=reps)
        torch.cuda.synchronize()
        dt = time.time() - t
        res = osle.residual(opc, sol.cores, rhsc)
        print(f"r={r} repeats={reps} time {dt:.2f}s residual {res:.4e} ranks {sol.ranks[:4]}.. norm {sol.norm():.3e}", flush=True)
    if r == 2:
        ref = osle.als(opc, x0.cores, rhsc, repeats=1)
        print("oracle r=2 residual", osle.residual(opc, ref, rhsc), "rel diff", ott.norm(ott.sub(sol.cores, osle.als(opc, x0.cores, rhsc, repeats=2))) / ott.norm(ref), flush=True)
