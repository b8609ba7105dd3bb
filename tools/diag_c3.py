"""Diagnostics (GPU box): the bench problem family at reduced sizes with micro-solve tracing, checked against the
oracle where the dense reference algorithm can still run."""
import os, sys, time
os.environ["SKTT_TRACE"] = os.environ.get("SKTT_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, _local
from oracle import sle as osle, tt as ott

def run(d, n, r, dense_limit, check_oracle):
    print(f"=== d={d} n={n} r={r} dense_limit={dense_limit}", flush=True)
    opc, rhsc, x0c = workload_cores(d, n, r)
    x0c = ott.ortho_right(x0c)
    _local.DENSE_LIMIT = dense_limit
    t = time.time()
    sol = sle.als(TT(opc), TT([c.copy() for c in x0c]), TT(rhsc), repeats=1)
    torch.cuda.synchronize()
    print("time", time.time() - t, "residual", osle.residual(opc, sol.cores, rhsc), flush=True)
    if check_oracle:
        ref = osle.als(opc, x0c, rhsc, repeats=1)
        print("oracle residual", osle.residual(opc, ref, rhsc), "rel diff", ott.norm(ott.sub(sol.cores, ref)) / ott.norm(ref), flush=True)
    # GPU ortho_right against the oracle's
    g = TT([c.copy() for c in workload_cores(d, n, r)[2]]).ortho_right()
    print("ortho_right rel diff vs oracle", ott.norm(ott.sub(g.cores, x0c)) / ott.norm(x0c), flush=True)

run(6, 16, 16, 64, True)
run(6, 16, 16, 1 << 20, True)
os.environ["SKTT_TRACE"] = "1"
run(6, 64, 32, 8192, False)
