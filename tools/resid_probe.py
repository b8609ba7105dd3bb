"""Diagnostic (GPU box): global residual of one bench step (sle.als, repeats given) for the SKTT_DEBUG variant in force."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload_cores
from scikit_tt_b200 import TT, tensor_train as ttm
from scikit_tt_b200.solvers import sle
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 1
opc, rhsc, x0c = workload_cores(32, 64, 64)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
t = time.perf_counter()
sol = sle.als(op, x0, rhs, repeats=rep)
torch.cuda.synchronize()
dt = time.perf_counter() - t
res = float(ttm.residual_error(op, sol, rhs) / np.prod([np.linalg.norm(c) for c in rhs.cores]))
print(json.dumps(dict(debug=os.environ.get("SKTT_DEBUG", "0"), repeats=rep, seconds=dt, residual=res)))
