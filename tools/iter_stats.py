"""Diagnostic (GPU box): CG iterations / cycles per micro solve of one bench step (SKTT_TRACE=1 makes the solves synchronous)."""
import os, sys, re, io, contextlib
os.environ["SKTT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle
opc, rhsc, x0c = workload_cores(32, 64, 64)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    sle.als(op, x0, rhs, repeats=reps)
its = [(int(m.group(2)), int(m.group(3)), float(m.group(1))) for m in
       re.finditer(r"true relres ([0-9.e+-]+) after (\d+) iterations, (\d+) cycles", buf.getvalue())]
per_half = len(its) // (2 * reps) + 1
print("solves", len(its), "iterations", sum(i for i, _, _ in its), "cycles", sum(c for _, c, _ in its))
for h in range(0, len(its), 32):
    chunk = its[h:h + 32]
    print("solves %3d..%3d: iterations %s" % (h, h + len(chunk) - 1, [i for i, _, _ in chunk]))
print("max relres", max(r for _, _, r in its))
