"""Golden vectors for evp.power_method (scikit_tt/solvers/evp.py:182-250) from the LIVE reference:
inverse power iteration on a small Laplacian-type operator, plain and generalised.  Build container only:
    OPENBLAS_NUM_THREADS=1 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_power_golden.py"""
import os, sys
import numpy as np
import scikit_tt.tensor_train as tt
from scikit_tt.tensor_train import TT
import scikit_tt.solvers.evp as evp
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import workloads
d, n, r = 4, 6, 3
op = TT(workloads.laplace_cores(d, n, c=0.05))
x0 = TT(workloads.random_guess(d, n, r, seed=7)).ortho_right()
gev = tt.eye(op.row_dims) + 0.1 * TT(workloads.laplace_cores(d, n, c=0.0))
out = {}
def pack(prefix, t):
    out[prefix + "/n"] = np.array(len(t.cores))
    for i, c in enumerate(t.cores):
        out[f"{prefix}/{i}"] = np.asarray(c)
pack("x0", x0); pack("gev", gev)
for tag, kw in (("plain", {}), ("gevp", {"operator_gevp": gev})):
    for reps in (1, 3):
        lam, x = evp.power_method(op, x0, repeats=reps, sigma=0.3, **kw)
        out[f"{tag}/rep{reps}/lam"] = np.array(lam)
        pack(f"{tag}/rep{reps}/x", x)
        print(tag, reps, lam)
np.savez_compressed(os.path.join(HERE, "power_method.npz"), **out)
