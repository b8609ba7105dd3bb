"""Golden vectors for ode.tdvp1site / ode.tdvp2site from the LIVE reference: transverse-field Ising
chain (models.ising(5, J=1, h=1.2), tests/test_ode.py:135-158), real- and imaginary-time steps, exact and local-Krylov
micro solvers.  Build container only:
    OPENBLAS_NUM_THREADS=1 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_tdvp_golden.py"""
import os

import numpy as np

import scikit_tt.tensor_train as tt
from scikit_tt.tensor_train import TT
import scikit_tt.solvers.ode as ode
import scikit_tt.models as mdl

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


def pack(prefix, t):
    out[prefix + "/n"] = np.array(len(t.cores))
    for i, c in enumerate(t.cores):
        out[f"{prefix}/{i}"] = np.asarray(c)


N = 5
op = mdl.ising(N, J=1.0, h=1.2)
rng = np.random.default_rng(12)
ranks = [1, 2, 3, 3, 2, 1]
x0 = TT([rng.standard_normal((ranks[i], 2, 1, ranks[i + 1])) + 1j * rng.standard_normal((ranks[i], 2, 1, ranks[i + 1]))
         for i in range(N)]).ortho()
x0 = (1 / x0.norm()) * x0
pack("op", op)
pack("x0", x0)
cases = {"real_exact": (0.05, None, 0), "imag_exact": (-1j * 0.05, None, 2), "real_krylov": (0.05, {"method": "local_krylov", "dimension": 4}, 0)}
for tag, (h, solver, normalize) in cases.items():
    sol = ode.tdvp1site(op, x0, h, 3, local_solver=solver, normalize=normalize)
    for k in range(1, 4):
        pack(f"tdvp1/{tag}/step{k}", sol[k])
    sol = ode.tdvp2site(op, x0, h, 3, local_solver=solver, threshold=1e-10, max_rank=4, normalize=normalize)
    for k in range(1, 4):
        pack(f"tdvp2/{tag}/step{k}", sol[k])
    print(tag, "ranks 1site", sol[-1].ranks)
energy = lambda t: complex(t.transpose(conjugate=True) @ op @ t)
out["energy0"] = np.array(energy(x0))
np.savez_compressed(os.path.join(HERE, "tdvp.npz"), **out)
print(os.path.getsize(os.path.join(HERE, "tdvp.npz")) / 1024, "KiB")
