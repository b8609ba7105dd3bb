"""Golden vectors for the time steppers that sit on top of sle.als, from the LIVE reference:
`ode.trapezoidal_rule` and `ode.adaptive_step_size` on the signaling cascade of tests/golden/euler_cascade.npz
(same operator, initial value and guess).  Build container only:
    OPENBLAS_NUM_THREADS=1 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_ode_golden.py
"""
import os

import numpy as np

import scikit_tt.tensor_train as tt
from scikit_tt.tensor_train import TT
import scikit_tt.solvers.ode as ode
import scikit_tt.models as mdl

HERE = os.path.dirname(os.path.abspath(__file__))
z = np.load(os.path.join(HERE, "euler_cascade.npz"))
d = int(z["d"])
op = mdl.signaling_cascade(d)
iv = TT([z[f"iv/{i}"] for i in range(d)])
guess = TT([z[f"guess/{i}"] for i in range(d)])
out = {"d": np.array(d)}


def pack(prefix, t):
    out[prefix + "/n"] = np.array(len(t.cores))
    for i, c in enumerate(t.cores):
        out[f"{prefix}/{i}"] = np.asarray(c)


sol = ode.trapezoidal_rule(op, iv, guess, [0.5, 1.0, 0.5], repeats=2, progress=False)
for k in range(1, 4):
    pack(f"trap/step{k}", sol[k])
for method in ("two_step_Euler", "trapezoidal_rule"):
    sol, times = ode.adaptive_step_size(op, iv, guess, 2.0, step_size_first=0.1, repeats=2, second_method=method,
                                        progress=False)
    out[f"adapt/{method}/times"] = np.array(times, dtype=float)
    for k in range(1, len(sol)):
        pack(f"adapt/{method}/step{k}", sol[k])
    print(method, "accepted steps", len(sol) - 1, "times", times)
np.savez_compressed(os.path.join(HERE, "ode_steppers.npz"), **out)
print(os.path.getsize(os.path.join(HERE, "ode_steppers.npz")) / 1024, "KiB")
