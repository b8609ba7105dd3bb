"""Generates tests/golden/c2_cooxidation20.npz from the LIVE reference (run in the build container only):
BASELINE config 2 inputs -- `models.co_oxidation(20, 1e4).ortho_left().ortho_right()` plus identity, and the rank-8 guess
`tt.ones(..., ranks=8).ortho_left().ortho_right()` (examples/co_oxidation.py:77-101) -- together with the
reference's own result of `evp.als(..., repeats=2, conv_eps=0, solver='eig')` on them.
    OPENBLAS_NUM_THREADS=1 PYTHONPATH=/root/reference python tests/golden/make_c2_fixture.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import scikit_tt.models as mdl                      # noqa: E402
import scikit_tt.tensor_train as tt                 # noqa: E402
import scikit_tt.solvers.evp as evp                 # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
op = mdl.co_oxidation(20, 1e4).ortho_left().ortho_right()
full = tt.eye(op.row_dims) + op
guess = tt.ones(op.row_dims, [1] * op.order, ranks=8).ortho_left().ortho_right()
lam, x, it = evp.als(full, guess, repeats=2, conv_eps=0, solver='eig')
out = {"d": op.order, "lam": lam, "it": it}
for name, t in (("op", full), ("x0", guess), ("x", x)):
    out[name + "/n"] = t.order
    for i, c in enumerate(t.cores):
        out[f"{name}/{i}"] = c
np.savez_compressed(os.path.join(HERE, "c2_cooxidation20.npz"), **out)
print("ranks", full.ranks, "lambda", lam, "iterations", it)
