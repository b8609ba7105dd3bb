"""Diagnostics: find the C5 micro matrices on which the shift-invert Arnoldi does not reach its tolerance, dump them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import workloads
from scikit_tt_b200 import TT
import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200.solvers import evp
from scikit_tt_b200._device import get_device
dev = get_device()
d = 20
ks = workloads.c5_pressures(64)
guess = None
orig = dev.eig_shift_invert
dumped = [0]
def spy(M, sigma, k, **kw):
    keep = M.clone()
    try:
        return orig(M, sigma, k, **kw)
    except np.linalg.LinAlgError as e:
        if dumped[0] < 3:
            np.save(f"gpurun_out/badM_{dumped[0]}.npy", keep.cpu().numpy())
            dumped[0] += 1
            print("dumped", keep.shape, str(e), flush=True)
        raise
dev.eig_shift_invert = spy
for j, k in enumerate(ks):
    t = TT(workloads.co_oxidation_cores(d, k)).ortho_left().ortho_right()
    op = tt.eye(t.row_dims) + t
    if guess is None:
        guess = tt.ones(op.row_dims, [1] * d, ranks=8).ortho_left().ortho_right()
    try:
        lam, x, it = evp.als(op, guess, repeats=2, conv_eps=0, solver='eigs')
        print(j, "ok", lam, flush=True)
    except np.linalg.LinAlgError as e:
        print(j, "FAILED", e, flush=True)
    if dumped[0] >= 3:
        break
