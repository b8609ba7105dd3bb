"""Timing of BASELINE config 4 on the GPU box (not the bench line): d=10, n=16, operator rank 8 (SPD variant of SURVEY.md
8d: rank-diagonal cores, M = I + 0.1 sym(randn)), solution rank 128 / 256, one sle.als sweep.  The reference cannot run this
size (dense micro matrices of 512 GiB / 8 TiB); the check is the global residual ||A x - b|| / ||b|| after the sweep."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200 import TT
import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200.solvers import sle
from scikit_tt_b200._device import get_device
dev = get_device()
d, n, R = 10, 16, 8
rng = np.random.default_rng(2)
def sym(a): return 0.5 * (a + a.T)
blocks = [[np.eye(n) + 0.1 * sym(rng.standard_normal((n, n))) for _ in range(R)] for _ in range(d)]
opc = []
for k in range(d):
    Rl, Rr = (1 if k == 0 else R), (1 if k == d - 1 else R)
    c = np.zeros((Rl, n, n, Rr))
    for b in range(R):
        c[0 if k == 0 else b, :, :, 0 if k == d - 1 else b] = blocks[k][b]
    opc.append(c)
op = TT(opc).pin_memory()
rhs = TT([np.random.default_rng(0).standard_normal((1, n, 1, 1)) for _ in range(d)])
bnorm = np.prod([np.linalg.norm(c) for c in rhs.cores])
for r in [int(a) for a in sys.argv[1:]] or [128]:
    ranks = [1] + [r] * (d - 1) + [1]
    for i in range(1, d): ranks[i] = min(ranks[i], ranks[i - 1] * n)
    for i in range(d - 1, 0, -1): ranks[i] = min(ranks[i], ranks[i + 1] * n)
    x0 = TT([np.random.default_rng(1 + i).standard_normal((ranks[i], n, 1, ranks[i + 1])) for i in range(d)]).ortho_right()
    l0 = dev.launches()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sol = sle.als(op, x0, rhs, repeats=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res = float(tt.residual_error(op, sol, rhs) / bnorm)
    print(json.dumps(dict(cfg=f"C4 r={r}", ranks=sol.ranks, half_sweeps=2, seconds=dt, hs_per_s=2 / dt, launches=dev.launches() - l0,
                          residual=res)), flush=True)
