import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from util import load, cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import evp, _local
from scikit_tt_b200._device import get_device
dev = get_device()
z = load("evp_laplace")
opt, x0 = TT(cores(z, "op")), TT(cores(z, "x0"))
orig = _local.eigh_matrix_free
step = [0]
def spy(dev_, matvec, shape, dtype, k, v0=None, **kw):
    theta, vec = orig(dev_, matvec, shape, dtype, k, v0=v0, **kw)
    N = int(np.prod(shape))
    # dense matrix column by column from the matvec
    I = torch.eye(N, dtype=dtype, device=dev_.device)
    M = torch.stack([matvec(I[:, j].reshape(shape).contiguous()).reshape(-1) for j in range(N)], dim=1).cpu().numpy()
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    v = vec[:, 0].cpu().numpy()
    ref = V[:, -1]
    s = np.sign(np.dot(v, ref))
    print(step[0], N, "lam diff", float(theta[0]) - w[-1], "vec diff", np.linalg.norm(v - s * ref), "asym", np.abs(M - M.T).max(),
          "res", np.linalg.norm(M @ v - float(theta[0]) * v), "stats", dict(_local.lanczos_stats), flush=True)
    step[0] += 1
    return theta, vec
_local.eigh_matrix_free = spy
_local.EIGH_DENSE_LIMIT = 8
lam, x, it = evp.als(opt, x0, repeats=2, conv_eps=0, solver='eigh')
print(lam, float(z["eigh/lam"]))
