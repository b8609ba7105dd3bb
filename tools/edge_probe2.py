"""Diagnostic (GPU box): the last micro system of the bench step (core 0, backward) -- persistent vs two-kernel matvec."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload_cores
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, _local
from scikit_tt_b200._device import get_device
from oracle import kernels as K
dev = get_device()
opc, rhsc, x0c = workload_cores(32, 64, 64)
op, rhs = TT(opc), TT(rhsc)
x0 = TT(x0c).ortho_right()
st = sle._State(op, x0, rhs)
grab = {}
orig = _local.solve_micro
calls = [0]
def spy(dev_, solver, dense_builder, lop, f, guess, cache=None):
    calls[0] += 1
    if calls[0] == 63:
        grab.update(lop=lop, f=f.clone(), guess=None if guess is None else guess.clone())
    return orig(dev_, solver, dense_builder, lop, f, guess, cache)
_local.solve_micro = spy
sle._run_als(st, 1, 'solve')
torch.cuda.synchronize()
lop, f, guess = grab["lop"], grab["f"], grab["guess"]
L, A1, A2, Rt = lop._keep
print("shapes", tuple(L.shape), tuple(A1.shape), tuple(Rt.shape), tuple(f.shape))
Lh, Ah, Rh = L.cpu().numpy(), A1.cpu().numpy(), Rt.cpu().numpy()
M = K.micro_matrix_als(Lh, Ah, Rh)
w = np.linalg.eigvalsh(0.5 * (M + M.T))
print("sym err", np.abs(M - M.T).max() / np.abs(M).max(), "eig min/max", w[0], w[-1])
dev.prepare_local_op(lop)
nt = dev.tiled_len(lop)
r, n = L.shape[0], A1.shape[2]
rp = (r + 3) // 4 * 4
v = torch.randn(r, n, 64, dtype=torch.float64, device="cuda")
vt = torch.zeros(n, rp, 68, dtype=torch.float64, device="cuda"); vt[:, :r, :64] = v.permute(1, 0, 2)
y1 = dev.local_matvec_tiled(lop, vt.reshape(-1)).clone()
y2 = torch.zeros_like(y1); dev.local_matvec_tiled_repeat(lop, vt.reshape(-1), y2, 1); torch.cuda.synchronize()
want = (M @ v.cpu().numpy().reshape(-1)).reshape(r, n, 64)
g1 = y1.view(n, rp, 68)[:, :r, :64].permute(1, 0, 2).cpu().numpy()
g2 = y2.view(n, rp, 68)[:, :r, :64].permute(1, 0, 2).cpu().numpy()
print("two-kernel vs dense", np.linalg.norm(g1 - want) / np.linalg.norm(want), " persistent vs dense", np.linalg.norm(g2 - want) / np.linalg.norm(want))
print("padding of persistent result: rows", float(y2.view(n, rp, 68)[:, r:, :].abs().max()), "cols", float(y2.view(n, rp, 68)[:, :, 64:].abs().max()))
for dbg in (16, 0):
    dev.set_debug(dbg)
    u = guess.reshape(-1).clone()
    st_, iters, relres, cycles = dev.krylov_solve_refined(lop, f, u, tol=1e-14, max_iters=20000, max_cycles=5)
    pk = dev.scratch_peek(65536 + 4 * 256 * 8, 4, ctype=ctypes.c_double)
    uref = np.linalg.solve(M, f.cpu().numpy().reshape(-1))
    print(json.dumps(dict(debug=dbg, iters=iters, relres=relres, cycles=cycles, err=float(np.linalg.norm(u.cpu().numpy() - uref) / np.linalg.norm(uref)), persistent_out=[float(x) for x in pk])))
