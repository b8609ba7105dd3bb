"""Diagnostic (GPU box): persistent matvec vs the two-kernel matvec (values and time), then one refined solve both ways."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
from oracle import kernels as K
dev = get_device()
rng = np.random.default_rng(0)
r = n = 64
S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1); D = np.sqrt(1e-3) * 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1)); I = np.eye(n)
A = np.zeros((3, n, n, 3)); A[0, :, :, 0], A[1, :, :, 0], A[2, :, :, 0], A[2, :, :, 1], A[2, :, :, 2] = I, D, S, D, I
def spd(lo, hi):
    q, _ = np.linalg.qr(rng.standard_normal((r, r))); return (q * np.geomspace(lo, hi, r)) @ q.T
L = np.stack([spd(1.0, 3.0), np.zeros((r, r)), np.eye(r)], axis=1)
Rt = np.stack([np.eye(r), np.zeros((r, r)), spd(1.0, 3.0)], axis=1)
f = rng.standard_normal((r, n, r))
dL, dA, dR, df = (dev.to_device(x) for x in (L, A, Rt, f))
op = dev.local_op(dL, dA, dR, prepare=True)
nt = dev.tiled_len(op)
vt = torch.zeros(nt, dtype=torch.float64, device="cuda"); vt.view(n, r, 68)[:, :, :64] = df.permute(1, 0, 2)
y1 = dev.local_matvec_tiled(op, vt).clone()
y2 = torch.zeros_like(y1); dev.local_matvec_tiled_repeat(op, vt, y2, 1); torch.cuda.synchronize()
print("persistent vs two-kernel matvec: max abs diff", float((y1 - y2).abs().max()), "norm", float(y1.norm()))
want = K.micro_matvec_als(L, A, Rt, f)
got = y2.view(n, r, 68)[:, :, :64].permute(1, 0, 2).cpu().numpy()
print("vs oracle rel err", np.linalg.norm(got - want) / np.linalg.norm(want))
for reps in (1, 10, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.local_matvec_tiled_repeat(op, vt, y2, reps); torch.cuda.synchronize()
    e0.record(); dev.local_matvec_tiled_repeat(op, vt, y2, reps); e1.record(); torch.cuda.synchronize()
    print("reps", reps, "us per matvec", e0.elapsed_time(e1) * 1e3 / reps)
dev.local_matvec_tiled_repeat(op, vt, y2, 3)
st_ = dev.scratch_peek(65536 + (4 * 256 + 8) * 8, 40)
k = int(st_[0]); tt = np.array(st_[1:1 + k], dtype=np.float64)
print("phase stamps (us) [start, S1 done, barrier, S23 done | ...]:", [round(float(x), 1) for x in np.diff(tt) / 1e3])
for dbg in (16, 0):
    dev.set_debug(dbg)
    u = torch.zeros(f.size, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize(); t = time.perf_counter()
    st, iters, relres, cycles = dev.krylov_solve_refined(op, df, u, tol=1e-14, max_iters=4000, max_cycles=5)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    uh = u.cpu().numpy().reshape(f.shape)
    true = np.linalg.norm(f - K.micro_matvec_als(L, A, Rt, uh)) / np.linalg.norm(f)
    print(json.dumps(dict(debug=dbg, status=st, iters=iters, relres=relres, cycles=cycles, ms=dt * 1e3, true_relres=true)))
