"""Diagnostic (GPU box): where one iteration of the persistent CG kernel goes -- %globaltimer stamps of CTA 0 during a real
solve at the bench shape (ctx debug bit 0): [matvec start, stage 1 done, grid barrier, stage 2+3 done] per matvec and
[update done, reduction done] per iteration."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
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
yt = torch.zeros_like(vt)
for dbg, name in ((0, "plain"), (256, "with the fused dot + grid sum")):
    dev.set_debug(dbg)
    for _ in range(2):
        dev.local_matvec_tiled_repeat(op, vt, yt, 200)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.local_matvec_tiled_repeat(op, vt, yt, 200); e1.record(); torch.cuda.synchronize()
    st_ = dev.scratch_peek(65536 + (4 * 256 + 8) * 8, 40)
    kk = int(st_[0]); tt = np.array(st_[1:1 + kk], dtype=np.float64)
    print(f"matvec loop {name}: {e0.elapsed_time(e1) * 1e3 / 200:.2f} us per matvec; first stamps", [round(float(x), 2) for x in np.diff(tt)[:12] / 1e3])
dev.set_debug(0)
for rep in range(3):
    u = torch.zeros(f.size, dtype=torch.float64, device="cuda")
    dev.set_debug(1 if rep == 2 else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st, iters, relres, cycles = dev.krylov_solve_refined(op, df, u, tol=1e-14, max_iters=4000, max_cycles=5)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(status=st, iters=iters, relres=relres, cycles=cycles, ms=e0.elapsed_time(e1),
                          us_per_iteration=1e3 * e0.elapsed_time(e1) / max(iters, 1))))
st_ = dev.scratch_peek(3600, 31)
dev.set_debug(0)
k = int(st_[0]); t = np.array(st_[1:1 + k], dtype=np.float64)
d = np.diff(t) / 1e3
print("stamps:", k)
print("initial true residual: stage 1 %.2f, barrier %.2f, stage 2+3 %.2f" % tuple(d[0:3]))
i = 4
print("then per iteration [gap before matvec, stage 1, barrier, stage 2+3, reduce(delta) + update, reduce(rr)]:")
while i + 5 < k:
    print("   ", " ".join(f"{x:6.2f}" for x in d[i - 1:i + 5]), "| total %.2f" % float(sum(d[i - 1:i + 5])))
    i += 6
