"""ncu target (GPU box): a few persistent-kernel launches of 20 matvecs each at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
rng = np.random.default_rng(0)
r = n = 64
S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1); D = np.sqrt(1e-3) * 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1)); I = np.eye(n)
A = np.zeros((3, n, n, 3)); A[0, :, :, 0], A[1, :, :, 0], A[2, :, :, 0], A[2, :, :, 1], A[2, :, :, 2] = I, D, S, D, I
L, Rt = rng.standard_normal((r, 3, r)), rng.standard_normal((r, 3, r))
dL, dA, dR = (dev.to_device(x) for x in (L, A, Rt))
op = dev.local_op(dL, dA, dR, prepare=True)
nt = dev.tiled_len(op)
vt = torch.randn(nt, dtype=torch.float64, device="cuda"); vt.view(n, r, 68)[:, :, 64:] = 0
y = torch.zeros_like(vt)
for _ in range(4):
    dev.local_matvec_tiled_repeat(op, vt, y, 20)
torch.cuda.synchronize()
