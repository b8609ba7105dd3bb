"""ncu target (GPU box): two LU factorisations of a 1024 x 1024 matrix (the second one is the one to read)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = np.random.default_rng(0).standard_normal((N, N)) + N ** 0.5 * np.eye(N)
dA = dev.to_device(A)
for _ in range(2):
    dev.lu_factor(dA.clone())
torch.cuda.synchronize()
