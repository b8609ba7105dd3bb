import sys, os
sys.path.insert(0, '/root/repo')
import torch
from scikit_tt_b200._device import get_device
dev = get_device()
r, R, n = 256, 8, 16
L = torch.randn((r, R, r), dtype=torch.float64, device="cuda")
x = torch.randn((r, n, r), dtype=torch.float64, device="cuda")
A = torch.randn((R, n, n, R), dtype=torch.float64, device="cuda")
for _ in range(3):
    dev.stack_left_op(L, x, A)
torch.cuda.synchronize()
