"""GPU probe: tall QR at the C4 shapes (4096 x 128 / 256: beyond the 64 columns the sketched CholeskyQR kernel takes)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
for (m, n) in ((4096, 64), (4096, 128), (4096, 256), (8192, 256)):
    A = torch.randn(m, n, dtype=torch.float64, device="cuda")
    for _ in range(2): q = dev.qr(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = dev.launches()
    e0.record()
    for _ in range(5): q = dev.qr(A)
    e1.record(); torch.cuda.synchronize()
    qh = q.cpu().numpy()
    print(json.dumps(dict(m=m, n=n, us=e0.elapsed_time(e1) / 5 * 1e3, launches=(dev.launches() - l0) / 5,
                          orth=float(np.linalg.norm(qh.T @ qh - np.eye(n))))))
