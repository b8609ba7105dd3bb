"""Times the micro-matvec at a given shape (GPU box).  usage: matvec_probe.py [r R n] [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scikit_tt_b200._device import get_device
dev = get_device()
r, R, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 3, 64)
reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 200
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
L, Rt, x, A = mk(r, R, r), mk(r, R, r), mk(r, n, r), mk(R, n, n, R)
F = 2 * r * R * r * n * r + 2 * r * r * R * R * n * n + 2 * r * R * r * n * r
for mode in (0, 1) if reps > 10 else (0,):
    dev.set_gemm_mode(mode)
    op = dev.local_op(L, A, Rt, prepare=True)      # as CG / GMRES use it: images built once per local operator
    nt = dev.tiled_len(op)
    if nt > 0:                                     # the Krylov inner step: tiled vectors, two launches
        xt = torch.randn(nt, dtype=torch.float64, device="cuda", generator=g)
        yt = torch.zeros(nt, dtype=torch.float64, device="cuda")
        run = lambda: dev.local_matvec_tiled(op, xt, yt)
    else:
        run = lambda: dev.local_matvec(op, x)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(json.dumps(dict(shape=[r, R, n], mode=mode, us=round(us, 2), tflops=round(F / us / 1e6, 3), frac=round(F / us / 1e6 / 37.1, 3))), flush=True)
dev.set_gemm_mode(0)
