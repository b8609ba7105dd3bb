"""GPU probe: the interface-stack update at the bench shape (r = 64, R = 3, n = 64) -- entry-point time of the natural-layout
kernel and of the image-based kernel it replaces (CUDA events, 200 calls each), phase time stamps of CTA 0 (%globaltimer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scikit_tt_b200._device import get_device  # noqa: E402

dev = get_device(0)
r, R, n = 64, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
L, x = rng.standard_normal((r, R, r)), rng.standard_normal((r, n, r))
A = rng.standard_normal((R, n, n, R))
for (b, q) in ((0, 1), (0, 2), (1, 1), (1, 2)):
    A[b, :, :, q] = 0.0
dL, dx, dA = dev.to_device(L), dev.to_device(x), dev.to_device(A)
F = 2 * r * R * r * n * r + 2 * r * r * R * R * n * n + 2 * r * R * r * n * r


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, dbg in (("natural-layout kernel", 0), ("image-based kernel", 64)):
    dev.set_debug(dbg)
    for side, fn in (("left", lambda: dev.stack_left_op(dL, dx, dA)), ("right", lambda: dev.stack_right_op(dL, dx, dA))):
        us = timed(fn)
        print(f"{name:24s} {side:5s}: {us:7.2f} us per update  {F / us / 1e6:6.2f} TFLOP/s  frac {F / us / 1e6 / 37.1:.3f}")
dev.set_debug(1)
for side, fn in (("left", lambda: dev.stack_left_op(dL, dx, dA)), ("right", lambda: dev.stack_right_op(dL, dx, dA))):
    for _ in range(3):
        fn()
    st = dev.scratch_peek(3600, 10)
    k = int(st[0])
    t = [int(v) for v in st[1:1 + k]]
    names = ["phase 1 (+ mask scan)", "grid barrier", "phase 2", "grid barrier", "ordered reduction"]
    print(side, "phases of CTA 0 [us]:", ", ".join(f"{nm} {(b - a) / 1e3:.2f}" for nm, a, b in zip(names, t, t[1:])),
          f"| total {(t[-1] - t[0]) / 1e3:.2f}")
dev.set_debug(0)
