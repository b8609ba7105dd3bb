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


import time
t0 = time.perf_counter()
for _ in range(2000):
    dev.stack_left_op(dL, dx, dA)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host: {1e6 * (t1 - t0) / 2000:.1f} us per call to issue, {1e6 * (t2 - t0) / 2000:.1f} us per call until the device is done")
for name, dbg in (("natural-layout kernel", 0), ("natural-layout, plain launch", 128), ("image-based kernel", 64)):
    dev.set_debug(dbg)
    for side, fn in (("left", lambda: dev.stack_left_op(dL, dx, dA)), ("right", lambda: dev.stack_right_op(dL, dx, dA))):
        us = timed(fn)
        print(f"{name:30s} {side:5s}: {us:7.2f} us per update  {F / us / 1e6:6.2f} TFLOP/s  frac {F / us / 1e6 / 37.1:.3f}")
dev.set_debug(1)
for side, fn in (("left", lambda: dev.stack_left_op(dL, dx, dA)), ("right", lambda: dev.stack_right_op(dL, dx, dA))):
    for _ in range(3):
        fn()
    st = [int(v) for v in dev.scratch_peek(3600, 24)]
    us = lambda i, j: (st[j] - st[i]) / 1e3
    print(f"{side}: CTA 0 [us]: phase 1 (+ conj-side rows) {us(0, 1):.2f}, grid barrier {us(1, 2):.2f}, phase 2 {us(2, 3):.2f}, "
          f"grid barrier {us(3, 4):.2f}, ordered reduction {us(4, 5):.2f} | total {us(0, 5):.2f}")
    print(f"      phase 1: until the consumers wait {us(0, 20):.2f}, first K group arrives {us(20, 21):.2f}, contraction "
          f"{us(21, 22):.2f}, K-half exchange + store {us(22, 23):.2f}")
    print(f"      phase 2: operator rows there {us(2, 12):.2f}, first T1 slot arrives {us(12, 13):.2f}, second contraction "
          f"{us(13, 14):.2f}, T2 exchange {us(14, 15):.2f}, last contraction {us(15, 16):.2f}, partial store {us(16, 17):.2f}")
dev.set_debug(0)
