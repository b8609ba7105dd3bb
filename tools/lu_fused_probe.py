"""GPU probe: np.linalg.solve of one dense fp64 system in one launch (csrc/lu_fused.cu) -- dataflow form against the first
form (ctx debug bit 10), CUDA events, matrix restored by a device copy before every solve (the copy is timed separately)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device(0)
for N in (256, 512, 1024, 1536):
    rng = np.random.default_rng(N)
    M = rng.standard_normal((N, N)); f = rng.standard_normal(N)
    dM0, df = dev.to_device(M), dev.to_device(f)
    ipiv = torch.empty(N, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    work = dM0.clone(); x = df.clone()
    from scikit_tt_b200._device import _ptr
    def run():
        work.copy_(dM0); x.copy_(df)
        dev._check(dev.lib.sktt_lu_solve_fused(dev.h, N, _ptr(work), _ptr(x), _ptr(ipiv), _ptr(info)))
    def copies():
        work.copy_(dM0); x.copy_(df)
    def timed(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tc = timed(copies)
    res = {}
    for name, dbg in (("dataflow", 0), ("first form", 1024)):
        dev.set_debug(dbg)
        t = timed(run) - tc
        res[name] = t
        err = np.linalg.norm(x.cpu().numpy() - np.linalg.solve(M, f)) / np.linalg.norm(np.linalg.solve(M, f))
        print(f"N = {N:5d} {name:11s}: {t:9.1f} us per solve  ({2 / 3 * N ** 3 / t / 1e6:6.3f} TFLOP/s), rel err {err:.1e}")
    dev.set_debug(0)

N = 1024
rng = np.random.default_rng(N)
M = rng.standard_normal((N, N)); f = rng.standard_normal(N)
dev.set_debug(1)
x = dev.solve_fused(dev.to_device(M), dev.to_device(f))
st = [int(v) for v in dev.scratch_peek(3600, 14)]
dev.set_debug(0)
names = ["L11 + pivots + row list", "gather", "U12", "scatter", "rank-16 update", "panel (load, 16 columns, store)", "publish"]
print("CTA of the middle block, last update and own panel [us]:", ", ".join(f"{n} {(b - a) / 1e3:.2f}" for n, a, b in zip(names, st, st[1:])))
