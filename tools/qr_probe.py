"""Diagnostic (GPU box): CholeskyQR kernel -- passes, phase time stamps, accuracy on matrices of graded conditioning."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scikit_tt_b200._device import get_device
dev = get_device()
dev.set_debug(True)
rng = np.random.default_rng(0)
m, n = 4096, 64
U0, _ = np.linalg.qr(rng.standard_normal((m, n)))
V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
names = ["load", "loaded"] + ["gram", "sync1", "reduce", "sync2", "scaled", "chol", "subst"] * 6
for cond in (1e2, 1e8, 1e16, 1e30):
    A = (U0 * np.logspace(0, -np.log10(cond), n)) @ V0.T
    for kind in ("qr", "rq"):
        a = dev.to_device(A if kind == "qr" else np.ascontiguousarray(A.T))
        for _ in range(3):
            q = dev.qr(a) if kind == "qr" else dev.rq(a)
        status = dev.scratch_peek(1024, 2, ctype=__import__("ctypes").c_int)
        st = dev.scratch_peek(1536, 62)
        k = int(st[0]); t = np.array(st[1:1 + k], dtype=np.float64)
        dt = np.diff(t) / 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            q = dev.qr(a) if kind == "qr" else dev.rq(a)
        e1.record(); torch.cuda.synchronize()
        Q = q.cpu().numpy()
        if kind == "rq": Q = Q.T
        orth = np.linalg.norm(Q.T @ Q - np.eye(n))
        # how much of A lies outside span(Q), relative to ||A||
        out = np.linalg.norm(A - Q @ (Q.T @ A)) / np.linalg.norm(A)
        print(json.dumps(dict(cond=cond, kind=kind, fail=status[0], passes=status[1], us_per_call=round(e0.elapsed_time(e1) / 20 * 1e3, 1),
                              kernel_us=round((t[-1] - t[0]) / 1e3, 1), orth=orth, outside=out,
                              phases_us=[round(float(x), 1) for x in dt])), flush=True)
