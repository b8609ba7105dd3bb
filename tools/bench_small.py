"""Timing of the reference-runnable configs on the GPU box: C1 (signaling_cascade(20), implicit Euler via sle.als, r=4)
and C2 (co_oxidation(20), evp.als, r=8), GPU path against the numpy/scipy oracle on the host cores.  Not the bench."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from util import load, cores, cascade_operator, rel_diff
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import ode, evp
from scikit_tt_b200._device import get_device
from oracle import ode as oode, evp as oevp
dev = get_device()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps, out

z = load("euler_cascade")
opc = cascade_operator(z)
iv, guess = cores(z, "iv"), cores(z, "guess")
K = 3
l0 = dev.launches()
tg, sol = timed(lambda: ode.implicit_euler(TT(opc), TT(iv), TT(guess), [1.0] * K, progress=False))
nl = (dev.launches() - l0) // 4
t0 = time.perf_counter(); ref = oode.implicit_euler(opc, iv, guess, [1.0] * K); tc = time.perf_counter() - t0
print(json.dumps(dict(cfg="C1", half_sweeps=2 * K, gpu_s=tg, cpu_s=tc, gpu_hs_per_s=2 * K / tg, cpu_hs_per_s=2 * K / tc,
                      launches_per_run=nl, step1_rel_diff=rel_diff(sol[1].cores, ref[1]))), flush=True)

z = load("c2_cooxidation20")
op, x0 = cores(z, "op"), cores(z, "x0")
R = 2
l0 = dev.launches()
tg, (lam, x, it) = timed(lambda: evp.als(TT(op), TT(x0), repeats=R, conv_eps=0, solver='eig'))
nl = (dev.launches() - l0) // 4
t0 = time.perf_counter(); lam_o, x_o, it_o = oevp.als(op, x0, repeats=R, conv_eps=0, solver='eig'); tc = time.perf_counter() - t0
print(json.dumps(dict(cfg="C2 r=8 eig", half_sweeps=2 * R, gpu_s=tg, cpu_s=tc, gpu_hs_per_s=2 * R / tg, cpu_hs_per_s=2 * R / tc,
                      launches_per_run=nl, lam_gpu=float(lam), lam_oracle=float(np.real(lam_o[0]) if np.ndim(lam_o) else lam_o),
                      lam_reference=float(z["lam"]))), flush=True)
