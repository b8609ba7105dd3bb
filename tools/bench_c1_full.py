"""Timing + parity of BASELINE config 1 at FULL size on the GPU box (not the bench line): signaling_cascade(d=20) built from
the committed cores of the live reference's operator (tests/golden/euler_cascade.npz holds first / middle / last core; the
cascade repeats its middle core), implicit Euler via sle.als, solution rank 4 (dense 1024 x 1024 micro systems, LU with
partial pivoting on the GPU as in the reference).  CPU column: the numpy/scipy oracle on the host cores."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from util import load, rel_diff
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import ode
from oracle import ode as oode, tt as ott
z = load("euler_cascade")
d, n = 20, 64
opc = [z["op/first"]] + [z["op/mid"].copy() for _ in range(d - 2)] + [z["op/last"]]
iv = [np.zeros((1, n, 1, 1)) for _ in range(d)]
for c in iv: c[0, 0, 0, 0] = 1.0
ranks = [1] + [4] * (d - 1) + [1]
guess = ott.ortho_right([np.ones((ranks[i], n, 1, ranks[i + 1])) for i in range(d)])
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
run = lambda k: ode.implicit_euler(TT(opc), TT(iv), TT(guess), [1.0] * k, progress=False)
run(1); torch.cuda.synchronize()
t0 = time.perf_counter(); sol = run(K); torch.cuda.synchronize(); tg = time.perf_counter() - t0
t0 = time.perf_counter(); ref = oode.implicit_euler(opc, iv, guess, [1.0]); tc = time.perf_counter() - t0
print(json.dumps(dict(cfg="C1 full (d=20, n=64, r=4)", gpu_half_sweeps=2 * K, gpu_s=tg, gpu_hs_per_s=2 * K / tg, cpu_half_sweeps=2, cpu_s=tc,
                      cpu_hs_per_s=2 / tc, cpu_threads=len(os.sched_getaffinity(0)), step1_rel_diff=rel_diff(sol[1].cores, ref[1]))), flush=True)
