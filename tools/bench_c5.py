"""Timing of BASELINE config 5 on the GPU box (64 co_oxidation(20) pressures, evp.als, rank-8 guess): batched GPU path vs the
one-at-a-time GPU path vs the oracle on one host core.  Not the bench line (bench.py carries the c5 leg)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import workloads
from scikit_tt_b200 import TT
import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200.solvers import evp
from scikit_tt_b200._device import get_device
dev = get_device()
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 8
solver = sys.argv[3] if len(sys.argv) > 3 else 'eigs'
d = 20
ops = []
for k in workloads.c5_pressures(64)[:nsys]:
    t = TT(workloads.co_oxidation_cores(d, k)).ortho_left().ortho_right()
    ops.append(tt.eye(t.row_dims) + t)
guess = tt.ones(ops[0].row_dims, [1] * d, ranks=rank).ortho_left().ortho_right()
R = 2
evp.als_batch(ops, guess, repeats=1, conv_eps=0, solver=solver); torch.cuda.synchronize()
l0 = dev.launches(); t0 = time.perf_counter()
res = evp.als_batch(ops, guess, repeats=R, conv_eps=0, solver=solver)
torch.cuda.synchronize(); tb = time.perf_counter() - t0; nl = dev.launches() - l0
t0 = time.perf_counter()
one = [evp.als(ops[j], guess, repeats=R, conv_eps=0, solver=solver) for j in range(min(4, nsys))]
torch.cuda.synchronize(); ts = (time.perf_counter() - t0) / min(4, nsys)
from threadpoolctl import threadpool_limits
from oracle import evp as oevp
with threadpool_limits(limits=1):
    t0 = time.perf_counter(); oevp.als(ops[0].cores, guess.cores, repeats=R, conv_eps=0, solver=solver); tc = time.perf_counter() - t0
print(json.dumps(dict(cfg=f"C5 {nsys} x co_oxidation(20) r={rank} {solver}", half_sweeps=2 * R * nsys, batch_s=tb, batch_hs_per_s=2 * R * nsys / tb,
                      launches_per_half_sweep=nl / (2 * R), single_gpu_hs_per_s=2 * R / ts, oracle_1core_hs_per_s=2 * R / tc,
                      stats=dict(evp.batch_stats), lam_batch=[float(r[0]) for r in res[:4]], lam_single=[float(r[0]) for r in one])), flush=True)
