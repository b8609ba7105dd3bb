"""C4 under torchrun (N ranks): one sle.als sweep at solution rank r with the micro-matvec rank-sharded over peer memory.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_c4_sharded.py 256"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import workloads
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from scikit_tt_b200 import TT
import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200.solvers import sle, multi
d, n, R = 10, 16, 8
for r in [int(a) for a in sys.argv[1:]] or [128]:
    op, rhs = TT(workloads.c4_spd_cores(d, n, R)), TT(workloads.rank1_rhs(d, n))
    x0 = TT(workloads.random_guess(d, n, r, seed=1)).ortho_right()
    kw = dict(group=dist.group.WORLD) if world > 1 else {}
    sle.als(op, x0, rhs, repeats=1, **kw)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    sol = sle.als(op, x0, rhs, repeats=1, **kw)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        bn = np.prod([np.linalg.norm(c) for c in rhs.cores])
        print(json.dumps(dict(cfg=f"C4 r={r}", world=world, seconds=dt, hs_per_s=2 / dt, residual=float(tt.residual_error(op, sol, rhs) / bn),
                              sharded=dict(multi.sharded_stats))), flush=True)
if world > 1:
    dist.destroy_process_group()
