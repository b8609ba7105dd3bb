"""CPU, world size 2 over gloo: the two multi-GPU partitions of the hot path -- block partition of
independent systems with a result gather, and the output-rank-sharded micro-matvec with its collective.  The per-item /
per-block arithmetic is the numpy oracle here (no GPU in this suite); the GPU forms are covered by
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scikit_tt_b200.solvers import multi
from oracle import kernels as K


def rows_numpy(L, A, Rt, v, lo, hi):
    """One rank's block y[lo:hi] of the micro-matvec on the host (the checker standing in for the CUDA block kernel)."""
    return np.einsum('abc,anp,bmnq,pqs->cms', L[:, :, lo:hi], v, A, Rt, optimize=True)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 63, 64, 65):
        for w in (1, 2, 3, 4, 8):
            blocks = [multi.shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multi.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # partition 1: 7 independent "systems" (ragged over 2 ranks), results gathered in input order everywhere
        items = [np.arange(3.0) + i for i in range(7)]
        calls = []
        res = multi.map_sharded(lambda a: (calls.append(1), float(a.sum()))[1], items)
        assert res == [float(a.sum()) for a in items]
        lo, hi = multi.shard_bounds(7, world, rank)
        assert len(calls) == hi - lo                                  # each rank only worked on its block
        # partition 2: sharded micro-matvec, equal (r = 6) and ragged (r = 5) blocks, real and complex
        rng = np.random.default_rng(3)                                # same operands on every rank
        for r, cplx in ((6, False), (5, False), (5, True)):
            R, m, r2, R2 = 3, 4, 5, 2
            def rnd(*s):
                a = rng.standard_normal(s)
                return a + 1j * rng.standard_normal(s) if cplx else a
            L, A, Rt, v = rnd(r, R, r), rnd(R, m, m, R2), rnd(r2, R2, r2), rnd(r, m, r2)
            y = multi.sharded_micro_matvec(L, A, Rt, v, rows=rows_numpy)
            want = K.micro_matvec_als(L, A, Rt, v)
            assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)
        # the CG driver of the sharded micro solve (host logic of multi.solve_sharded): identical scalar recurrences on
        # every rank around a matvec whose rows are computed rank by rank and assembled by the collective
        r, m = 6, 4
        X, Y, Z = rng.standard_normal((r, r)), rng.standard_normal((m, m)), rng.standard_normal((r, r))
        L = np.stack([np.eye(r), X @ X.T / r], axis=1)
        A = np.stack([np.stack([Y @ Y.T / m + np.eye(m), np.zeros((m, m))], -1),
                      np.stack([np.zeros((m, m)), np.eye(m)], -1)], 0)
        Rt = np.stack([np.eye(r), Z @ Z.T / r + np.eye(r)], axis=1)
        f = torch.from_numpy(rng.standard_normal((r, m, r)))

        class HostDev:                                                # level-1 helpers of the device wrapper, on the host
            def axpby(self, alpha, x, beta, y, out=None):
                res = alpha * x + beta * y
                if out is None:
                    return res
                out.copy_(res)
                return out

            def dotc(self, x, y):
                return float(torch.dot(x, y))
        mv = lambda v: torch.from_numpy(multi.sharded_micro_matvec(L, A, Rt, v.numpy(), rows=rows_numpy))
        multi.sharded_stats.update(solves=0, matvecs=0, worst_relres=0.0)
        for guess in (None, torch.from_numpy(rng.standard_normal((r, m, r))), 1e30 * torch.ones((r, m, r), dtype=torch.float64)):
            u = multi.solve_sharded(HostDev(), mv, f, guess)
            want = np.linalg.solve(K.micro_matrix_als(L, A, Rt), f.numpy().reshape(-1))
            assert np.linalg.norm(u.numpy() - want) <= 1e-11 * np.linalg.norm(want)
        assert multi.sharded_stats["solves"] == 3 and multi.sharded_stats["worst_relres"] <= 1e-12
        out.put((rank, "ok"))
    except Exception as e:                                            # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
