"""GPU: the multi-GPU partitions of SURVEY.md 8e.  The block kernels of the sharded micro-matvec are checked on one GPU
against the oracle; the collective forms (NCCL, world size 2) run when the box has two GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import kernels as K, sle as osle, tt as ott

pytestmark = pytest.mark.gpu


def _operands(rng, r, R, m, r2, R2, cplx=False):
    def rnd(*s):
        a = rng.standard_normal(s)
        return a + 1j * rng.standard_normal(s) if cplx else a
    return rnd(r, R, r), rnd(R, m, m, R2), rnd(r2, R2, r2), rnd(r, m, r2)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(6, 3, 4, 5, 2), (128, 8, 16, 128, 8), (33, 2, 8, 17, 3)])
def test_block_rows_of_the_matvec(dev, shape, cplx):
    from scikit_tt_b200.solvers import multi
    r, R, m, r2, R2 = shape
    rng = np.random.default_rng(r + 7 * cplx)
    L, A, Rt, v = _operands(rng, r, R, m, r2, R2, cplx)
    want = K.micro_matvec_als(L, A, Rt, v)
    dL, dA, dR, dv = (dev.to_device(x) for x in (L, A, Rt, v))
    for world in (1, 2, 3, 8):
        got = np.zeros_like(want)
        for rank in range(world):
            lo, hi = multi.shard_bounds(r, world, rank)
            if hi > lo:
                got[lo:hi] = multi.matvec_rows_device(dev, dL, dA, dR, dv, lo, hi).cpu().numpy()
        assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want), (world, shape)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from scikit_tt_b200 import TT
        from scikit_tt_b200._device import get_device
        from scikit_tt_b200.solvers import multi
        dev = get_device()
        rng = np.random.default_rng(11)
        # sharded matvec (equal and ragged blocks) and CG on an SPD micro system
        for r in (64, 37):
            R, m, r2, R2 = 3, 16, 24, 3
            L, A, Rt, v = _operands(rng, r, R, m, r2, R2)
            dL, dA, dR, dv = (dev.to_device(x) for x in (L, A, Rt, v))
            y = multi.sharded_micro_matvec(dL, dA, dR, dv, dev=dev).cpu().numpy()
            want = K.micro_matvec_als(L, A, Rt, v)
            assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)
        r, m = 16, 8
        X = rng.standard_normal((r, r)); Y = rng.standard_normal((m, m)); Z = rng.standard_normal((r, r))
        L = np.stack([np.eye(r), X @ X.T / r], axis=1)                      # [a, b, c], symmetric slices
        A = np.stack([np.stack([Y @ Y.T / m + np.eye(m), np.zeros((m, m))], -1),
                      np.stack([np.zeros((m, m)), np.eye(m)], -1)], 0)      # [b, m, n, b2]
        Rt = np.stack([np.eye(r), Z @ Z.T / r + np.eye(r)], axis=1)
        f = rng.standard_normal((r, m, r))
        dL, dA, dR, df = (dev.to_device(x) for x in (L, A, Rt, f))
        u, it, rel = multi.cg_sharded(dL, dA, dR, df, tol=1e-13, dev=dev)
        M = K.micro_matrix_als(L, A, Rt)
        want = np.linalg.solve(M, f.reshape(-1)).reshape(f.shape)
        err = np.linalg.norm(u.cpu().numpy() - want) / np.linalg.norm(want)
        assert err < 1e-10, err
        # the fused form: rows of y stored into every rank's exchange buffer from the last contraction's epilogue
        # (peer memory, csrc/peer.cu) + flag barrier; ping-pong buffers over several consecutive matvecs
        px = multi.PeerExchange(dev, 64 * 16 * 24, torch.float64, None)
        try:
            for r in (64, 37):
                R, m, r2, R2 = 3, 16, 24, 3
                L, A, Rt, v = _operands(rng, r, R, m, r2, R2)
                dL, dA, dR, dv = (dev.to_device(x) for x in (L, A, Rt, v))
                for rep in range(3):
                    y = px.matvec(dL, dA, dR, dv).clone().cpu().numpy()
                    want = K.micro_matvec_als(L, A, Rt, v)
                    assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want), (r, rep)
                    dv = dev.to_device(y / np.linalg.norm(y))
                    v = y / np.linalg.norm(y)
            px.check()
        finally:
            px.close()
        # sle.als(..., group=): one system swept by both ranks with the micro-matvec sharded, against the one-GPU sweep
        import workloads
        from scikit_tt_b200.solvers import sle
        d, n, rk = 4, 16, 24
        opc = workloads.c4_spd_cores(d, n, 4)
        rhs = TT(workloads.rank1_rhs(d, n))
        x0 = TT(ott.ortho_right(workloads.random_guess(d, n, rk, seed=5)))
        one = sle.als(TT(opc), x0, rhs, repeats=1, solver='cg')
        keep, sle.SHARD_MIN_UNKNOWNS = sle.SHARD_MIN_UNKNOWNS, 0       # the size policy would keep a system this small replicated
        try:
            two = sle.als(TT(opc), x0, rhs, repeats=1, solver='cg', group=dist.group.WORLD)
        finally:
            sle.SHARD_MIN_UNKNOWNS = keep
        assert two.ranks == one.ranks
        assert ott.norm(ott.sub(two.cores, one.cores)) / ott.norm(one.cores) < 1e-8
        assert multi.sharded_stats["solves"] > 0 and multi.sharded_stats["worst_relres"] <= 1e-12
        # independent systems: four right-hand sides of one SPD system, two per GPU, gathered in order
        d, n, rk = 4, 6, 3
        S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
        I = np.eye(n)
        op = [np.stack([S, I], -1)[None]] + [np.stack([np.stack([I, np.zeros((n, n))], -1), np.stack([S, I], -1)], 0)
                                             for _ in range(d - 2)] + [np.stack([I, S], 0)[..., None]]
        x0 = ott.ortho_right([rng.standard_normal((1 if i == 0 else rk, n, 1, 1 if i == d - 1 else rk)) for i in range(d)])
        rhss = [[rng.standard_normal((1, n, 1, 1)) for _ in range(d)] for _ in range(4)]
        sols = multi.sle_als_batch(TT(op), TT(x0), [TT(b) for b in rhss], repeats=2)
        for b, s in zip(rhss, sols):
            ref = osle.als(op, x0, b, repeats=2)
            assert ott.norm(ott.sub(s.cores, ref)) / ott.norm(ref) < 1e-8
        out.put((rank, "ok"))
    except Exception as e:
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        multi.close_peer_exchanges()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_world_size_two_nccl():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
