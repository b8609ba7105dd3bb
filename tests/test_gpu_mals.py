"""GPU: MALS at the block size of BASELINE config 3 (two-site super-cores of 4096 x 4096, 16.7 million unknowns per micro
system): the truncated SVD by subspace iteration against numpy, and a full-size two-site sweep (properties: no reference
can exist -- the reference's two-site micro matrix would be 2 PiB)."""
import numpy as np
import pytest
import torch

import workloads
from oracle import tt as ott, sle as osle

pytestmark = pytest.mark.gpu


def test_truncated_svd_of_a_4096_block(dev):
    """Top-64 singular triplets of a 4096 x 4096 matrix with a decaying spectrum: singular values to 1e-12 relative to the
    largest, dominant left / right subspaces to 1e-9, the reference's rank rule (strict relative threshold, then max_rank)."""
    rng = np.random.default_rng(0)
    n, k = 4096, 64
    U0, _ = np.linalg.qr(rng.standard_normal((n, 200)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, 200)))
    s = 10.0 ** (-0.12 * np.arange(200))                      # 1 ... 1e-24: 100 of them above 1e-12
    A = (U0 * s) @ V0.T
    dA = dev.to_device(A)
    before = dict(dev.svd_topk_stats)
    U, S, Vh, rank = dev.svd_truncated(dA, threshold=1e-12, max_rank=k)
    assert dev.svd_topk_stats["calls"] == before["calls"] + 1 and dev.svd_topk_stats["fallbacks"] == before["fallbacks"]
    assert rank == k
    Sg = S[:k].cpu().numpy()
    assert np.abs(Sg - s[:k]).max() <= 1e-12 * s[0]
    Ug, Vg = U[:, :k].cpu().numpy(), Vh[:k, :].cpu().numpy()
    assert np.linalg.norm(Ug.T @ Ug - np.eye(k)) < 1e-12 and np.linalg.norm(Vg @ Vg.T - np.eye(k)) < 1e-12
    assert np.linalg.norm(Ug @ Ug.T @ U0[:, :k] - U0[:, :k]) < 1e-8     # gap s_64 / s_65 = 1.3
    assert np.linalg.norm(Vg.T @ Vg @ V0[:, :k] - V0[:, :k]) < 1e-8
    # the threshold decides before max_rank does: only 40 singular values above 1e-4.75 relative
    U, S, Vh, rank = dev.svd_truncated(dA, threshold=10.0 ** -4.75, max_rank=k)
    assert rank == int((s / s[0] > 10.0 ** -4.75).sum()) == 40
    # small matrices and unbounded ranks keep the full Jacobi SVD
    c0 = dev.svd_topk_stats["calls"]
    dev.svd_truncated(dev.to_device(A[:300, :300]), threshold=1e-12, max_rank=8)
    dev.svd_truncated(dev.to_device(A[:1100, :1100]), threshold=1e-12, max_rank=np.inf)
    assert dev.svd_topk_stats["calls"] == c0


def test_mals_at_the_c3_block_size(dev):
    """sle.mals with n = 64 and solution rank 64 on a four-core instance of the C3 operator family: the middle two-site
    system has 64 * 64 * 64 * 64 = 16 777 216 unknowns and its super-core is a 4096 x 4096 matrix.  Properties: max_rank and
    the boundary ranks are respected, the global residual falls from one sweep to two, two runs return bit-identical cores."""
    from scikit_tt_b200 import TT
    import scikit_tt_b200.tensor_train as tt
    from scikit_tt_b200.solvers import sle
    d, n, r = 4, 64, 64
    opc, rhsc, x0c = workloads.workload_cores(d, n, r)
    op, rhs = TT(opc), TT(rhsc)
    x0 = TT(ott.ortho_right(x0c))
    bnorm = np.prod([np.linalg.norm(c) for c in rhsc])
    c0 = dev.svd_topk_stats["calls"]
    one = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=64)
    assert dev.svd_topk_stats["calls"] > c0                               # the 4096 x 4096 blocks took the subspace route
    assert one.ranks[0] == one.ranks[-1] == 1 and max(one.ranks) <= 64 and one.row_dims == [n] * d
    r1 = tt.residual_error(op, one, rhs) / bnorm
    two = sle.mals(op, x0, rhs, repeats=2, threshold=1e-12, max_rank=64)
    r2 = tt.residual_error(op, two, rhs) / bnorm
    assert r2 <= r1 * (1 + 1e-9) and r1 < 1e-6, (r1, r2)
    again = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=64)
    assert all(np.array_equal(a, b) for a, b in zip(one.cores, again.cores))
