"""CPU: the host side of the drop-in surface -- TT container semantics, argument validation of ortho_*
(tests/test_tensor_train.py:371-382, :411-422 of the reference), the C-ABI library loads and exports every
symbol include/sktt_b200.h declares, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200 import TT, _lib
from oracle import tt as ott

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "sktt_b200.h")).read()
    declared = set(re.findall(r"\b(sktt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sktt_ctx", "sktt_idx2", "sktt_local_op"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().sktt_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scikit_tt_b200.solvers import sle
    t = tt.ones([2, 2], [1, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sle.als(tt.eye([2, 2]), t, t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t.ortho_right()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scikit_tt_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_tt_container_and_algebra():
    rng = np.random.default_rng(0)
    a = TT([rng.standard_normal(s) for s in ((1, 2, 3, 2), (2, 3, 2, 3), (3, 2, 2, 1))])
    b = TT([rng.standard_normal(s) for s in ((1, 2, 3, 1), (1, 3, 2, 2), (2, 2, 2, 1))])
    assert (a.order, a.row_dims, a.col_dims, a.ranks) == (3, [2, 3, 2], [3, 2, 2], [1, 2, 3, 1])
    assert np.allclose((a + b).full(), a.full() + b.full())
    assert np.allclose((a - b).full(), a.full() - b.full())
    assert np.allclose((2.5 * a).full(), 2.5 * a.full()) and np.allclose((a * 2).full(), 2 * a.full())
    assert (a + b).ranks == [1, 3, 5, 1]
    x = TT([rng.standard_normal(s) for s in ((1, 3, 1, 2), (2, 2, 1, 2), (2, 2, 1, 1))])
    assert np.allclose((a @ x).matricize(), a.matricize() @ x.matricize())
    assert np.allclose(a.transpose().matricize(), a.matricize().T)
    assert a.isoperator() and not x.isoperator()
    c = a.copy()
    c.cores[0][:] = 0
    assert np.abs(a.cores[0]).sum() > 0
    assert np.allclose(tt.eye([2, 3]).matricize(), np.eye(6))
    assert tt.ones([2, 3], [1, 1], ranks=4).ranks == [1, 4, 1]
    assert abs(np.linalg.norm(tt.uniform([2, 3, 4], ranks=3, norm=2.0).full()) - 2.0) < 1e-12
    assert np.allclose(TT(a.full()).full(), a.full())
    pos = TT([np.abs(c) for c in a.cores])
    assert abs(pos.norm(p=1) - np.max(pos.matricize().sum(axis=0))) < 1e-10
    v = TT([np.abs(c) for c in x.cores])
    assert abs(v.norm(p=1) - v.full().sum()) < 1e-12
    with pytest.raises(ValueError):
        TT([np.zeros((1, 2, 2, 2)), np.zeros((3, 2, 2, 1))])
    with pytest.raises(ValueError):
        TT([np.zeros((2, 2, 2))])
    with pytest.raises(TypeError):
        TT("x")
    res = tt.residual_error(a, x, a @ x)
    assert res < 1e-12 * (a @ x).full().size
    y = TT([rng.standard_normal(c.shape) for c in (a @ x).cores[:1]] + (a @ x).cores[1:])
    dense = np.linalg.norm(a.matricize() @ x.matricize() - y.matricize())
    assert abs(tt.residual_error(a, x, y) - dense) < 1e-10 * dense


def test_ortho_argument_validation():
    t = tt.ones([2, 2, 2], [1, 1, 1], ranks=2)
    for fn in (t.ortho_left, t.ortho_right):
        with pytest.raises(ValueError):
            fn(max_rank=0)
        with pytest.raises(ValueError):
            fn(threshold=-1)
        with pytest.raises(TypeError):
            fn(start_index="a")
        with pytest.raises(TypeError):
            fn(end_index="b")


def test_warm_start_plumbing_of_the_sweep():
    """Host logic of the warm starts (solvers/sle.py): which solver settings want a starting vector, and how the gauge factor
    of the neighbour that was just orthonormalised is pushed into the next core (left: factor @ core, right: core @ factor;
    mismatching factors are ignored).  The device product is replaced by torch on the CPU here."""
    import torch
    from scikit_tt_b200.solvers import sle, _local

    class FakeDev:
        def gauge_push(self, carry, core, left):
            return carry @ core if left else core @ carry

    assert sle._wants_guess('cg', 10) and sle._wants_guess('gmres', 10)
    assert not sle._wants_guess('dense', 10 ** 9)
    assert sle._wants_guess('solve', _local.SMALL_DENSE_LIMIT + 1) and not sle._wants_guess('solve', _local.SMALL_DENSE_LIMIT)
    g = torch.Generator().manual_seed(0)
    core = torch.randn(5, 7, 3, dtype=torch.float64, generator=g)
    R = torch.randn(4, 5, dtype=torch.float64, generator=g)                   # k x r: from the QR of the core to the left
    out = sle._pushed(FakeDev(), R, core, left=True)
    assert out.shape == (4, 7, 3) and torch.allclose(out, torch.einsum('ka,anb->knb', R, core))
    Rp = torch.randn(3, 2, dtype=torch.float64, generator=g)                  # r2 x k: from the RQ of the core to the right
    out = sle._pushed(FakeDev(), Rp, core, left=False)
    assert out.shape == (5, 7, 2) and torch.allclose(out, torch.einsum('anb,bk->ank', core, Rp))
    assert sle._pushed(FakeDev(), None, core, left=True) is None
    assert sle._pushed(FakeDev(), torch.zeros(4, 6, dtype=torch.float64), core, left=True) is None     # rank changed
    assert sle._pushed(FakeDev(), torch.zeros(4, 2, dtype=torch.float64), core, left=False) is None


def test_multi_gpu_policy_surface():
    """Host-side surface of the sharded sweep: the size policy of sle.als(group=) and the cache of peer exchange buffers exist
    with the documented defaults (their behaviour on GPUs is covered by tests/test_gpu_multi.py)."""
    from scikit_tt_b200.solvers import sle, multi
    assert sle.SHARD_MIN_UNKNOWNS == 1 << 20
    assert callable(multi.peer_exchange) and callable(multi.close_peer_exchanges)
    multi.close_peer_exchanges()                              # nothing cached: a no-op
    assert multi._PX_CACHE == {}
