"""GPU: the public solver surface (scikit_tt_b200.solvers.sle / evp / ode, TT.ortho_*) against the golden
vectors produced by the live reference (tests/golden/*.npz) and against the numpy oracle.

Tolerances are the north star's: 1e-8 on the solution TT measured as ||x - x_ref|| / ||x_ref||, 1e-10 on
residual / eigenvalue values -- except where the reference itself is only reproducible to a coarser level,
which is stated at the assertion.
"""
import numpy as np
import pytest

import scikit_tt_b200.tensor_train as tt
from scikit_tt_b200 import TT
from scikit_tt_b200.solvers import sle, evp, ode, _local
from oracle import sle as osle, evp as oevp, tt as ott
from util import load, cores, rel_diff, rel_diff_up_to_phase, cascade_operator

pytestmark = pytest.mark.gpu

SOL_TOL = 1e-8
VAL_TOL = 1e-10


def T(z, prefix):
    return TT(cores(z, prefix))


def check_residual(op, sol, rhs, ref):
    """1e-10 agreement of the residual VALUE with the reference's; residuals that are themselves at rounding level
    (<= 1e-11: the solution is exactly representable at this rank) can only be compared by magnitude."""
    res = osle.residual(op.cores, sol.cores, rhs.cores)
    assert abs(res - ref) <= VAL_TOL * max(ref, 1e-300) + 1e-11, (res, ref)


def test_sle_toeplitz(dev):
    """The reference's own acceptance problem (tests/test_sle.py:13-85)."""
    z = load("sle_toeplitz")
    op, rhs, x0 = T(z, "op"), T(z, "rhs"), T(z, "x0")
    x0_before = [c.copy() for c in x0.cores]
    for solver in ("solve", "lu"):
        sol = sle.als(op, x0, rhs, repeats=1, solver=solver)
        assert isinstance(sol, TT) and sol.ranks == TT(cores(z, f"als_{solver}")).ranks
        assert rel_diff(sol.cores, cores(z, f"als_{solver}")) < SOL_TOL
        m = sle.mals(op, x0, rhs, repeats=1, solver=solver, threshold=1e-14, max_rank=10)
        assert m.ranks == TT(cores(z, f"mals_{solver}")).ranks
        assert rel_diff(m.cores, cores(z, f"mals_{solver}")) < SOL_TOL
    assert all(np.array_equal(a, b) for a, b in zip(x0.cores, x0_before))      # inputs are never mutated (sle.py:45)
    sol = sle.als(op, x0, rhs)
    dense = sol.matricize().reshape(-1)
    assert np.linalg.norm(dense - z["dense_solution"]) / np.linalg.norm(z["dense_solution"]) < 1e-7
    check_residual(op, sol, rhs, float(z["als_solve_residual"]))


@pytest.mark.parametrize("name", ["sle_laplace", "sle_random_spd"])
def test_sle_spd_dense_and_matrix_free(dev, name):
    z = load(name)
    op, rhs, x0 = T(z, "op"), T(z, "rhs"), T(z, "x0")
    reps = [1, 2] if name == "sle_laplace" else [2]
    for rep in reps:
        for solver in ("solve", "cg", "gmres"):
            sol = sle.als(op, x0, rhs, repeats=rep, solver=solver)
            assert rel_diff(sol.cores, cores(z, f"als_rep{rep}")) < SOL_TOL, (rep, solver)
            check_residual(op, sol, rhs, float(z[f"als_rep{rep}_residual"]))
    rmax = max(x0.ranks)
    for solver in ("solve", "cg"):
        m = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=rmax, solver=solver)
        assert m.ranks == TT(cores(z, "mals")).ranks
        assert rel_diff(m.cores, cores(z, "mals")) < SOL_TOL, solver


def test_sle_complex(dev):
    z = load("sle_complex")
    op, rhs, x0 = T(z, "op"), T(z, "rhs"), T(z, "x0")
    for solver in ("solve", "gmres"):
        sol = sle.als(op, x0, rhs, repeats=2, solver=solver)
        assert sol.cores[0].dtype == np.complex128
        assert rel_diff(sol.cores, cores(z, "als_rep2")) < SOL_TOL, solver
    m = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=3)
    assert rel_diff(m.cores, cores(z, "mals")) < SOL_TOL


def test_sle_matrix_free_beyond_the_reference(dev, monkeypatch):
    """A micro system larger than the dense limit: the residual must decrease monotonically over sweeps and the
    result must agree with the dense path run on the same problem with the limit lifted."""
    rng = np.random.default_rng(42)
    d, n, r = 5, 16, 8
    S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1)) * np.sqrt(1e-3)
    first = tt.build_core([[S, D, 1]])
    mid = tt.build_core([[1, 0, 0], [D, 0, 0], [S, D, 1]])
    last = tt.build_core([[1], [D], [S]])
    op = TT([first] + [mid.copy() for _ in range(d - 2)] + [last])
    rhs = TT([rng.standard_normal((1, n, 1, 1)) for _ in range(d)])
    x0 = TT(ott.ortho_right([rng.standard_normal((1 if i == 0 else r, n, 1, 1 if i == d - 1 else r)) for i in range(d)]))
    monkeypatch.setattr(_local, "DENSE_LIMIT", 64)            # r*n*r = 1024 > 64 -> Krylov path
    free = [sle.als(op, x0, rhs, repeats=k) for k in (1, 2)]
    monkeypatch.setattr(_local, "DENSE_LIMIT", 1 << 20)
    dense = [sle.als(op, x0, rhs, repeats=k) for k in (1, 2)]
    res = [osle.residual(op.cores, s.cores, rhs.cores) for s in free]
    assert res[1] <= res[0] * (1 + 1e-9)
    for a, b in zip(free, dense):
        assert rel_diff(a.cores, b.cores) < SOL_TOL
    ref = osle.als(op.cores, x0.cores, rhs.cores, repeats=2)
    assert rel_diff(free[1].cores, ref) < SOL_TOL


def test_implicit_euler(dev):
    z = load("euler_cascade")
    op = TT(cascade_operator(z))
    iv, guess = T(z, "iv"), T(z, "guess")
    sol = ode.implicit_euler(op, iv, guess, [1.0] * 3, progress=False)
    assert sol[0] is iv and len(sol) == 4                                       # ode.py:301
    # one-repeat ALS on this lossy rank-3 problem amplifies rounding differences 3-50x per step:
    # compare per step, feeding the reference's own previous step to the solver
    assert rel_diff(sol[1].cores, cores(z, "als/step1")) < SOL_TOL
    prev, g = iv, guess
    for k in range(1, 4):
        nxt = ode.implicit_euler(op, prev, g, [1.0], progress=False)[1]
        assert rel_diff(nxt.cores, cores(z, f"als/step{k}")) < SOL_TOL, k
        prev = g = T(z, f"als/step{k}")
    gen, iv2, g2 = T(z, "gen"), T(z, "iv2"), T(z, "guess2")
    for p in (1, 2, 0):
        sol = ode.implicit_euler(gen, iv2, g2, [0.1, 0.2, 0.1], repeats=2, tt_solver='mals', max_rank=3, normalize=p,
                                 progress=False)
        for k in range(1, 4):
            assert rel_diff(sol[k].cores, cores(z, f"mals_norm{p}/step{k}")) < SOL_TOL, (p, k)


def test_evp_laplace(dev):
    z = load("evp_laplace")
    op, x0 = T(z, "op"), T(z, "x0")
    lam, x, it = evp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    ref = float(z["eigh/lam"])
    assert isinstance(x, TT) and it == int(z["eigh/it"])
    assert abs(lam - ref) < VAL_TOL * abs(ref)
    assert rel_diff_up_to_phase(x.cores, cores(z, "eigh/x")) < SOL_TOL
    for solver in ("eig", "eigs"):
        lam, x, it = evp.als(op, x0, repeats=4, conv_eps=0, solver=solver, sigma=0.0)
        ref = float(z["eig/lam"])
        assert abs(lam - ref) < VAL_TOL * max(abs(ref), 1.0), solver
        assert rel_diff_up_to_phase(x.cores, cores(z, "eig/x")) < 1e-7, solver
    lam, xs, it = evp.als(op, x0, repeats=3, conv_eps=0, solver='eigh', number_ev=2)
    assert np.allclose(lam, z["eigh2/lam"], rtol=VAL_TOL, atol=0)
    assert len(xs) == 2
    for j in range(2):
        assert rel_diff_up_to_phase(xs[j].cores, cores(z, f"eigh2/x{j}")) < 1e-7
    lam, x, it = evp.als(op, x0, operator_gevp=T(z, "gevp"), repeats=3, conv_eps=0, solver='eigh')
    ref = float(z["gevp_eigh/lam"])
    assert abs(lam - ref) < VAL_TOL * abs(ref)
    assert rel_diff_up_to_phase(x.cores, cores(z, "gevp_eigh/x")) < 1e-7
    lam1, x1, _ = evp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    lam2, x2, _ = evp.als(op, x0, previous=[x1], shift=-lam1, repeats=4, conv_eps=0, solver='eigh')
    ref = float(z["defl/lam2"])
    assert abs(lam2 - ref) < 1e-9 * abs(ref)


def test_evp_convergence_stop(dev):
    z = load("evp_laplace")
    op, x0 = T(z, "op"), T(z, "x0")
    ref = oevp.als(op.cores, x0.cores, repeats=20, conv_eps=1e-6, solver='eigh')
    lam, x, it = evp.als(op, x0, repeats=20, conv_eps=1e-6, solver='eigh')
    assert it == ref[2] and it < 20
    assert abs(lam - ref[0]) < VAL_TOL * abs(ref[0])


def test_evp_cooxidation_reference_noise_limited(dev):
    """co_oxidation with a rank-deficient guess is rounding-chaotic in the reference itself (SURVEY.md 8c:
    reference-vs-reference spread with 1 vs 8 BLAS threads).  What is stable is that every implementation returns an
    eigenpair of the projected problem close to sigma = 1 with a small residual; compare residual quality, not values."""
    z = load("evp_cooxidation")
    op = T(z, "op_raw")
    opI = tt.eye(op.row_dims) + op
    lam, x, it = evp.als(opI, T(z, "x0"), repeats=5, conv_eps=0, solver='eig', sigma=1)
    xr = TT(cores(z, "eig/x"))

    def quality(l, v):
        return ott.norm(ott.sub(ott.matmul(opI.cores, v.cores), ott.scale(v.cores, l))) / ott.norm(v.cores)

    q_new, q_ref = quality(lam, x), quality(float(z["eig/lam"]), xr)
    assert np.isfinite(lam) and q_new <= 10 * q_ref + 1e-6, (lam, q_new, q_ref)


def test_ortho(dev):
    z = load("ortho")
    t0 = cores(z, "t")
    cases = (("left", lambda t: t.ortho_left()), ("right", lambda t: t.ortho_right()),
             ("left_thr", lambda t: t.ortho_left(threshold=0.2)), ("right_mr", lambda t: t.ortho_right(max_rank=2)),
             ("ortho", lambda t: t.ortho(threshold=1e-12, max_rank=3)))
    for key, fn in cases:
        t = TT([c.copy() for c in t0])
        out = fn(t)
        assert out is t                                                         # in place and returned
        ref = cores(z, key)
        assert t.ranks == ott.ranks_of(ref), key
        assert rel_diff(t.cores, ref) < 1e-12, key
    t = TT([c.copy() for c in t0]).ortho_left()
    for c in t.cores[:-1]:                                                      # tests/test_tensor_train.py:352-358
        m = c.reshape(-1, c.shape[3])
        assert np.linalg.norm(m.T @ m - np.eye(m.shape[1])) < 1e-12
    t = TT([c.copy() for c in t0]).ortho_right()
    for c in t.cores[1:]:
        m = c.reshape(c.shape[0], -1)
        assert np.linalg.norm(m @ m.T - np.eye(m.shape[0])) < 1e-12
    assert abs(TT([c.copy() for c in t0]).norm() - float(z["norm2"])) < 1e-12 * float(z["norm2"])
    assert abs(TT([np.abs(c) for c in t0]).norm(p=1) - float(z["norm1"])) < 1e-12 * float(z["norm1"])
    o = TT([c.copy() for c in cores(z, "ones")]).ortho_right(threshold=1e-10)
    assert o.ranks == ott.ranks_of(cores(z, "ones_right"))
    assert rel_diff(o.cores, cores(z, "ones_right")) < 1e-12
    tc = cores(z, "tc")
    assert rel_diff(TT([c.copy() for c in tc]).ortho_left().cores, cores(z, "tc_left")) < 1e-12
    assert rel_diff(TT([c.copy() for c in tc]).ortho_right().cores, cores(z, "tc_right")) < 1e-12
    # rank-deficient input without threshold keeps the rank and stays orthonormal (tt.ones(..., ranks=4))
    o = tt.ones([3, 4, 5], [1, 1, 1], ranks=4).ortho_right()
    assert o.ranks == [1, 4, 4, 1]
    for c in o.cores[1:]:
        m = c.reshape(c.shape[0], -1)
        assert np.linalg.norm(m @ m.T - np.eye(m.shape[0])) < 1e-12
    assert abs(o.norm() - tt.ones([3, 4, 5], [1, 1, 1], ranks=4).norm()) < 1e-9


def test_c2_first_step_and_completion(dev):
    """BASELINE config 2 at full size.  The reference is chaotic there (tests/test_c2_noise.py), so the check is: the
    first micro-step -- identical inputs on both sides -- picks the same eigenvalue to the 1e-7 its conditioning allows
    (|M| ~ 1e8, eigenvalue ~ 1), the sweep completes, and the return types / ranks follow evp.py:168-179."""
    from scikit_tt_b200.solvers import evp as gevp
    from oracle import evp as oevp
    z = load("c2_cooxidation20")
    op, x0 = cores(z, "op"), cores(z, "x0")
    first = {}
    orig, oorig = gevp._local_eig, oevp._local_eig

    def spy(d, M, B, k, solver, sigma):
        lam, vec = orig(d, M, B, k, solver, sigma)
        first.setdefault("gpu", complex(lam[0].item()))
        return lam, vec

    def ospy(M, B, k, solver, sigma, real):
        lam, vec = oorig(M, B, k, solver, sigma, real)
        first.setdefault("oracle", complex(lam[0]))
        return lam, vec
    gevp._local_eig, oevp._local_eig = spy, ospy
    try:
        lam, x, it = gevp.als(TT(op), TT(x0), repeats=1, conv_eps=0, solver='eig')
        oevp.als(op, x0, repeats=1, conv_eps=0, solver='eig')
    finally:
        gevp._local_eig, oevp._local_eig = orig, oorig
    assert abs(first["gpu"] - first["oracle"]) < 1e-6 * abs(first["oracle"])
    assert isinstance(lam, float) and isinstance(x, TT) and it == 1
    assert x.ranks[0] == x.ranks[-1] == 1 and max(x.ranks) <= 8 and np.isfinite(lam)


def test_c3_full_size_properties(dev):
    """BASELINE config 3 at full size (d=32, n=64, R=3, r=64; 262 144 unknowns per micro system) -- no reference result can
    exist (512 GiB micro matrices), so the checks are the size-independent ones: the global residual ||A x - b|| / ||b||
    falls from one sweep to two and ends below 1e-11, ranks and shapes are those of the guess, and two runs on the same
    inputs return bit-identical cores (every reduction of the path is ordered)."""
    import bench
    opc, rhsc, x0c = bench.workload_cores(32, 64, 64)
    op, rhs = TT(opc), TT(rhsc)
    x0 = TT(ott.ortho_right(x0c))
    one = sle.als(op, x0, rhs, repeats=1)
    two = sle.als(op, x0, rhs, repeats=2)
    bnorm = np.prod([np.linalg.norm(c) for c in rhsc])
    r1, r2 = tt.residual_error(op, one, rhs) / bnorm, tt.residual_error(op, two, rhs) / bnorm
    assert r2 < r1 < 1e-9 and r2 < 1e-11, (r1, r2)
    assert two.ranks == x0.ranks and two.row_dims == x0.row_dims and two.col_dims == [1] * 32
    again = sle.als(op, x0, rhs, repeats=1)
    assert all(np.array_equal(a, b) for a, b in zip(one.cores, again.cores))


def test_mals_matrix_free_two_site(dev):
    """MALS where the two-site micro matrix cannot be formed (n = 64: r n n r'' = 65 536 unknowns, a 32 GiB matrix): the
    matrix-free two-site path must reduce the residual sweep over sweep and respect max_rank (sle.py:603-650)."""
    rng = np.random.default_rng(7)
    d, n, r = 4, 64, 4
    S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1)) * np.sqrt(1e-3)
    first = tt.build_core([[S, D, 1]])
    mid = tt.build_core([[1, 0, 0], [D, 0, 0], [S, D, 1]])
    last = tt.build_core([[1], [D], [S]])
    op = TT([first] + [mid.copy() for _ in range(d - 2)] + [last])
    rhs = TT([rng.standard_normal((1, n, 1, 1)) for _ in range(d)])
    x0 = TT(ott.ortho_right([rng.standard_normal((1 if i == 0 else r, n, 1, 1 if i == d - 1 else r)) for i in range(d)]))
    sols = [sle.mals(op, x0, rhs, repeats=k, threshold=1e-12, max_rank=6) for k in (1, 2)]
    res = [osle.residual(op.cores, s.cores, rhs.cores) for s in sols]
    assert res[1] <= res[0] * (1 + 1e-9) and res[1] < 1e-2
    assert max(sols[1].ranks) <= 6 and sols[1].ranks[0] == sols[1].ranks[-1] == 1


def test_trapezoidal_rule_and_adaptive_step_size(dev):
    """The steppers next to implicit_euler (ode.py:366-450, :487-636) against the live reference's
    states on the signaling cascade.  Chained steps amplify rounding differences, hence the per-step
    tolerance schedule; the accepted time grid of the step-size control must be the reference's."""
    z, zc = load("ode_steppers"), load("euler_cascade")
    op = TT(cascade_operator(zc))
    iv, guess = T(zc, "iv"), T(zc, "guess")
    sol = ode.trapezoidal_rule(op, iv, guess, [0.5, 1.0, 0.5], repeats=2, progress=False)
    assert sol[0] is iv and len(sol) == 4
    for k in range(1, 4):
        assert rel_diff(sol[k].cores, cores(z, f"trap/step{k}")) < 10 ** (k - 1) * SOL_TOL, k
    for method in ("two_step_Euler", "trapezoidal_rule"):
        ref_times = z[f"adapt/{method}/times"]
        sol, times = ode.adaptive_step_size(op, iv, guess, 2.0, step_size_first=0.1, repeats=2, second_method=method,
                                            progress=False)
        assert len(times) == len(ref_times) and np.allclose(times, ref_times, rtol=1e-6, atol=0)
        assert rel_diff(sol[1].cores, cores(z, f"adapt/{method}/step1")) < SOL_TOL
        assert rel_diff(sol[-1].cores, cores(z, f"adapt/{method}/step{len(ref_times) - 1}")) < 1e-5


def test_transfer_paths_and_deferred_redo(dev, monkeypatch):
    """Host-side plumbing of sle.als at a size where it is active (results >= 1 MiB: streamed out of the last backward half
    sweep into one page-locked block; page-locked inputs DMA'd in place; micro-solve outcomes inspected once per half
    sweep): pageable and page-locked inputs give bit-identical cores, the result cores are ordinary writable numpy arrays
    that can be fed back, the inputs are untouched, and a missed deferred outcome redoes the sweep synchronously with the
    same result."""
    import torch
    import bench
    opc, rhsc, x0c = bench.workload_cores(6, 64, 32)
    op, rhs = TT(opc), TT(rhsc)
    x0 = TT(ott.ortho_right(x0c))
    before = [c.copy() for c in x0.cores]
    a = sle.als(op, x0, rhs, repeats=2)
    assert all(np.array_equal(c, b) for c, b in zip(x0.cores, before))
    assert all(isinstance(c, np.ndarray) and c.flags.writeable and c.dtype == np.float64 for c in a.cores)
    assert torch.from_numpy(a.cores[2].reshape(-1)).is_pinned()
    opp, rhsp, x0p = TT(opc).pin_memory(), TT(rhsc).pin_memory(), TT([c.copy() for c in x0.cores]).pin_memory()
    assert torch.from_numpy(x0p.cores[2].reshape(-1)).is_pinned()
    b = sle.als(opp, x0p, rhsp, repeats=2)
    assert all(np.array_equal(p, q) for p, q in zip(a.cores, b.cores))
    c = sle.als(op, a, rhs, repeats=1)                        # a solver result fed back as the next guess
    bnorm = np.prod([np.linalg.norm(v) for v in rhsc])
    assert tt.residual_error(op, c, rhs) / bnorm <= 1.0001 * tt.residual_error(op, a, rhs) / bnorm + 1e-13
    a.cores[0][...] = 0.0                                     # results own their memory as far as the caller can tell
    assert np.any(b.cores[0] != 0.0)
    # a deferred outcome that is judged bad: the sweeps are redone with a host decision after every solve
    monkeypatch.setattr(_local.Deferred, "check", lambda self: False)
    sle._DEFER_MISSES.clear()
    redo = sle.als(op, x0, rhs, repeats=2)
    sle._DEFER_MISSES.clear()
    assert all(np.array_equal(p, q) for p, q in zip(redo.cores, b.cores))


def test_krylov_workspace_right_rank_above_64(dev):
    """Matrix-free micro systems whose right solution rank exceeds the 64 (+4) columns of the tiled vector layout (config 4:
    r = 128 / 256): the Krylov workspace must be sized by the natural length r n r' -- regression for an undersized bound
    that made the r = 256 sweep fault.  CG and GMRES solve the same micro systems, so their sweeps must agree."""
    import bench
    rng = np.random.default_rng(5)
    d, n = 4, 16
    opc = bench.laplace_cores(d, n)
    ranks = [1, 16, 80, 16, 1]
    rhs = TT([rng.standard_normal((1, n, 1, 1)) for _ in range(d)])
    x0 = TT(ott.ortho_right([rng.standard_normal((ranks[i], n, 1, ranks[i + 1])) for i in range(d)]))
    op = TT(opc)
    a = sle.als(op, x0, rhs, repeats=1, solver='cg')
    b = sle.als(op, x0, rhs, repeats=1, solver='gmres')
    assert a.ranks == ranks
    assert rel_diff(a.cores, b.cores) < SOL_TOL
    bnorm = np.prod([np.linalg.norm(c) for c in rhs.cores])
    assert tt.residual_error(op, a, rhs) / bnorm < 1e-2
