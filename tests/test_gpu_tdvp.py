"""GPU: ode.tdvp1site / ode.tdvp2site against states produced by the live reference
(tests/golden/make_tdvp_golden.py: Ising chain of tests/test_ode.py:135-158, real and imaginary time, exact and
local-Krylov micro solvers), and the small matrix exponential against scipy."""
import numpy as np
import pytest
import scipy.linalg as sla

from util import load, cores, rel_diff

pytestmark = pytest.mark.gpu


def _T(c):
    from scikit_tt_b200 import TT
    return TT([np.array(x) for x in c])


def test_expm_small_and_action(dev):
    import torch
    from scikit_tt_b200.solvers import ode
    rng = np.random.default_rng(5)
    for m, c in ((1, 0.3), (7, -0.5j), (20, 2.0 - 1.0j), (48, -3.0j), (64, 0.7)):
        H = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        got = dev.expm_small(dev.to_device(H), c).cpu().numpy()
        want = sla.expm(c * H)
        assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want), (m, c)
    for N, c in ((12, -0.4j), (150, -0.2j), (300, 0.05)):
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        M = 0.5 * (A + A.conj().T) / np.sqrt(N)
        if np.isreal(c):
            M = A / np.sqrt(N)                                    # the action must not rely on Hermitian structure
        v = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        got = ode._expm_action(dev, dev.to_device(M), dev.to_device(v), c).cpu().numpy()
        want = sla.expm(c * M) @ v
        assert np.linalg.norm(got - want) <= 1e-11 * np.linalg.norm(want), (N, c)


@pytest.mark.parametrize("tag,h,solver,normalize", [("real_exact", 0.05, None, 0), ("imag_exact", -1j * 0.05, None, 2),
                                                     ("real_krylov", 0.05, {"method": "local_krylov", "dimension": 4}, 0)])
def test_tdvp_against_the_reference(dev, tag, h, solver, normalize):
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import ode
    z = load("tdvp")
    op, x0 = _T(cores(z, "op")), _T(cores(z, "x0"))
    sol = ode.tdvp1site(op, x0, h, 3, local_solver=solver, normalize=normalize)
    assert sol[0] is x0 and len(sol) == 4 and all(isinstance(t, TT) for t in sol)
    for k in range(1, 4):
        ref = cores(z, f"tdvp1/{tag}/step{k}")
        assert sol[k].ranks == [c.shape[0] for c in ref] + [1]
        assert rel_diff(sol[k].cores, ref) < 1e-8, (tag, k)
    sol = ode.tdvp2site(op, x0, h, 3, local_solver=solver, threshold=1e-10, max_rank=4, normalize=normalize)
    for k in range(1, 4):
        ref = cores(z, f"tdvp2/{tag}/step{k}")
        assert sol[k].ranks == [c.shape[0] for c in ref] + [1]
        assert rel_diff(sol[k].cores, ref) < 1e-8, (tag, k)
    if normalize == 0 and solver is None:
        # unitary evolution: norm and energy are conserved by the exact one-site integrator
        from oracle import tt as ott
        nrm = [ott.norm(t.cores) for t in ode.tdvp1site(op, x0, h, 3)]
        assert max(abs(n - nrm[0]) for n in nrm) < 1e-10


def test_tt_algebra_on_device(dev, monkeypatch):
    """`@` and `tt.residual_error` through the device kernels against the host statement of the same
    formulas (tensor_train.py:422-503, :2035-2074), real and complex, and against the oracle."""
    import workloads
    import scikit_tt_b200.tensor_train as tt
    from scikit_tt_b200 import TT
    from oracle import tt as ott, sle as osle
    rng = np.random.default_rng(9)
    d, n, r = 5, 8, 6
    op = TT(workloads.laplace_cores(d, n))
    x = TT(workloads.random_guess(d, n, r, seed=2))
    b = TT(workloads.rank1_rhs(d, n))
    xc = TT([c + 1j * rng.standard_normal(c.shape) for c in x.cores])
    monkeypatch.setattr(tt, "DEVICE_ALGEBRA_MIN_WORK", 1 << 60)
    host = [op @ x, op @ xc, op @ op]
    res_host = [tt.residual_error(op, x, b), tt.residual_error(op, xc, b)]
    l0 = dev.launches()
    monkeypatch.setattr(tt, "DEVICE_ALGEBRA_MIN_WORK", 0)
    devi = [op @ x, op @ xc, op @ op]
    res_dev = [tt.residual_error(op, x, b), tt.residual_error(op, xc, b)]
    assert dev.launches() > l0
    for h, g in zip(host, devi):
        assert g.ranks == h.ranks and g.row_dims == h.row_dims and g.col_dims == h.col_dims
        for ch, cg in zip(h.cores, g.cores):
            assert cg.dtype == ch.dtype and np.linalg.norm(cg - ch) <= 1e-14 * max(np.linalg.norm(ch), 1e-300)
    for h, g in zip(res_host, res_dev):
        assert abs(g - h) <= 1e-11 * h
    want = osle.residual(op.cores, x.cores, b.cores) * ott.norm(b.cores)
    assert abs(res_dev[0] - want) <= 1e-9 * want
    assert np.isscalar(TT([np.ones((1, 1, 3, 1))]) @ TT([np.ones((1, 3, 1, 1))]))       # fully contracted product -> scalar


def test_power_method_against_the_reference(dev):
    """evp.power_method (evp.py:182-250): inverse power iteration on top of the GPU sle.als, plain and generalised, against
    the live reference (tests/golden/make_power_golden.py) -- eigenvalue to 1e-10, eigentensor to 1e-8."""
    import workloads
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import evp
    z = load("power_method")
    op = TT(workloads.laplace_cores(4, 6, c=0.05))
    x0, gev = _T(cores(z, "x0")), _T(cores(z, "gev"))
    for tag, kw in (("plain", {}), ("gevp", {"operator_gevp": gev})):
        for reps in (1, 3):
            lam, x = evp.power_method(op, x0, repeats=reps, sigma=0.3, **kw)
            ref = float(z[f"{tag}/rep{reps}/lam"])
            assert abs(lam - ref) <= 1e-10 * abs(ref), (tag, reps, lam, ref)
            assert rel_diff(x.cores, cores(z, f"{tag}/rep{reps}/x")) < 1e-8, (tag, reps)


@pytest.mark.parametrize("N", [1, 5, 16, 17, 100, 256, 1000, 1024, 1536])
def test_fused_lu_solve(dev, N):
    """np.linalg.solve in one cooperative launch (csrc/lu_fused.cu): solution, LU factors and LAPACK's pivot sequence."""
    import scipy.linalg as sla
    rng = np.random.default_rng(N)
    M = rng.standard_normal((N, N))
    if N >= 100:
        M[7, :] = M[3, :] * (1 + 1e-3 * rng.standard_normal(N))        # ill-conditioned rows: pivoting matters
    f = rng.standard_normal(N)
    dM = dev.to_device(M)
    x, piv = dev.solve_fused(dM, dev.to_device(f), want_pivots=True)
    lu_ref, piv_ref = sla.lu_factor(M)
    assert np.array_equal(piv.cpu().numpy(), piv_ref)
    got = dM.cpu().numpy()
    assert np.linalg.norm(got - lu_ref) <= 1e-11 * np.linalg.norm(lu_ref)
    want = np.linalg.solve(M, f)
    assert np.linalg.norm(x.cpu().numpy() - want) <= 1e-9 * np.linalg.norm(want)
    again_M = dev.to_device(M)
    again = dev.solve_fused(again_M, dev.to_device(f))
    assert np.array_equal(again.cpu().numpy(), x.cpu().numpy()) and np.array_equal(again_M.cpu().numpy(), got)   # bit-reproducible
    if N >= 16:
        S = M.copy()
        S[:, 9] = 0.0
        with pytest.raises(np.linalg.LinAlgError):
            dev.solve(dev.to_device(S), dev.to_device(f))


def test_arr_against_the_reference(dev):
    """data_driven.regression.arr (regression.py:15-142): alternating ridge regression with sample-indexed stacks and an
    SVD-based least-squares micro solve, against the live reference (tests/golden/make_arr_golden.py): coefficient tensors to
    1e-8, and the fit they represent reproduces the targets."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_arr_golden import basis
    from scikit_tt_b200 import TT
    from scikit_tt_b200.data_driven import regression as reg
    z = load("arr")
    x, y = z["x"], z["y"]
    d = x.shape[0]
    guess = TT([z[f"guess/{i}"] for i in range(d)])
    for reps in (1, 3):
        sol = reg.arr(x, y, basis(d), guess, repeats=reps, rcond=1e-10, progress=False)
        assert len(sol) == 2 and all(isinstance(t, TT) for t in sol)
        for k, t in enumerate(sol):
            ref = [z[f"rep{reps}/row{k}/{i}"] for i in range(d)]
            assert t.ranks == [c.shape[0] for c in ref] + [1]
            assert rel_diff(t.cores, ref) < 1e-8, (reps, k)
    # the fit the coefficient tensors represent is the reference's fit (same training error to 1e-8)
    Phi = [np.array([[f(x[:, j]) for j in range(x.shape[1])] for f in mode]) for mode in basis(d)]

    def predict(cores_):
        acc = np.ones((1, x.shape[1]))
        for c, P in zip(cores_, Phi):
            acc = np.einsum('aj,kj,akl->lj', acc, P, c[:, :, 0, :])
        return acc[0]
    for k, t in enumerate(sol):
        ref = [z[f"rep3/row{k}/{i}"] for i in range(d)]
        e_new, e_ref = np.linalg.norm(predict(t.cores) - y[k]), np.linalg.norm(predict(ref) - y[k])
        assert abs(e_new - e_ref) <= 1e-8 * np.linalg.norm(y[k]), (k, e_new, e_ref)
