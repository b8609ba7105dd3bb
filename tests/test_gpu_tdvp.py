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
    for m, c in ((1, 0.3), (7, -0.5j), (20, 2.0 - 1.0j), (48, -3.0j), (96, 0.7)):
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
