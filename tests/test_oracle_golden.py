"""CPU: the oracle (oracle/*.py) against the golden vectors produced by the live reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA
path with the oracle and with the same golden vectors."""
import numpy as np
import pytest

from oracle import kernels as K
from oracle import sle as osle, evp as oevp, ode as oode, tt as ott
from util import load, cores, rel_diff, rel_diff_up_to_phase, cascade_operator

TOL = 1e-10


def x3(c):
    return c[:, :, 0, :]


@pytest.mark.parametrize("tag", ["real", "cplx"])
def test_kernels_match_reference_private_functions(tag):
    z = load("kernels")
    op, x, b = cores(z, f"{tag}/op"), cores(z, f"{tag}/x"), cores(z, f"{tag}/b")
    d = len(op)

    def close(a, ref):
        ref = np.asarray(ref)
        assert a.shape == ref.shape
        assert np.linalg.norm(a - ref) <= 1e-13 * max(1.0, np.linalg.norm(ref))

    for i in range(1, d):
        close(K.stack_left_op(z[f"{tag}/Lop/{i-1}"], x3(x[i - 1]), op[i - 1]), z[f"{tag}/Lop/{i}"])
        close(K.stack_left_op(z[f"{tag}/evpL/{i-1}"], x3(x[i - 1]), op[i - 1], conj_col=True), z[f"{tag}/evpL/{i}"])
        close(K.stack_left_rhs(z[f"{tag}/Lrhs/{i-1}"], x3(b[i - 1]), x3(x[i - 1])), z[f"{tag}/Lrhs/{i}"])
    for i in range(d - 2, -1, -1):
        close(K.stack_right_op(z[f"{tag}/Rop/{i+1}"], x3(x[i + 1]), op[i + 1]), z[f"{tag}/Rop/{i}"])
        close(K.stack_right_rhs(z[f"{tag}/Rrhs/{i+1}"], x3(b[i + 1]), x3(x[i + 1])), z[f"{tag}/Rrhs/{i}"])
    for i in range(d):
        L, R = z[f"{tag}/Lop/{i}"], z[f"{tag}/Rop/{i}"]
        M = K.micro_matrix_als(L, op[i], R)
        close(M, z[f"{tag}/M1/{i}"])
        close(K.micro_rhs_als(z[f"{tag}/Lrhs/{i}"], x3(b[i]), z[f"{tag}/Rrhs/{i}"]).reshape(-1, 1), z[f"{tag}/f1/{i}"])
        v = x3(x[i])
        close(K.micro_matvec_als(L, op[i], R, v).reshape(-1), M @ v.reshape(-1))
    for i in range(d - 1):
        L, R = z[f"{tag}/Lop/{i}"], z[f"{tag}/Rop/{i+1}"]
        M = K.micro_matrix_mals(L, op[i], op[i + 1], R)
        close(M, z[f"{tag}/M2/{i}"])
        close(K.micro_rhs_mals(z[f"{tag}/Lrhs/{i}"], x3(b[i]), x3(b[i + 1]), z[f"{tag}/Rrhs/{i+1}"]).reshape(-1, 1),
              z[f"{tag}/f2/{i}"])
        v = np.einsum('ane,ejf->anjf', x3(x[i]), x3(x[i + 1]))
        close(K.micro_matvec_mals(L, op[i], op[i + 1], R, v).reshape(-1), M @ v.reshape(-1))


def test_sle_toeplitz():
    z = load("sle_toeplitz")
    op, rhs, x0 = cores(z, "op"), cores(z, "rhs"), cores(z, "x0")
    for solver in ("solve", "lu"):
        assert rel_diff(osle.als(op, x0, rhs, repeats=1, solver=solver), cores(z, f"als_{solver}")) < 1e-9
        m = osle.mals(op, x0, rhs, repeats=1, solver=solver, threshold=1e-14, max_rank=10)
        assert ott.ranks_of(m) == ott.ranks_of(cores(z, f"mals_{solver}"))
        assert rel_diff(m, cores(z, f"mals_{solver}")) < 1e-9
    sol = osle.als(op, x0, rhs)
    # the reference's own acceptance test (tests/test_sle.py:44-59), tolerance 1e-7
    dense = ott.matricize(sol).reshape(-1)
    assert np.linalg.norm(dense - z["dense_solution"]) / np.linalg.norm(z["dense_solution"]) < 1e-7
    assert abs(osle.residual(op, sol, rhs) - float(z["als_solve_residual"])) < 1e-9


@pytest.mark.parametrize("name", ["sle_laplace", "sle_random_spd"])
def test_sle_als_mals_spd(name):
    z = load(name)
    op, rhs, x0 = cores(z, "op"), cores(z, "rhs"), cores(z, "x0")
    r = ott.ranks_of(x0)[1:-1]
    reps = [1, 2] if name == "sle_laplace" else [2]
    for rep in reps:
        sol = osle.als(op, x0, rhs, repeats=rep)
        assert rel_diff(sol, cores(z, f"als_rep{rep}")) < TOL
        res_ref = float(z[f"als_rep{rep}_residual"])
        assert abs(osle.residual(op, sol, rhs) - res_ref) <= 1e-10 * max(res_ref, 1e-300) + 1e-13
    m = osle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=max(r))
    assert ott.ranks_of(m) == ott.ranks_of(cores(z, "mals"))
    assert rel_diff(m, cores(z, "mals")) < TOL


def test_sle_complex():
    z = load("sle_complex")
    op, rhs, x0 = cores(z, "op"), cores(z, "rhs"), cores(z, "x0")
    assert rel_diff(osle.als(op, x0, rhs, repeats=2), cores(z, "als_rep2")) < TOL
    assert rel_diff(osle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=3), cores(z, "mals")) < TOL


def test_implicit_euler():
    z = load("euler_cascade")
    op = cascade_operator(z)
    sol = oode.implicit_euler(op, cores(z, "iv"), cores(z, "guess"), [1.0] * 3)
    for k in range(1, 4):
        assert rel_diff(sol[k], cores(z, f"als/step{k}")) < 1e-9
    gen, iv2, g2 = cores(z, "gen"), cores(z, "iv2"), cores(z, "guess2")
    for p in (1, 2, 0):
        sol = oode.implicit_euler(gen, iv2, g2, [0.1, 0.2, 0.1], repeats=2, tt_solver='mals', max_rank=3, normalize=p)
        for k in range(1, 4):
            assert rel_diff(sol[k], cores(z, f"mals_norm{p}/step{k}")) < 1e-9


def test_evp_laplace():
    z = load("evp_laplace")
    op, x0 = cores(z, "op"), cores(z, "x0")
    lam, x, it = oevp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    assert abs(lam - float(z["eigh/lam"])) < 1e-10 * abs(float(z["eigh/lam"]))
    assert it == int(z["eigh/it"])
    assert rel_diff_up_to_phase(x, cores(z, "eigh/x")) < 1e-8
    lam, x, it = oevp.als(op, x0, repeats=4, conv_eps=0, solver='eig', sigma=0.0)
    assert abs(lam - float(z["eig/lam"])) < 1e-10 * max(abs(float(z["eig/lam"])), 1.0)
    assert rel_diff_up_to_phase(x, cores(z, "eig/x")) < 1e-7
    lam, x, it = oevp.als(op, x0, repeats=3, conv_eps=0, solver='eigh', number_ev=2)
    assert np.allclose(lam, z["eigh2/lam"], rtol=1e-10, atol=0)
    gev = cores(z, "gevp")
    lam, x, it = oevp.als(op, x0, op_gevp=gev, repeats=3, conv_eps=0, solver='eigh')
    assert abs(lam - float(z["gevp_eigh/lam"])) < 1e-10 * abs(float(z["gevp_eigh/lam"]))
    lam1, x1, _ = oevp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    lam2, x2, _ = oevp.als(op, x0, previous=[x1], shift=-lam1, repeats=4, conv_eps=0, solver='eigh')
    assert abs(lam2 - float(z["defl/lam2"])) < 1e-9 * abs(float(z["defl/lam2"]))


def test_evp_cooxidation_eig():
    z = load("evp_cooxidation")
    op = cores(z, "op_raw")
    opI = ott.add(ott.eye([c.shape[1] for c in op]), op)
    lam, x, it = oevp.als(opI, cores(z, "x0"), repeats=5, conv_eps=0, solver='eig', sigma=1)
    # reference-vs-reference reproducibility floor on this operator is ~1e-5
    assert abs(lam - float(z["eig/lam"])) < 1e-5 * max(1.0, abs(float(z["eig/lam"])))


def test_ortho():
    z = load("ortho")
    t = cores(z, "t")
    for key, fn in (("left", lambda c: ott.ortho_left(c)), ("right", lambda c: ott.ortho_right(c)),
                    ("left_thr", lambda c: ott.ortho_left(c, threshold=0.2)),
                    ("right_mr", lambda c: ott.ortho_right(c, max_rank=2)),
                    ("ortho", lambda c: ott.ortho_right(ott.ortho_left(c, threshold=1e-12), threshold=1e-12, max_rank=3))):
        got = fn(ott.copy_cores(t))
        ref = cores(z, key)
        assert ott.ranks_of(got) == ott.ranks_of(ref)
        assert rel_diff(got, ref) < 1e-12
    assert abs(ott.norm(t) - float(z["norm2"])) < 1e-12 * float(z["norm2"])
    assert abs(ott.norm([np.abs(c) for c in t], p=1) - float(z["norm1"])) < 1e-12 * float(z["norm1"])
    got = ott.ortho_right(ott.copy_cores(cores(z, "ones")), threshold=1e-10)
    assert ott.ranks_of(got) == ott.ranks_of(cores(z, "ones_right"))
    assert rel_diff(got, cores(z, "ones_right")) < 1e-12
    tc = cores(z, "tc")
    assert rel_diff(ott.ortho_left(ott.copy_cores(tc)), cores(z, "tc_left")) < 1e-12
    assert rel_diff(ott.ortho_right(ott.copy_cores(tc)), cores(z, "tc_right")) < 1e-12


def test_trapezoidal_rule_oracle_matches_reference():
    """ode.trapezoidal_rule (ode.py:366-450) on the signaling cascade: the oracle against the live reference's steps."""
    from oracle import ode as oode
    z, zc = load("ode_steppers"), load("euler_cascade")
    op = cascade_operator(zc)
    sol = oode.trapezoidal_rule(op, cores(zc, "iv"), cores(zc, "guess"), [0.5, 1.0, 0.5], repeats=2)
    for k in range(1, 4):
        assert rel_diff(sol[k], cores(z, f"trap/step{k}")) < 1e-9, k
