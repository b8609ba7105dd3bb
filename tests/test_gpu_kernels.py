"""GPU: every entry point of the C-ABI (include/sktt_b200.h) against the numpy oracle on seeded inputs.

Tolerances are relative Frobenius-norm errors in fp64; the contraction kernels are held to 1e-13,
the factorisations to 1e-11 x a mild condition-number allowance written next to each check.
"""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import kernels as K

pytestmark = pytest.mark.gpu


def rnd(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return a


def relerr(a, ref):
    ref = np.asarray(ref)
    a = np.asarray(a)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return np.linalg.norm((a - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300)


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------ gemm
GEMM_SHAPES = [(1, 1, 1), (3, 5, 7), (17, 33, 65), (64, 64, 64), (100, 130, 50), (192, 4096, 64), (4096, 192, 192),
               (40, 24, 3000), (129, 257, 31)]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm2_plain_and_strided(dev, cplx, mode):
    rng = np.random.default_rng(100 + mode + 10 * cplx)
    dev.set_gemm_mode(mode)
    try:
        for (M, N, Kd) in GEMM_SHAPES:
            A = rnd(rng, (M, Kd), cplx)
            B = rnd(rng, (Kd, N), cplx)
            C0 = rnd(rng, (M, N), cplx)
            dA, dB, dC = dev.to_device(A), dev.to_device(B), dev.to_device(C0)
            big = 1 << 40
            # plain row-major, alpha/beta
            dev.gemm2(M, N, Kd, dA, (big, 0, Kd), (big, 0, 1), dB, (big, 0, N), (big, 0, 1), dC, (big, 0, N), (big, 0, 1),
                      alpha=(0.5, 0.0), beta=(2.0, 0.0))
            assert relerr(host(dC), 0.5 * A @ B + 2.0 * C0) < 1e-13, (M, N, Kd)
            # transposed operands through the strides + conjugation flags
            dAt, dBt = dev.to_device(A.T.copy()), dev.to_device(B.T.copy())
            dC2 = dev.empty((M, N), dA.dtype)
            dev.gemm2(M, N, Kd, dAt, (big, 0, 1), (big, 0, M), dBt, (big, 0, 1), (big, 0, Kd), dC2, (big, 0, N), (big, 0, 1),
                      conjA=1, conjB=1)
            assert relerr(host(dC2), np.conj(A) @ np.conj(B)) < 1e-13, (M, N, Kd)
        # two-level index maps: C[i1,j1,i2,j2] = sum_{k1,k2} A[k1,i1,k2,i2] B[j2,k2,k1,j1]
        i1, i2, j1, j2, k1, k2 = 5, 7, 3, 6, 4, 9
        A = rnd(rng, (k1, i1, k2, i2), cplx)
        B = rnd(rng, (j2, k2, k1, j1), cplx)
        dA, dB = dev.to_device(A), dev.to_device(B)
        dC = dev.empty((i1, j1, i2, j2), dA.dtype)
        dev.gemm2(i1 * i2, j1 * j2, k1 * k2, dA, (i2, k2 * i2, 1), (k2, i1 * k2 * i2, i2),
                  dB, (k2, j1, k1 * j1), (j2, 1, k2 * k1 * j1), dC, (i2, j1 * i2 * j2, j2), (j2, i2 * j2, 1))
        ref = np.einsum('aibj,dbac->icjd', A, B)
        assert relerr(host(dC), ref) < 1e-13
    finally:
        dev.set_gemm_mode(0)


# ------------------------------------------------------------------------------------------------ stacks
STACK_SHAPES = [  # r, R, n, r2, R2
    (1, 1, 3, 2, 2), (3, 2, 4, 5, 3), (4, 4, 64, 4, 4), (8, 21, 3, 8, 21), (16, 3, 16, 16, 3), (64, 3, 64, 64, 3),
    (7, 5, 9, 1, 1), (32, 3, 64, 64, 3), (64, 3, 64, 32, 3), (8, 2, 32, 64, 3), (64, 3, 96, 4, 5)]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("pp", [(1, 1), (4, 4), (1, 4), (5, 6), (2, 7)])
@pytest.mark.parametrize("shape", [(1, 3, 2), (5, 9, 1), (64, 64, 64), (33, 16, 70)])
def test_rhs_stacks_small_and_general_rhs_ranks(dev, cplx, pp, shape):
    """Right-hand-side stacks and micro right-hand sides (sle.py:222-247, :279-305, :393-430): rhs ranks <= 4 run the
    one-kernel forms, larger ones the two-GEMM chains; both against the einsum restatement, and twice for bit-identity
    (the left stack sums CTA partials in a fixed order)."""
    r, n, r2 = shape
    p, p2 = pp
    rng = np.random.default_rng(11 + r + 7 * p + 13 * p2 + 101 * cplx)
    x = rnd(rng, (r, n, r2), cplx)
    bL, bR, b = rnd(rng, (p, r), cplx), rnd(rng, (p2, r2), cplx), rnd(rng, (p, n, p2), cplx)
    dx, dbL, dbR, db = map(dev.to_device, (x, bL, bR, b))
    tol = 2e-13
    first = host(dev.stack_left_rhs(dbL, db, dx))
    assert relerr(first, K.stack_left_rhs(bL, b, x)) < tol
    assert np.array_equal(first, host(dev.stack_left_rhs(dbL, db, dx)))
    assert relerr(host(dev.stack_right_rhs(dbR, db, dx)), K.stack_right_rhs(bR, b, x)) < tol
    assert relerr(host(dev.micro_rhs_als(dbL, db, dbR)), K.micro_rhs_als(bL, b, bR)) < tol


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", STACK_SHAPES)
def test_stacks_and_micro_systems(dev, cplx, shape):
    r, R, n, r2, R2 = shape
    rng = np.random.default_rng(7 + r + 31 * R + 101 * cplx)
    L, Rt = rnd(rng, (r, R, r), cplx), rnd(rng, (r2, R2, r2), cplx)
    x, A = rnd(rng, (r, n, r2), cplx), rnd(rng, (R, n, n, R2), cplx)
    p, p2 = 2, 3
    bL, bR, b = rnd(rng, (p, r), cplx), rnd(rng, (p2, r2), cplx), rnd(rng, (p, n, p2), cplx)
    dL, dR, dx, dA = map(dev.to_device, (L, Rt, x, A))
    dbL, dbR, db = map(dev.to_device, (bL, bR, b))
    tol = 2e-13
    assert relerr(host(dev.stack_left_op(dL, dx, dA, 0)), K.stack_left_op(L, x, A)) < tol
    assert relerr(host(dev.stack_left_op(dL, dx, dA, 1)), K.stack_left_op(L, x, A, conj_col=True)) < tol
    assert relerr(host(dev.stack_right_op(dR, dx, dA)), K.stack_right_op(Rt, x, A)) < tol
    assert relerr(host(dev.stack_left_rhs(dbL, db, dx)), K.stack_left_rhs(bL, b, x)) < tol
    assert relerr(host(dev.stack_right_rhs(dbR, db, dx)), K.stack_right_rhs(bR, b, x)) < tol
    assert relerr(host(dev.micro_rhs_als(dbL, db, dbR)), K.micro_rhs_als(bL, b, bR)) < tol
    v = rnd(rng, (r, n, r2), cplx)
    yref = K.micro_matvec_als(L, A, Rt, v)
    assert relerr(host(dev.micro_matvec_als(dL, dA, dR, dev.to_device(v))), yref) < tol
    if r * n * r2 <= 2048:
        Mref = K.micro_matrix_als(L, A, Rt)
        assert relerr(host(dev.micro_matrix_als(dL, dA, dR)), Mref) < tol
        assert relerr((Mref @ v.reshape(-1)).reshape(yref.shape), yref) < 1e-12


@pytest.mark.parametrize("shape", [(64, 3, 64, 64, 3), (32, 3, 64, 64, 3), (8, 2, 32, 64, 3), (4, 5, 96, 64, 3),
                                   (128, 3, 32, 64, 3), (20, 1, 64, 64, 3),
                                   # padded: edge cores (r = 1 / r2 = 1), ranks that are not multiples of 4, small R2
                                   (1, 1, 64, 64, 3), (64, 3, 64, 1, 1), (3, 2, 32, 5, 2), (64, 3, 64, 16, 3), (6, 4, 32, 64, 1)])
def test_fused_matvec_path_matches_generic_chain_and_oracle(dev, shape):
    """Shapes covered by the two-kernel fused matvec (fused.cu): against the oracle and against the generic chain."""
    r, R, n, r2, R2 = shape
    rng = np.random.default_rng(5 + r + n)
    L, Rt = rnd(rng, (r, R, r), False), rnd(rng, (r2, R2, r2), False)
    v, A = rnd(rng, (r, n, r2), False), rnd(rng, (R, n, n, R2), False)
    dL, dR, dv, dA = map(dev.to_device, (L, Rt, v, A))
    yref = K.micro_matvec_als(L, A, Rt, v)
    l0 = dev.launches()
    y_fused = host(dev.micro_matvec_als(dL, dA, dR, dv))
    assert dev.launches() - l0 == 5                      # stateless: image build, to-tiled, 2 fused kernels, from-tiled
    op = dev.local_op(dL, dA, dR, prepare=True)
    assert op.image is not None
    y_prepared = host(dev.local_matvec(op, dv))
    assert relerr(y_prepared, yref) < 2e-13
    # the Krylov inner step: vectors in the tiled layout [n][a][r2 + 4], two kernel launches
    # (rows padded to a multiple of 4, columns to 64 + 4; the padding is zero on input and stays zero on output)
    rp = (r + 3) // 4 * 4
    vt = np.zeros((n, rp, 68))
    vt[:, :r, :r2] = v.transpose(1, 0, 2)
    l0 = dev.launches()
    yt = host(dev.local_matvec_tiled(op, dev.to_device(vt.reshape(-1)))).reshape(n, rp, 68)
    assert dev.launches() - l0 == 2
    assert relerr(yt[:, :r, :r2].transpose(1, 0, 2), yref) < 2e-13
    assert np.all(yt[:, :, r2:] == 0.0) and np.all(yt[:, r:, :] == 0.0)
    dev.set_gemm_mode(1)
    try:
        y_generic = host(dev.micro_matvec_als(dL, dA, dR, dv))
    finally:
        dev.set_gemm_mode(0)
    assert relerr(y_fused, yref) < 2e-13
    assert relerr(y_generic, yref) < 2e-13


@pytest.mark.parametrize("cplx", [False, True])
def test_two_site_micro_systems(dev, cplx):
    rng = np.random.default_rng(23 + cplx)
    for (r, R, n, R2, n2, R3, r3) in [(2, 2, 3, 3, 4, 2, 3), (1, 1, 2, 3, 2, 1, 1), (4, 3, 5, 3, 5, 3, 4),
                                      (6, 2, 16, 2, 16, 2, 6)]:
        L, Rt = rnd(rng, (r, R, r), cplx), rnd(rng, (r3, R3, r3), cplx)
        A1, A2 = rnd(rng, (R, n, n, R2), cplx), rnd(rng, (R2, n2, n2, R3), cplx)
        v = rnd(rng, (r, n, n2, r3), cplx)
        p, p2, p3 = 2, 3, 2
        bL, bR = rnd(rng, (p, r), cplx), rnd(rng, (p3, r3), cplx)
        b1, b2 = rnd(rng, (p, n, p2), cplx), rnd(rng, (p2, n2, p3), cplx)
        dL, dR, dA1, dA2, dv = map(dev.to_device, (L, Rt, A1, A2, v))
        yref = K.micro_matvec_mals(L, A1, A2, Rt, v)
        assert relerr(host(dev.micro_matvec_mals(dL, dA1, dA2, dR, dv)), yref) < 2e-13
        assert relerr(host(dev.micro_rhs_mals(*map(dev.to_device, (bL, b1, b2, bR)))), K.micro_rhs_mals(bL, b1, b2, bR)) < 2e-13
        if r * n * n2 * r3 <= 2400:
            Mref = K.micro_matrix_mals(L, A1, A2, Rt)
            assert relerr(host(dev.micro_matrix_mals(dL, dA1, dA2, dR)), Mref) < 2e-13


@pytest.mark.parametrize("cplx", [False, True])
def test_rank1_update(dev, cplx):
    rng = np.random.default_rng(5)
    N = 75
    M, t = rnd(rng, (N, N), cplx), rnd(rng, (N,), cplx)
    dM = dev.to_device(M)
    dev.rank1_update(dM, dev.to_device(t), -0.7)
    assert relerr(host(dM), M - 0.7 * np.outer(t, np.conj(t))) < 1e-14


# ------------------------------------------------------------------------------------------------ dense solves
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 100, 256, 257, 600, 1024])
def test_lu_solve_matches_lapack(dev, cplx, N):
    rng = np.random.default_rng(N + 1000 * cplx)
    M = rnd(rng, (N, N), cplx)
    f = rnd(rng, (N,), cplx)
    xref = np.linalg.solve(M, f)
    dM = dev.to_device(M)
    ipiv, info = dev.lu_factor(dM)
    assert info == 0
    # same pivot sequence as LAPACK getrf (partial pivoting, first maximal entry) on generic input
    _, piv_ref = sla.lu_factor(M)
    assert np.array_equal(host(ipiv)[:N], piv_ref)
    dx = dev.lu_solve(dM, ipiv, dev.to_device(f))
    cond = np.linalg.cond(M)
    assert relerr(host(dx), xref) < 1e-14 * cond + 1e-13
    # several right-hand sides
    F = rnd(rng, (N, 3), cplx)
    dX = dev.lu_solve(dM, ipiv, dev.to_device(F))
    assert relerr(host(dX), np.linalg.solve(M, F)) < 1e-14 * cond + 1e-13


def test_lu_reports_singular(dev):
    M = np.ones((40, 40))
    M[:, 7] = 0.0
    with pytest.raises(np.linalg.LinAlgError):
        dev.solve(dev.to_device(M), dev.to_device(np.ones(40)))


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("N", [1, 5, 32, 70, 300, 1000])
def test_cholesky(dev, cplx, N):
    rng = np.random.default_rng(N + 7 * cplx)
    G = rnd(rng, (N, N + 3), cplx)
    M = G @ np.conj(G.T) + N * np.eye(N)
    F = rnd(rng, (N, 2), cplx)
    dM = dev.to_device(M)
    assert dev.chol_factor(dM) == 0
    assert relerr(np.tril(host(dM)), np.linalg.cholesky(M)) < 1e-12
    dX = dev.chol_solve(dM, dev.to_device(F))
    assert relerr(host(dX), np.linalg.solve(M, F)) < 1e-12
    Mbad = M.copy()
    Mbad[N // 2, N // 2] = -1.0
    assert dev.chol_factor(dev.to_device(Mbad)) != 0


# ------------------------------------------------------------------------------------------------ QR / RQ / SVD
QR_SHAPES = [(256, 4), (16, 16), (5, 9), (1, 6), (7, 1), (1000, 64), (4096, 64), (40, 130), (300, 200)]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", QR_SHAPES)
def test_qr_rq_match_lapack(dev, cplx, shape):
    """Q, R against scipy.linalg.qr / rq.  The economic Q of a full-rank matrix is unique up to the sign of each column
    (row for RQ); tall unfoldings go through the CholeskyQR kernel (R with a positive diagonal), the others through the
    Householder kernel (LAPACK's signs), so the comparison fixes the signs from the diagonals of R."""
    m, n = shape
    rng = np.random.default_rng(m * 131 + n + cplx)
    A = rnd(rng, (m, n), cplx)
    Q, R = dev.qr(dev.to_device(A), want_r=True)
    Q, R = host(Q), host(R)
    k = min(m, n)
    assert relerr(Q @ R, A) < 1e-13
    assert relerr(np.conj(Q.T) @ Q, np.eye(k)) < 1e-13
    assert np.allclose(R, np.triu(R))
    qref, rref = sla.qr(A, mode='economic')
    sg = np.sign(np.real(np.diag(R))) * np.sign(np.real(np.diag(rref)))
    assert np.max(np.abs(np.imag(np.diag(R)))) < 1e-12 * np.abs(R).max()
    assert relerr(Q * sg, qref) < 1e-11 and relerr(sg[:, None] * R, rref) < 1e-11
    # RQ on the transposed shape (the backward ALS step factors r x (n r2))
    B = rnd(rng, (n, m), cplx)
    Rr, Qr = dev.rq(dev.to_device(B), want_r=True)
    Rr, Qr = host(Rr), host(Qr)
    assert relerr(Rr @ Qr, B) < 1e-13
    assert relerr(Qr @ np.conj(Qr.T), np.eye(k)) < 1e-13
    rref, qref = sla.rq(B, mode='economic')
    sg = np.sign(np.real(np.diag(Rr[-k:, :] if Rr.shape[0] > k else Rr))) * \
        np.sign(np.real(np.diag(rref[-k:, :] if rref.shape[0] > k else rref)))
    assert relerr(sg[:, None] * Qr, qref) < 1e-11 and relerr(Rr * sg, rref) < 1e-11


QR_HARD = [(4096, 64, 1e6), (4096, 64, 1e12), (1024, 48, 1e15), (512, 96, 1e10), (4096, 64, np.inf), (256, 4, np.inf)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", QR_HARD)
def test_qr_ill_conditioned_and_rank_deficient(dev, case):
    """CholeskyQR passes / shift / Householder fallback: orthonormal Q and a backward-stable factorisation whatever the
    conditioning (inf = exactly rank-deficient: duplicated and zero columns)."""
    m, n, cond = case
    rng = np.random.default_rng(m + n)
    U0, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    if np.isinf(cond):
        A = rng.standard_normal((m, n))
        A[:, 1] = A[:, 0]
        A[:, -1] = 0.0
    else:
        A = (U0 * np.logspace(0, -np.log10(cond), n)) @ V0.T
    Q, R = dev.qr(dev.to_device(A), want_r=True)
    Q, R = host(Q), host(R)
    assert np.all(np.isfinite(Q)) and np.all(np.isfinite(R))
    assert relerr(Q.T @ Q, np.eye(n)) < 1e-13
    assert relerr(Q @ R, A) < 1e-13
    assert np.allclose(R, np.triu(R))
    Rr, Qr = dev.rq(dev.to_device(np.ascontiguousarray(A.T)), want_r=True)
    Rr, Qr = host(Rr), host(Qr)
    assert relerr(Qr @ Qr.T, np.eye(n)) < 1e-13
    assert relerr(Rr @ Qr, A.T) < 1e-13


SVD_SHAPES = [(1, 1), (6, 4), (4, 6), (64, 64), (256, 12), (12, 256), (1000, 64), (192, 192), (4096, 64), (300, 260)]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", SVD_SHAPES)
def test_svd_values_vectors_and_rank_rule(dev, cplx, shape):
    m, n = shape
    rng = np.random.default_rng(m * 17 + n + cplx)
    k = min(m, n)
    # prescribed spectrum spanning 12 decades so that the strict threshold rule is exercised
    U0, _ = np.linalg.qr(rnd(rng, (m, k), cplx))
    V0, _ = np.linalg.qr(rnd(rng, (n, k), cplx))
    s0 = np.logspace(0, -12, k) if k > 1 else np.ones(1)
    A = (U0 * s0) @ np.conj(V0.T)
    sref = sla.svd(A, compute_uv=False, lapack_driver='gesvd')
    thr = 1e-6
    U, S, Vh, rank = dev.svd(dev.to_device(A), threshold=thr, max_rank=0)
    U, S, Vh = host(U), host(S), host(Vh)
    assert np.max(np.abs(S - sref) / np.maximum(sref, 1e-16 * sref[0])) < 1e-3       # relative, even for tiny s
    assert np.max(np.abs(S - sref)) < 1e-14 * sref[0] * max(m, n)
    assert relerr((U * S) @ Vh, A) < 5e-13
    assert relerr(np.conj(U.T) @ U, np.eye(k)) < 1e-12
    assert relerr(Vh @ np.conj(Vh.T), np.eye(k)) < 1e-12
    assert rank == int(np.sum(sref / sref[0] > thr))
    _, _, _, rank2 = dev.svd(dev.to_device(A), threshold=0.0, max_rank=3)
    assert rank2 == min(3, k)
    _, _, _, rank3 = dev.svd(dev.to_device(A), threshold=0.0, max_rank=0)
    assert rank3 == k


def test_svd_rank_deficient_completion(dev):
    # tt.ones(..., ranks=4) unfoldings are numerically rank 1: U must still be orthonormal
    A = np.ones((12, 4))
    U, S, Vh, rank = dev.svd(dev.to_device(A), threshold=0.0, max_rank=0)
    U, S, Vh = host(U), host(S), host(Vh)
    assert rank == 4
    assert abs(S[0] - np.sqrt(48.0)) < 1e-13 and np.all(S[1:] < 1e-14)
    assert relerr(U.T @ U, np.eye(4)) < 1e-13
    assert relerr((U * S) @ Vh, A) < 1e-13


# ------------------------------------------------------------------------------------------------ eigen
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("N", [1, 2, 24, 150, 400])
def test_eigh(dev, cplx, N):
    rng = np.random.default_rng(N + cplx)
    G = rnd(rng, (N, N), cplx)
    M = G + np.conj(G.T)
    wref = sla.eigh(M, eigvals_only=True)
    W, V = dev.eigh(dev.to_device(M))
    W, V = host(W), host(V)
    scale = max(np.abs(wref).max(), 1e-300)
    assert np.max(np.abs(W - wref)) < 1e-13 * scale * max(N, 8)
    assert relerr(M @ V, V * W) < 1e-12
    assert relerr(np.conj(V.T) @ V, np.eye(N)) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("N", [6, 60, 300])
def test_eig_shift_invert(dev, cplx, N):
    rng = np.random.default_rng(N + 3 * cplx)
    M = rnd(rng, (N, N), cplx) / np.sqrt(N) + np.diag(np.linspace(0.0, 3.0, N))
    sigma, k = 1.234, min(3, N - 1)
    spectrum = sla.eigvals(M)
    dref = np.sort(np.abs(spectrum - sigma))[:k]
    lam, vecs = dev.eig_shift_invert(dev.to_device(M), sigma, k, ncv=min(N, 40))
    lam, vecs = host(lam), host(vecs)
    order = np.argsort(np.abs(lam - sigma))
    lam, vecs = lam[order], vecs[:, order]
    # the k closest to sigma (conjugate pairs tie in |lambda - sigma|: compare distances, then membership)
    assert np.max(np.abs(np.abs(lam - sigma) - dref)) < 1e-10
    assert all(np.min(np.abs(spectrum - l)) < 1e-10 for l in lam)
    for j in range(k):
        v = vecs[:, j]
        assert np.linalg.norm(M @ v - lam[j] * v) < 1e-9 * np.linalg.norm(v)
    # generalised pencil
    Bm = np.eye(N) + 0.1 * (lambda g: g @ np.conj(g.T))(rnd(rng, (N, N), cplx)) / N
    spectrum = sla.eigvals(M, Bm)
    dref = np.sort(np.abs(spectrum - sigma))[:k]
    lam, vecs = dev.eig_shift_invert(dev.to_device(M), sigma, k, B=dev.to_device(Bm), ncv=min(N, 40))
    lam = host(lam)
    assert np.max(np.abs(np.sort(np.abs(lam - sigma)) - dref)) < 1e-9
    assert all(np.min(np.abs(spectrum - l)) < 1e-9 for l in lam)


# ------------------------------------------------------------------------------------------------ Krylov
def spd_local(rng, r, R, n, r2, cplx):
    """A Hermitian positive definite one-site local operator: L, Rt Hermitian PSD per operator index, A symmetric."""
    def psd(k):
        g = rnd(rng, (k, k), cplx)
        return g @ np.conj(g.T) / k + np.eye(k)
    L = np.stack([psd(r) for _ in range(R)], axis=1)           # [a, b, c], Hermitian in (a, c) -> conj layout below
    Rt = np.stack([psd(r2) for _ in range(R)], axis=1)
    A = np.zeros((R, n, n, R), dtype=L.dtype)
    for b in range(R):
        g = rnd(rng, (n, n), cplx)
        A[b, :, :, b] = np.eye(n) + 0.05 * (g + np.conj(g.T))
    # M[(c,m,g),(a,n,e)] = sum_b L[a,b,c] A[b,m,n,b] Rt[e,b,g]: Hermitian if L[a,b,c] = conj(L[c,b,a]) etc.
    return L.transpose(2, 1, 0).copy(), A, Rt.transpose(2, 1, 0).copy()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("method", ["cg", "gmres"])
def test_krylov_one_site(dev, cplx, method):
    rng = np.random.default_rng(77 + cplx)
    r, R, n, r2 = 5, 3, 7, 4
    L, A, Rt = spd_local(rng, r, R, n, r2, cplx)
    M = K.micro_matrix_als(L, A, Rt)
    assert relerr(M, np.conj(M.T)) < 1e-14
    f = rnd(rng, (r, n, r2), cplx)
    uref = np.linalg.solve(M, f.reshape(-1)).reshape(f.shape)
    dL, dA, dR = map(dev.to_device, (L, A, Rt))
    op = dev.local_op(dL, dA, dR)
    u = dev.to_device(np.zeros_like(f))
    st, iters, relres = dev.krylov_solve(op, dev.to_device(f), u, method=method, tol=1e-13, max_iters=2000, restart=30)
    assert st == 0 and relres <= 1e-13
    assert relerr(host(u), uref) < 1e-11


def test_krylov_two_site_and_nonsymmetric(dev):
    rng = np.random.default_rng(99)
    r, R, n, R2, n2, R3, r3 = 3, 2, 4, 2, 5, 2, 3
    L, Rt = rnd(rng, (r, R, r), False), rnd(rng, (r3, R3, r3), False)
    A1, A2 = rnd(rng, (R, n, n, R2), False), rnd(rng, (R2, n2, n2, R3), False)
    M = K.micro_matrix_mals(L, A1, A2, Rt)
    f = rnd(rng, (r, n, n2, r3), False)
    uref = np.linalg.solve(M, f.reshape(-1)).reshape(f.shape)
    dL, dA1, dA2, dR = map(dev.to_device, (L, A1, A2, Rt))
    op = dev.local_op(dL, dA1, dR, A2=dA2)
    u = dev.to_device(np.zeros_like(f))
    N = f.size
    st, iters, relres = dev.krylov_solve(op, dev.to_device(f), u, method="gmres", tol=1e-12, max_iters=4 * N, restart=N)
    assert st == 0
    assert relerr(host(u), uref) < 1e-9 * np.linalg.cond(M)


def test_blas1_helpers(dev):
    rng = np.random.default_rng(3)
    for cplx in (False, True):
        x = rnd(rng, (100003,), cplx)
        assert abs(dev.nrm2(dev.to_device(x)) - np.linalg.norm(x)) < 1e-12 * np.linalg.norm(x)


def test_refined_cg_ill_conditioned_uses_the_mode_preconditioner(dev):
    """One micro system at the bench shape (r = 64, n = 64, R = 3) with converged-sweep-like stacks: spread spectra in the
    accumulated left / right operators, the 64-point Laplacian in the mode factor (condition number ~ 1e4).  The one-call
    solve must switch to the mode-preconditioned CG and reach a TRUE relative residual <= 1e-12, checked with the
    einsum restatement of the matvec."""
    import torch
    rng = np.random.default_rng(5)
    r = n = 64
    S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = np.sqrt(1e-3) * 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1))
    I = np.eye(n)
    A = np.zeros((3, n, n, 3))
    A[0, :, :, 0], A[1, :, :, 0], A[2, :, :, 0], A[2, :, :, 1], A[2, :, :, 2] = I, D, S, D, I

    def spd(lo, hi):
        q, _ = np.linalg.qr(rng.standard_normal((r, r)))
        return (q * np.geomspace(lo, hi, r)) @ q.T

    def skew(scale):
        x = rng.standard_normal((r, r))
        return scale * (x - x.T)
    L = np.stack([spd(1e-2, 30.0), skew(1e-3), np.eye(r)], axis=1)          # [a, b, c]
    Rt = np.stack([np.eye(r), skew(1e-3), spd(1e-2, 30.0)], axis=1)
    f = rng.standard_normal((r, n, r))
    dL, dA, dR, df = (dev.to_device(x) for x in (L, A, Rt, f))
    op = dev.local_op(dL, dA, dR, prepare=True)
    assert dev.tiled_len(op) > 0
    u = torch.zeros(f.size, dtype=torch.float64, device=dev.device)
    # poison the reusable Krylov workspace: nothing the solver reads may depend on what an earlier call left there (the
    # padding columns of the tiled vectors once did)
    import ctypes
    nwork = dev.lib.sktt_krylov_work(ctypes.byref(op), 0, 0)
    dev.work(nwork, torch.float64, tag="krylov").fill_(1e30)
    st, iters, relres, cycles = dev.krylov_solve_refined(op, df, u, tol=1e-13, max_iters=4000, max_cycles=5)
    assert st == 0 and relres <= 1e-12, (st, relres, iters)
    uh = u.cpu().numpy().reshape(f.shape)
    true = np.linalg.norm(f - K.micro_matvec_als(L, A, Rt, uh)) / np.linalg.norm(f)
    assert true <= 1e-12, true
    assert 40 < iters < 1500, iters                       # more than the unpreconditioned budget -> the preconditioned leg ran
    dev.set_debug(4)                                      # same solve without the preconditioner needs far more iterations
    try:
        u2 = torch.zeros_like(u)
        st2, iters2, relres2, _ = dev.krylov_solve_refined(op, df, u2, tol=1e-13, max_iters=20000, max_cycles=5)
    finally:
        dev.set_debug(0)
    assert st2 == 0 and relres2 <= 1e-12 and iters2 > iters, (iters, iters2)
    assert np.linalg.norm(u2.cpu().numpy() - u.cpu().numpy()) <= 1e-9 * np.linalg.norm(uh)


# ------------------------------------------------------------------------------------------------ full BASELINE sizes
@pytest.mark.parametrize("r", [128, 256])
def test_c4_full_size_contractions(dev, r):
    """BASELINE config 4 at full size (d=10 interior core: n=16, R=8, r=128 / 256): stack updates and the micro-matvec
    against the einsum restatement (SURVEY.md 8c parity plan, item 1).  The reference cannot run this size at all."""
    R, n = 8, 16
    rng = np.random.default_rng(r)
    L, Rt = rnd(rng, (r, R, r), False), rnd(rng, (r, R, r), False)
    x, A = rnd(rng, (r, n, r), False), rnd(rng, (R, n, n, R), False)
    dL, dR, dx, dA = (dev.to_device(a) for a in (L, Rt, x, A))
    assert relerr(host(dev.stack_left_op(dL, dx, dA)), K.stack_left_op(L, x, A)) < 1e-12
    assert relerr(host(dev.stack_right_op(dR, dx, dA)), K.stack_right_op(Rt, x, A)) < 1e-12
    assert relerr(host(dev.micro_matvec_als(dL, dA, dR, dx)), K.micro_matvec_als(L, A, Rt, x)) < 1e-12


@pytest.mark.parametrize("shape", [(4096, 64, 64), (4096, 64, 37), (100, 5, 3), (33, 1, 1), (20000, 64, 64)])
def test_gauge_products(dev, shape):
    """Q^H u / u Q^H and the push of that factor into the neighbouring core (warm starts of the matrix-free micro solves;
    the factor the reference discards at sle.py:525, :541), both storage orders, against numpy; bit-identical reruns."""
    L, k, r = shape
    rng = np.random.default_rng(L + 7 * k + r)
    q, u = rng.standard_normal((L, k)), rng.standard_normal((L, r))
    dq, du = dev.to_device(q), dev.to_device(u)
    R = host(dev.gauge_factor(dq, du, tall=True))
    assert relerr(R, q.T @ u) < 1e-13
    assert np.array_equal(R, host(dev.gauge_factor(dq, du, tall=True)))
    dqt, dut = dev.to_device(np.ascontiguousarray(q.T)), dev.to_device(np.ascontiguousarray(u.T))
    assert relerr(host(dev.gauge_factor(dqt, dut, tall=False)), u.T @ q) < 1e-13
    carry = rng.standard_normal((k, r))
    core = rng.standard_normal((r, L))
    assert relerr(host(dev.gauge_push(dev.to_device(carry), dev.to_device(core), left=True)), carry @ core) < 1e-13
    carry2 = rng.standard_normal((r, k))
    core2 = rng.standard_normal((L, r))
    assert relerr(host(dev.gauge_push(dev.to_device(carry2), dev.to_device(core2), left=False)), core2 @ carry2) < 1e-13


@pytest.mark.parametrize("n", [32, 64, 96])
@pytest.mark.parametrize("sparse", [False, True])
def test_stack_update_natural_layout_kernel(dev, n, sparse):
    """The one-launch interface-stack update that reads L, x and A where they lie (csrc/stack_nat.cu; solution ranks 64,
    operator ranks 3, mode size 32 or 64): left update and mirrored (right) update against the einsum restatement of sle.py:217-219 / :274-276
    and against the image-based kernel it replaces (debug bit 64), dense and block-sparse operator cores (the mask of zero
    rank blocks is found by the kernel itself and cleared again: a dense core right after a sparse one must not inherit
    it), bit-identical reruns."""
    r, R = 64, 3
    rng = np.random.default_rng(1000 + n + sparse)
    L, Rt = rnd(rng, (r, R, r), False), rnd(rng, (r, R, r), False)
    x = rnd(rng, (r, n, r), False)
    A = rnd(rng, (R, n, n, R), False)
    if sparse:                                            # the Laplacian-type pattern of the bench operator: 5 of 9 blocks
        for (b, q) in ((0, 1), (0, 2), (1, 1), (1, 2)):
            A[b, :, :, q] = 0.0
    dL, dR, dx, dA = (dev.to_device(a) for a in (L, Rt, x, A))
    l0 = dev.launches()
    got_l = host(dev.stack_left_op(dL, dx, dA))
    if n <= 64:                                           # mode sizes above 64 keep the image-based kernel
        assert dev.launches() - l0 == 1                   # one kernel: no image build, no tiling pass, no memset
    got_r = host(dev.stack_right_op(dR, dx, dA))
    assert relerr(got_l, K.stack_left_op(L, x, A)) < 1e-13
    assert relerr(got_r, K.stack_right_op(Rt, x, A)) < 1e-13
    assert np.array_equal(got_l, host(dev.stack_left_op(dL, dx, dA)))
    assert np.array_equal(got_r, host(dev.stack_right_op(dR, dx, dA)))
    dev.set_debug(64)
    try:
        old_l, old_r = host(dev.stack_left_op(dL, dx, dA)), host(dev.stack_right_op(dR, dx, dA))
    finally:
        dev.set_debug(0)
    assert relerr(got_l, old_l) < 1e-13 and relerr(got_r, old_r) < 1e-13
    # a dense core after the sparse one: the block mask of the previous launch is gone
    A2 = rnd(rng, (R, n, n, R), False)
    dA2 = dev.to_device(A2)
    assert relerr(host(dev.stack_left_op(dL, dx, dA2)), K.stack_left_op(L, x, A2)) < 1e-13
    assert relerr(host(dev.stack_right_op(dR, dx, dA2)), K.stack_right_op(Rt, x, A2)) < 1e-13


def test_qr_deferred_mode(dev):
    """Tall QR / RQ in the deferred mode the sweeps use (sktt_ctx_set_qr_deferred): the sketched CholeskyQR runs alone -- one
    launch, no Householder kernel behind its failure flag -- and a failure is reported through the sticky word instead."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((4096, 64))
    dA = dev.to_device(A)
    ref = host(dev.qr(dA))
    dev.set_qr_deferred(True)
    try:
        assert dev.qr_deferred_failures() == 0
        l0 = dev.launches()
        q = host(dev.qr(dA))
        assert dev.launches() - l0 == 1
        assert np.array_equal(q, ref)                     # the same kernel, the same bits
        assert np.linalg.norm(q.T @ q - np.eye(64)) < 1e-13
        assert dev.qr_deferred_failures() == 0
        bad = A.copy()
        bad[17, 3] = np.nan                               # nothing can be factorised here: the kernel must say so
        dev.qr(dev.to_device(bad))
        assert dev.qr_deferred_failures() == 1
        assert dev.qr_deferred_failures() == 0            # read once, cleared
    finally:
        dev.set_qr_deferred(False)


@pytest.mark.parametrize("shape", [(4096, 128), (4096, 256), (2048, 200), (1100, 96)])
def test_blocked_qr_of_wide_unfoldings(dev, shape):
    """Tall QR / RQ beyond the 64 columns of the sketched CholeskyQR kernel (BASELINE config 4: 4096 x 128 / 256): block
    Gram-Schmidt with re-orthogonalisation over 64-column blocks (scikit_tt_b200/_device.py::_qr_blocked).  Orthogonality and
    span to rounding level, equality with LAPACK's Q up to the sign of each column (row), a rank-deficient input, and the
    Householder kernel it replaces as the second reference."""
    m, n = shape
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n)) @ np.diag(np.geomspace(1.0, 1e-6, n)) @ rng.standard_normal((n, n))
    dA = dev.to_device(A)
    q = host(dev.qr(dA))
    assert np.linalg.norm(q.T @ q - np.eye(n)) < 1e-13
    assert np.linalg.norm(A - q @ (q.T @ A)) < 1e-12 * np.linalg.norm(A)
    qref = np.linalg.qr(A)[0]
    signs = np.sign(np.sum(q * qref, axis=0))
    assert np.linalg.norm(q * signs - qref) < 1e-7             # cond 1e6 x 1e6: the columns agree to eps * cond
    dev.blocked_qr = False
    try:
        qh = host(dev.qr(dA))
    finally:
        dev.blocked_qr = True
    assert np.linalg.norm(q * np.sign(np.sum(q * qh, axis=0)) - qh) < 1e-7
    # RQ of the transposed problem: orthonormal rows, same row space, LAPACK's Q up to the sign of each row
    B = np.ascontiguousarray(A.T)
    p = host(dev.rq(dev.to_device(B)))
    assert p.shape == (n, m) and np.linalg.norm(p @ p.T - np.eye(n)) < 1e-13
    assert np.linalg.norm(B - (B @ p.T) @ p) < 1e-12 * np.linalg.norm(B)
    pref = sla.rq(B, mode='economic')[1]
    assert np.linalg.norm(p * np.sign(np.sum(p * pref, axis=1))[:, None] - pref) < 1e-7
    # numerically rank-deficient (rank n - 40): still an orthonormal basis that contains the column space
    D = rng.standard_normal((m, n - 40)) @ rng.standard_normal((n - 40, n))
    qd = host(dev.qr(dev.to_device(D)))
    assert np.linalg.norm(qd.T @ qd - np.eye(n)) < 1e-12
    assert np.linalg.norm(D - qd @ (qd.T @ D)) < 1e-11 * np.linalg.norm(D)
