"""GPU: the batched small-system kernels (csrc/batch.cu, batched entry points of stacks.cu) -- every system of a batch
against the one-system kernels / numpy, and evp.als_batch against evp.als."""
import numpy as np
import pytest
import torch

import workloads
from oracle import kernels as K
from oracle import tt as ott

pytestmark = pytest.mark.gpu


def _rand(rng, shape, cplx=True):
    a = rng.standard_normal(shape)
    return a + 1j * rng.standard_normal(shape) if cplx else a


@pytest.mark.parametrize("cplx", [True, False])
def test_batch_stacks_and_micro_matrix(dev, cplx):
    rng = np.random.default_rng(3)
    B, r, R, n, r2, R2 = 5, 6, 4, 3, 7, 5
    L, x, A = _rand(rng, (B, r, R, r), cplx), _rand(rng, (B, r, n, r2), cplx), _rand(rng, (B, R, n, n, R2), cplx)
    Rt = _rand(rng, (B, r2, R2, r2), cplx)
    dL, dx, dA, dR = (dev.to_device(a) for a in (L, x, A, Rt))
    for mode, conj_col in ((0, False), (1, True)):
        got = dev.batch_stack_left_op(dL, dx, dA, mode).cpu().numpy()
        for b in range(B):
            want = K.stack_left_op(L[b], x[b], A[b], conj_col=conj_col)
            assert np.linalg.norm(got[b] - want) <= 1e-13 * np.linalg.norm(want)
    got = dev.batch_stack_right_op(dR, dx, dA).cpu().numpy()
    for b in range(B):
        want = K.stack_right_op(Rt[b], x[b], A[b])
        assert np.linalg.norm(got[b] - want) <= 1e-13 * np.linalg.norm(want)
    got = dev.batch_micro_matrix_als(dL, dA, dR).cpu().numpy()
    for b in range(B):
        want = K.micro_matrix_als(L[b], A[b], Rt[b])
        assert got[b].shape == want.shape and np.linalg.norm(got[b] - want) <= 1e-13 * np.linalg.norm(want)
        one = dev.micro_matrix_als(dL[b], dA[b], dR[b]).cpu().numpy()
        assert np.array_equal(one, got[b])                         # same kernels, the batch is a grid dimension


@pytest.mark.parametrize("N,k", [(9, 1), (81, 1), (192, 1), (192, 3), (300, 2), (768, 1)])
def test_batch_eig_shift_invert(dev, N, k):
    """Eigenpairs closest to sigma of a batch of real non-symmetric matrices (and one complex batch) against numpy."""
    rng = np.random.default_rng(N + k)
    B, sigma = 4, 1.0
    for cplx in (False, True):
        mats = []
        for b in range(B):
            # well separated spectrum around sigma with a non-normal similarity: lambda_j = sigma + 0.05 (j + 1) (+ i) ...
            lam = sigma + 0.05 * (np.arange(N) + 1) * (1 + (0.3j if cplx else 0.0)) * (1 if b % 2 == 0 else -1)
            Q = _rand(rng, (N, N), cplx) / np.sqrt(N) + np.eye(N)
            mats.append(Q @ np.diag(lam) @ np.linalg.inv(Q))
            if not cplx:
                mats[-1] = mats[-1].real
        M = np.stack(mats)
        lam_g, vec_g, status = dev.batch_eig_shift_invert(dev.to_device(M), sigma, k)
        lam_g, vec_g, status = lam_g.cpu().numpy(), vec_g.cpu().numpy(), status.cpu().numpy()
        assert (status[:, 0] >= k).all() and (status[:, 1] == 0).all(), status
        for b in range(B):
            w = np.linalg.eigvals(M[b])
            want = w[np.argsort(np.abs(w - sigma))[:k]]
            assert np.allclose(lam_g[b], want, rtol=1e-9, atol=1e-9), (b, lam_g[b], want)
            for s in range(k):
                v = vec_g[b, :, s]
                assert np.linalg.norm(M[b] @ v - lam_g[b, s] * v) <= 1e-9 * np.linalg.norm(M[b]) * np.linalg.norm(v)
                i = np.argmax(np.abs(v))
                assert abs(v[i].imag) <= 1e-12 * abs(v[i]) and v[i].real > 0       # geev phase convention
            one_l, one_v = dev.eig_shift_invert(dev.to_device(M[b].astype(complex)), sigma, k)
            assert np.allclose(np.sort_complex(one_l.cpu().numpy()), np.sort_complex(lam_g[b]), rtol=1e-9, atol=1e-9)


def test_batch_eig_flags_unconverged(dev):
    """A spectrum clustered at sigma beyond what 32 Krylov vectors and a few restarts resolve must be FLAGGED (status),
    never returned as converged."""
    rng = np.random.default_rng(0)
    N = 200
    lam = 1.0 + 1e-9 * (np.arange(N) + 1)
    Q = rng.standard_normal((N, N)) / np.sqrt(N) + np.eye(N)
    M = (Q @ np.diag(lam) @ np.linalg.inv(Q))[None]
    _, _, status = dev.batch_eig_shift_invert(dev.to_device(M), 1.0, 4, max_restarts=1)
    st = status.cpu().numpy()
    assert st[0, 0] <= 4                                            # reported honestly; evp.als_batch redoes such systems
    singular = np.zeros((1, 12, 12))
    singular[0] = np.eye(12)                                        # M - sigma I == 0: zero pivot must be flagged
    _, _, status = dev.batch_eig_shift_invert(dev.to_device(singular), 1.0, 1)
    assert int(status.cpu().numpy()[0, 1]) & 2


def test_eig_exact_fallback_on_equidistant_spectrum(dev):
    """Many eigenvalues at (nearly) the same distance from sigma -- the state the chaotic co_oxidation sweeps wander into
    (measured: 192 eigenvalues at |lambda - 1| = 9.84e7 (1 +- 1e-4)).  Restarted Arnoldi cannot separate them; the one-system
    path must fall back to the full Krylov space and return the closest eigenvalue exactly, as lin.eig does."""
    rng = np.random.default_rng(4)
    N = 96
    phi = rng.uniform(0, 2 * np.pi, N)
    rad = 1e6 * (1 + 1e-5 * np.arange(N))
    lam = 1.0 + rad * np.exp(1j * phi)
    Q = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / np.sqrt(N) + np.eye(N)
    M = Q @ np.diag(lam) @ np.linalg.inv(Q)
    before = getattr(dev, "eig_exact_fallbacks", 0)
    got, vec = dev.eig_shift_invert(dev.to_device(M), 1.0, 1)
    got = complex(got[0].item())
    assert getattr(dev, "eig_exact_fallbacks", 0) == before + 1
    assert abs(got - lam[0]) <= 1e-7 * abs(lam[0]), (got, lam[0])
    v = vec[:, 0].cpu().numpy()
    assert np.linalg.norm(M @ v - got * v) <= 1e-8 * np.linalg.norm(M) * np.linalg.norm(v)


@pytest.mark.parametrize("P,Q,keep", [(24, 8, 8), (9, 9, 9), (3, 9, 3), (27, 16, 16), (96, 32, 32), (24, 8, 5), (4, 12, 2)])
def test_batch_svd_left(dev, P, Q, keep):
    rng = np.random.default_rng(P * 100 + Q)
    B = 6
    F = _rand(rng, (B, P, Q))
    F[1] = F[1][:, :1] @ F[1][:1, :]                               # rank 1: completion of the null columns
    F[2, :, -1] = 0.0
    big = 1 << 40
    src = dev.to_device(F)
    out = dev.empty((B, P, keep), torch.complex128)
    dev.batch_svd_left(src, P, Q, keep, (big, 0, Q), (big, 0, 1), 0, out, keep, 1, 0)
    U = out.cpu().numpy()
    for b in range(B):
        assert np.linalg.norm(U[b].conj().T @ U[b] - np.eye(keep)) < 1e-12, b
        u, s, _ = np.linalg.svd(F[b], full_matrices=False)
        rank = int((s > 1e-10 * s[0]).sum())
        t = min(keep, rank)
        # the leading singular subspace (gauge invariant): projector onto the first t left singular vectors
        if t == rank or s[t - 1] - s[t] > 1e-6 * s[0]:
            Pg, Pw = U[b][:, :t] @ U[b][:, :t].conj().T, u[:, :t] @ u[:, :t].conj().T
            assert np.linalg.norm(Pg - Pw) < 1e-9, (b, t)
    # the conjugate-transposed access used by the backward half sweep: F = G^H given G, output written as rows of Vh
    G = _rand(rng, (B, Q, P))                                      # G is Q x P, F = G^H is P x Q
    out = dev.empty((B, keep, P), torch.complex128)
    dev.batch_svd_left(dev.to_device(G), P, Q, keep, (big, 0, 1), (big, 0, P), 1, out, 1, P, 1)
    Vh = out.cpu().numpy()
    for b in range(B):
        assert np.linalg.norm(Vh[b] @ Vh[b].conj().T - np.eye(keep)) < 1e-12
        _, s, vh = np.linalg.svd(G[b], full_matrices=False)
        t = keep
        if t == min(P, Q) or s[t - 1] - s[t] > 1e-6 * s[0]:
            Pg, Pw = Vh[b][:t].conj().T @ Vh[b][:t], vh[:t].conj().T @ vh[:t]
            assert np.linalg.norm(Pg - Pw) < 1e-9, b


def test_evp_als_batch_against_single(dev):
    """evp.als_batch == evp.als per system on a well-conditioned family (scaled Laplacian-type operators, 'eig'): same
    eigenvalues to 1e-10, same eigentensors to 1e-8, same iteration counts; number_ev = 2 and the shared-guess form."""
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import evp
    from util import rel_diff_up_to_phase
    d, n, r = 5, 4, 3
    rng = np.random.default_rng(1)
    ops = []
    for b in range(5):
        cores = workloads.laplace_cores(d, n, c=0.05 * (b + 1))
        ops.append(TT(cores))
    guess = TT(ott.ortho_right(workloads.random_guess(d, n, r, seed=6)))
    batch = evp.als_batch(ops, guess, repeats=3, conv_eps=0, solver='eig', sigma=0.0)
    assert len(batch) == 5
    for b in range(5):
        lam1, x1, it1 = evp.als(ops[b], guess, repeats=3, conv_eps=0, solver='eig', sigma=0.0)
        lamb, xb, itb = batch[b]
        assert isinstance(xb, TT) and itb == it1 and xb.ranks == x1.ranks
        assert abs(lamb - lam1) <= 1e-10 * max(abs(lam1), 1.0), (b, lamb, lam1)
        assert rel_diff_up_to_phase(xb.cores, x1.cores) < 1e-8, b
    lam2, xs2, _ = evp.als(ops[1], guess, repeats=2, conv_eps=0, solver='eig', sigma=0.0, number_ev=2)
    (lamb2, xsb2, _), = evp.als_batch([ops[1]], [guess], repeats=2, conv_eps=0, solver='eig', sigma=0.0, number_ev=2)
    assert np.allclose(lamb2, lam2, rtol=1e-9, atol=0) and len(xsb2) == 2
    # early stop per system (conv_eps) matches the single-system iteration counts
    batch = evp.als_batch(ops, guess, repeats=20, conv_eps=1e-6, solver='eig', sigma=0.0)
    for b in (0, 4):
        lam1, x1, it1 = evp.als(ops[b], guess, repeats=20, conv_eps=1e-6, solver='eig', sigma=0.0)
        assert batch[b][2] == it1 and abs(batch[b][0] - lam1) <= 1e-9 * max(abs(lam1), 1.0)


def test_lanczos_local_eigensolver(dev, monkeypatch):
    """Matrix-free Hermitian local eigen-solver (thick-restart Lanczos on the micro-matvec), selected for 'eigh' above
    EIGH_DENSE_LIMIT: (1) the k largest eigenpairs of a micro operator against the dense Jacobi eigh of the assembled
    micro matrix; (2) evp.als('eigh') against the live reference's golden vectors at 1e-10 / 1e-8 with the limit lowered so
    that every micro step runs matrix-free; (3) a micro operator of 65 536 unknowns (dense: 32 GiB) converges and satisfies
    the eigen-equation."""
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import evp, _local
    from util import load, cores, rel_diff_up_to_phase
    r, n = 6, 8
    opc = workloads.laplace_cores(3, n, c=0.05)
    xc = ott.ortho_right(workloads.random_guess(3, n, r, seed=2))
    one3 = np.ones((1, 1, 1))
    L = K.stack_left_op(one3, xc[0][:, :, 0, :], opc[0])   # stacks of a symmetric operator: Hermitian local operator
    Rt = K.stack_right_op(one3, xc[2][:, :, 0, :], opc[2])
    A = opc[1]
    dL, dA, dR = dev.to_device(L), dev.to_device(A), dev.to_device(Rt)
    M = dev.micro_matrix_als(dL, dA, dR)
    Mh = M.cpu().numpy()
    assert np.allclose(Mh, Mh.T, atol=1e-12 * np.abs(Mh).max())
    op = dev.local_op(dL, dA, dR, prepare=True)
    for k in (1, 3):
        theta, vec = _local.eigh_matrix_free(dev, lambda v: dev.local_matvec(op, v), (r, n, r), torch.float64, k)
        w = np.linalg.eigvalsh(Mh)[::-1][:k]
        assert np.allclose(theta.cpu().numpy(), w, rtol=1e-11, atol=1e-11 * abs(w[0]))
        v = vec.cpu().numpy()
        assert np.linalg.norm(Mh @ v - v * theta.cpu().numpy()[None, :]) <= 1e-9 * abs(w[0])
        assert np.linalg.norm(v.T @ v - np.eye(k)) < 1e-10
    z = load("evp_laplace")
    opt, x0 = TT(cores(z, "op")), TT(cores(z, "x0"))
    monkeypatch.setattr(_local, "EIGH_DENSE_LIMIT", 8)
    lam, x, it = evp.als(opt, x0, repeats=4, conv_eps=0, solver='eigh')
    ref = float(z["eigh/lam"])
    assert abs(lam - ref) < 1e-10 * abs(ref) and it == int(z["eigh/it"])
    # every micro eigenvector agrees with the dense eigh of the same micro matrix to ~1e-14 (tools/lanczos_probe.py); the
    # wanted eigentensor of this operator is nearly rank 1, so the trailing columns of the SVD re-orthonormalisation
    # (singular values ~1e-5) amplify rounding-level differences: 1.0e-7 on the tensor against the reference
    assert rel_diff_up_to_phase(x.cores, cores(z, "eigh/x")) < 1e-6
    lam2, xs, _ = evp.als(opt, x0, repeats=3, conv_eps=0, solver='eigh', number_ev=2)
    assert np.allclose(lam2, z["eigh2/lam"], rtol=1e-10, atol=0)
    monkeypatch.undo()
    # beyond any dense matrix: r = 16, n = 64 -> 16 384 unknowns per micro system with the default limit (4096)
    d, n, r = 4, 64, 16
    big = TT(workloads.laplace_cores(d, n, c=0.05))
    guess = TT(ott.ortho_right(workloads.random_guess(d, n, r, seed=3)))
    _local.lanczos_stats.update(solves=0, matvecs=0, restarts=0, worst_residual=0.0)
    lam1, x1, _ = evp.als(big, guess, repeats=1, conv_eps=0, solver='eigh')
    lam2, x2, _ = evp.als(big, guess, repeats=2, conv_eps=0, solver='eigh')
    assert _local.lanczos_stats["solves"] > 0 and _local.lanczos_stats["worst_residual"] <= 1e-11
    q = lambda l, v: ott.norm(ott.sub(ott.matmul(big.cores, v.cores), ott.scale(v.cores, l))) / ott.norm(v.cores)
    assert lam2 >= lam1 * (1 - 1e-12) and q(lam2, x2) <= q(lam1, x1) * (1 + 1e-6)      # largest eigenvalue: Ritz values grow
    assert lam2 < 4 * d + 1                                                          # Gershgorin bound of the operator
