"""ALS for (generalised) eigenvalue problems in TT format -- same call surface as
scikit_tt/solvers/evp.py of PGelss/scikit_tt (`als` :17-179, `power_method` :182-250), on the GPU.

Per micro-step: left / right interface stacks for the operator, the optional right-hand operator of
the pencil and the deflation vectors [evp.py:253-334] -> dense micro matrices plus the rank-one
deflation terms [:337-383] -> local eigen-solve [:417-443] -> SVD re-orthonormalisation of the
eigenvector block [:452-493].

Local eigen-solvers (all on the device):
  'eigh' : Hermitian micro matrix, the `number_ev` LARGEST eigenvalues (sigma ignored, as evp.py:434-439);
           cyclic Jacobi.  With operator_gevp the pencil is reduced through a Cholesky factor of B.
  'eig'  : general micro matrix, the `number_ev` eigenvalues closest to sigma, ascending |lambda - sigma|
           (evp.py:424-432); shift-invert Arnoldi on the LU factors of (M - sigma B).
  'eigs' : same eigenpairs, returned in DESCENDING |lambda - sigma| order as evp.py:417-422 does.

Conjugation: the left stacks conjugate the column-side copy of the solution core exactly as
evp.py:281-283 does; for real cores this is indistinguishable from the
Hermitian-consistent sle convention.  `hermitian_stacks=True` switches to the latter.
"""
import numpy as np
import torch

from .. import _device
from .. import tensor_train as tt
from ..tensor_train import TT
from . import _local, sle


def als(operator, initial_guess, previous=[], shift=0, operator_gevp=None, number_ev=1, repeats=1, conv_eps=1e-10,
        solver='eig', sigma=1, real=True, hermitian_stacks=False):
    dev = _device.get_device()
    d, k = operator.order, number_ev
    cplx = solver in ('eig', 'eigs') or _local.any_complex(operator, initial_guess, operator_gevp, *previous)
    dtype = torch.complex128 if cplx else torch.float64
    A = _local.Uploaded(dev, operator, dtype, vector=False)
    G = _local.Uploaded(dev, operator_gevp, dtype, vector=False) if operator_gevp is not None else None
    P = [_local.Uploaded(dev, t, dtype, vector=True) for t in previous]
    x = list(_local.Uploaded(dev, initial_guess, dtype, vector=True).cores)            # evp.py:90 (copy)
    one3, one2 = _local.ones(dev, (1, 1, 1), dtype), _local.ones(dev, (1, 1), dtype)
    Lop, Rop, Lg, Rg = [None] * d, [None] * d, [None] * d, [None] * d
    Lp = [[None] * d for _ in previous]
    Rp = [[None] * d for _ in previous]
    conj_mode = _device.CONJ_ROW if hermitian_stacks else _device.CONJ_COL

    def right(i):                                                                      # evp.py:295-334
        if i == d - 1:
            Rop[i], Rg[i] = one3, one3
            for j in range(len(P)):
                Rp[j][i] = one2
            return
        Rop[i] = dev.stack_right_op(Rop[i + 1], x[i + 1], A[i + 1])
        if G is not None:
            Rg[i] = dev.stack_right_op(Rg[i + 1], x[i + 1], G[i + 1])
        for j in range(len(P)):
            Rp[j][i] = dev.stack_right_rhs(Rp[j][i + 1], P[j][i + 1], x[i + 1])

    def left(i):                                                                       # evp.py:253-292
        if i == 0:
            Lop[i], Lg[i] = one3, one3
            for j in range(len(P)):
                Lp[j][i] = one2
            return
        Lop[i] = dev.stack_left_op(Lop[i - 1], x[i - 1], A[i - 1], conj_mode)
        if G is not None:
            Lg[i] = dev.stack_left_op(Lg[i - 1], x[i - 1], G[i - 1], conj_mode)
        for j in range(len(P)):
            Lp[j][i] = dev.stack_left_rhs(Lp[j][i - 1], P[j][i - 1], x[i - 1])

    def update(i, direction):                                                          # evp.py:337-495
        r, n, r2 = Lop[i].shape[0], A[i].shape[2], Rop[i].shape[0]
        M = dev.micro_matrix_als(Lop[i], A[i], Rop[i])
        B = dev.micro_matrix_als(Lg[i], G[i], Rg[i]) if G is not None else None
        for j in range(len(P)):
            t = dev.micro_rhs_als(Lp[j][i], P[j][i], Rp[j][i])
            dev.rank1_update(M, t, shift)                                              # evp.py:381
        lam, vec = _local_eig(dev, M, B, k, solver, sigma)                             # vec [N, k]
        if direction == 'forward':
            U, _, _, _ = dev.svd(vec.reshape(r * n, r2 * k))                           # evp.py:452-454
            rr = min(r2, U.shape[1])
            x[i] = U[:, :rr].contiguous().reshape(r, n, rr)
        elif i > 0:
            _, _, Vh, _ = dev.svd(vec.t().contiguous().reshape(k * r, n * r2))         # evp.py:469-477
            rr = min(r, Vh.shape[0])
            x[i] = Vh[:rr, :].contiguous().reshape(rr, n, r2)
        else:
            x[i] = vec.reshape(r, n, r2, k)                                            # evp.py:492-493
        lam = lam.detach().cpu().numpy()
        return np.real(lam) if real else lam                                           # evp.py:441-443

    for i in range(d - 1, -1, -1):                                                     # evp.py:103-104
        right(i)
    it = 1
    pre = np.array([np.inf] * k)[None, :]                                              # evp.py:110
    conv = False
    lam_opt, x_opt, lam = np.inf, None, None
    while it <= repeats and not conv:                                                  # evp.py:118
        for i in range(d):
            left(i)
            if i < d - 1:
                lam = update(i, 'forward')
        for i in range(d - 1, -1, -1):
            right(i)
            lam = update(i, 'backward')
        it += 1
        if k == 1 and np.abs(lam[0] - sigma) < np.abs(lam_opt - sigma):               # evp.py:151-155
            lam_opt = lam[0].copy()
            x_opt = [x[0][:, :, :, 0].contiguous()] + list(x[1:])
        last = pre[-min(3, pre.shape[0]):, :]                                          # evp.py:158-165
        if np.amax(np.abs(last - lam)) < conv_eps:
            conv = True
        pre = np.vstack((pre, lam))
    if k == 1:
        return lam_opt, TT(_local.download_vector_cores(x_opt)), it - 1
    tensors = [TT(_local.download_vector_cores([x[0][:, :, :, j].contiguous()] + list(x[1:]))) for j in range(k)]
    return lam, tensors, it - 1


def _local_eig(dev, M, B, k, solver, sigma):
    """The micro eigen-solve of evp.py:417-439 on the device; returns (lam [k], vec [N, k])."""
    N = M.shape[0]
    if solver == 'eigh':
        if B is None:
            W, V = dev.eigh(M)
        else:
            # pencil (M, B), B Hermitian positive definite: B = L L^H, C = L^-1 M L^-H, v = L^-H y
            if dev.chol_factor(B) != 0:
                raise np.linalg.LinAlgError("the micro matrix of operator_gevp is not positive definite")
            Y = dev.chol_trsm(B, M, backward=False)
            Z = dev.chol_trsm(B, Y.mH.contiguous(), backward=False)
            W, Yv = dev.eigh(Z)
            V = dev.chol_trsm(B, Yv, backward=True)
        kk = min(k, N)
        lam = torch.flip(W[N - kk:], dims=[0])                                        # largest first, evp.py:438-439
        vec = torch.flip(V[:, N - kk:], dims=[1]).contiguous()
        return lam, vec
    if solver in ('eig', 'eigs'):
        if M.dtype != torch.complex128:
            M = dev.widen(M)
            B = dev.widen(B) if B is not None else None
        lam, vec = dev.eig_shift_invert(M, sigma, k, B=B)
        order = torch.argsort(torch.abs(lam - sigma), descending=(solver == 'eigs'), stable=True)
        return lam[order], vec[:, order].contiguous()
    raise ValueError("solver must be 'eig', 'eigs' or 'eigh'")


def power_method(operator, initial_guess, operator_gevp=None, repeats=10, sigma=0.999):
    """Inverse power iteration on top of sle.als (evp.py:182-250)."""
    if operator_gevp is None:
        shifted = operator - sigma * tt.eye(operator.row_dims)
    else:
        shifted = operator - sigma * operator_gevp
    eigenvalue, eigentensor = 0, initial_guess
    for _ in range(repeats):
        rhs = eigentensor if operator_gevp is None else operator_gevp.dot(eigentensor)
        eigentensor = sle.als(shifted, eigentensor, rhs)
        eigentensor = eigentensor * (1 / eigentensor.norm())
        eigenvalue = eigentensor.transpose().dot(operator).dot(eigentensor)
        if operator_gevp is not None:
            eigenvalue *= 1 / (eigentensor.transpose().dot(operator_gevp).dot(eigentensor))
    return eigenvalue, eigentensor
