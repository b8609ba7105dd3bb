"""Oracle restatement of scikit_tt/solvers/evp.py:17-179, :253-495 (ALS for eigenvalue problems)."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse.linalg as spla

from . import kernels as K
from .tt import copy_cores


def _x3(core):
    return core[:, :, 0, :]


def _local_eig(M, B, k, solver, sigma, real):
    # evp.py:417-443
    if solver == 'eigs':
        lam, vec = spla.eigs(M, M=B, sigma=sigma, k=k, v0=np.ones(M.shape[0]))
        idx = np.abs(lam - sigma).argsort()[::-1]
        lam, vec = lam[idx], vec[:, idx]
    elif solver == 'eig':
        lam, vec = sla.eig(M, b=B, check_finite=False)
        idx = np.abs(lam - sigma).argsort()
        lam, vec = lam[idx[:k]], vec[:, idx[:k]]
    elif solver == 'eigh':
        n = M.shape[0]
        lam, vec = sla.eigh(M, b=B, check_finite=False, subset_by_index=(n - k, n - 1))
        lam, vec = lam[::-1], vec[:, ::-1]
    else:
        raise ValueError(solver)
    if real:
        lam = np.real(lam)
    return lam, vec


def als(op, x0, previous=(), shift=0, op_gevp=None, number_ev=1, repeats=1, conv_eps=1e-10, solver='eig', sigma=1,
        real=True, conj_fix=False):
    """evp.py:17-179 on core lists.  Returns (eigenvalues, list of core lists, iterations).

    conj_fix=False reproduces the reference's left-stack conjugation (evp.py:281-283, conj on the
    column-side core); conj_fix=True uses the Hermitian-consistent sle.py:217-219 convention.
    """
    x = copy_cores(x0)
    d = len(op)
    k = number_ev
    one3, one2 = np.ones((1, 1, 1)), np.ones((1, 1))
    Lop, Rop = [None] * d, [None] * d
    Lg, Rg = [None] * d, [None] * d
    Lp = [[None] * d for _ in previous]
    Rp = [[None] * d for _ in previous]

    def right(i):                                                         # evp.py:295-334
        if i == d - 1:
            Rop[i] = one3
            Rg[i] = one3
            for j in range(len(previous)):
                Rp[j][i] = one2
            return
        xi = _x3(x[i + 1])
        Rop[i] = K.stack_right_op(Rop[i + 1], xi, op[i + 1])
        if op_gevp is not None:
            Rg[i] = K.stack_right_op(Rg[i + 1], xi, op_gevp[i + 1])
        for j, t in enumerate(previous):
            Rp[j][i] = K.stack_right_rhs(Rp[j][i + 1], _x3(t[i + 1]), xi)

    def left(i):                                                          # evp.py:253-292
        if i == 0:
            Lop[i] = one3
            Lg[i] = one3
            for j in range(len(previous)):
                Lp[j][i] = one2
            return
        xi = _x3(x[i - 1])
        Lop[i] = K.stack_left_op(Lop[i - 1], xi, op[i - 1], conj_col=not conj_fix)
        if op_gevp is not None:
            Lg[i] = K.stack_left_op(Lg[i - 1], xi, op_gevp[i - 1], conj_col=not conj_fix)
        for j, t in enumerate(previous):
            Lp[j][i] = K.stack_left_rhs(Lp[j][i - 1], _x3(t[i - 1]), xi)

    def micro(i):                                                         # evp.py:337-383
        M = K.micro_matrix_als(Lop[i], op[i], Rop[i])
        B = K.micro_matrix_als(Lg[i], op_gevp[i], Rg[i]) if op_gevp is not None else None
        for j, t in enumerate(previous):
            v = K.micro_rhs_als(Lp[j][i], _x3(t[i]), Rp[j][i]).reshape(-1, 1)
            M = M + shift * (v @ np.conj(v.T))
        return M, B

    def update(i, direction):                                             # evp.py:386-495
        r, n, r2 = Lop[i].shape[0], op[i].shape[2], Rop[i].shape[0]
        M, B = micro(i)
        lam, vec = _local_eig(M, B, k, solver, sigma, real)
        if direction == 'forward':
            u, _, _ = sla.svd(vec.reshape(r * n, r2 * k), check_finite=False, lapack_driver='gesvd')
            rr = min(r2, u.shape[1])
            x[i] = u[:, :rr].reshape(r, n, 1, rr)
        elif i > 0:
            _, _, v = sla.svd(vec.transpose().reshape(k * r, n * r2), check_finite=False, lapack_driver='gesvd')
            rr = min(r, v.shape[0])
            x[i] = v[:rr, :].reshape(rr, n, 1, r2)
        else:
            x[i] = vec.reshape(r, n, 1, r2, k)
        return lam

    for i in range(d - 1, -1, -1):
        right(i)
    it = 1
    pre = np.array([np.inf] * k)[None, :]
    conv = False
    lam_opt, x_opt, lam = np.inf, None, None
    while it <= repeats and not conv:
        for i in range(d):
            left(i)
            if i < d - 1:
                lam = update(i, 'forward')
        for i in range(d - 1, -1, -1):
            right(i)
            lam = update(i, 'backward')
        it += 1
        if k == 1 and np.abs(lam[0] - sigma) < np.abs(lam_opt - sigma):   # evp.py:151-155
            lam_opt = lam[0].copy()
            x_opt = [x[0][:, :, :, :, 0]] + [c.copy() for c in x[1:]]
        last = pre[-min(3, pre.shape[0]):, :]                             # evp.py:158-165
        if np.amax(np.abs(last - lam)) < conv_eps:
            conv = True
        pre = np.vstack((pre, lam))
    if k == 1:
        return lam_opt, x_opt, it - 1
    return lam, [[x[0][:, :, :, :, j]] + [c.copy() for c in x[1:]] for j in range(k)], it - 1
