"""Einsum restatements of the per-micro-step contractions.

Index letters: a/c = left solution rank (column/row side), e/g = right solution rank,
b/d = operator ranks, n/m = column/row mode, p/q = right-hand-side ranks.
Cores: x [r, n, r2] (col_dims squeezed), A [R, m, n, R2], b [p, m, p2].
"""
import numpy as np


def stack_left_op(L, x, A, conj_col=False):
    """sle.py:217-219 (conj on the row-side copy); evp.py:281-283 when conj_col=True."""
    xn, xm = (np.conj(x), x) if conj_col else (x, np.conj(x))
    return np.einsum('abc,ane,bmnd,cmg->edg', L, xn, A, xm, optimize=True)


def stack_right_op(Rt, x, A):
    """sle.py:274-276 and evp.py:323-325."""
    return np.einsum('ane,bmnd,cmg,edg->abc', x, A, np.conj(x), Rt, optimize=True)


def stack_left_rhs(bL, b, x):
    """sle.py:246-247; evp.py:291-292."""
    return np.einsum('pc,pmq,cmg->qg', bL, b, np.conj(x), optimize=True)


def stack_right_rhs(bR, b, x):
    """sle.py:303-305; evp.py:333-334."""
    return np.einsum('pmq,cmg,qg->pc', b, np.conj(x), bR, optimize=True)


def micro_matrix_als(L, A, Rt):
    """sle.py:339-345: rows (c,m,g), columns (a,n,e)."""
    r, r2 = L.shape[0], Rt.shape[0]
    m, n = A.shape[1], A.shape[2]
    M = np.einsum('abc,bmnd,edg->cmgane', L, A, Rt, optimize=True)
    return M.reshape(r * m * r2, r * n * r2)


def micro_matvec_als(L, A, Rt, v):
    """Matrix-free product with the matrix of micro_matrix_als (SURVEY.md a4')."""
    return np.einsum('abc,ane,bmnd,edg->cmg', L, v, A, Rt, optimize=True)


def micro_matrix_mals(L, A1, A2, Rt):
    """sle.py:381-388: rows (c,m,m2,g), columns (a,n,n2,e)."""
    r, r3 = L.shape[0], Rt.shape[0]
    m, n, m2, n2 = A1.shape[1], A1.shape[2], A2.shape[1], A2.shape[2]
    M = np.einsum('abc,bmnd,dkjf,efg->cmkganje', L, A1, A2, Rt, optimize=True)
    return M.reshape(r * m * m2 * r3, r * n * n2 * r3)


def micro_matvec_mals(L, A1, A2, Rt, v):
    """Two-site matrix-free product; v [r, n, n2, r3]."""
    return np.einsum('abc,anje,bmnd,dkjf,efg->cmkg', L, v, A1, A2, Rt, optimize=True)


def micro_rhs_als(bL, b, bR):
    """sle.py:424-428 (returned unflattened [r, m, r2])."""
    return np.einsum('pc,pmq,qg->cmg', bL, b, bR, optimize=True)


def micro_rhs_mals(bL, b1, b2, bR):
    """sle.py:464-470 (returned unflattened [r, m, m2, r3])."""
    return np.einsum('pc,pmq,qks,sg->cmkg', bL, b1, b2, bR, optimize=True)
