"""Restatements of the per-micro-step contractions  --  oracle only.

Each function performs the same chain of pairwise np.tensordot contractions, in the same order,
as the reference line it cites, so that on the same BLAS the results are bit-identical to the
reference's private helpers (pinned by tests/test_oracle_golden.py against tests/golden/kernels.npz).

Cores: x [r, n, r2] (col_dims squeezed), A [R, m, n, R2], b [p, m, p2]; stacks are indexed
(solution, operator, conj-solution) as in the reference.
"""
import numpy as np

td = np.tensordot


def stack_left_op(L, x, A, conj_col=False):
    """sle.py:217-219 (conj on the row-side copy); evp.py:281-283 when conj_col=True."""
    first, last = (np.conj(x), x) if conj_col else (x, np.conj(x))
    t = td(L, first, axes=(0, 0))
    t = td(t, A, axes=([0, 2], [0, 2]))
    return td(t, last, axes=([0, 2], [0, 1]))


def stack_right_op(Rt, x, A):
    """sle.py:274-276 and evp.py:323-325."""
    t = td(np.conj(x), Rt, axes=(2, 2))
    t = td(A, t, axes=([1, 3], [1, 3]))
    return td(x, t, axes=([1, 2], [1, 3]))


def stack_left_rhs(bL, b, x):
    """sle.py:246-247; evp.py:291-292."""
    t = td(bL, b, axes=(0, 0))
    return td(t, np.conj(x), axes=([0, 1], [0, 1]))


def stack_right_rhs(bR, b, x):
    """sle.py:303-305; evp.py:333-334."""
    t = td(np.conj(x), bR, axes=(2, 1))
    return td(b, t, axes=([1, 2], [1, 2]))


def micro_matrix_als(L, A, Rt):
    """sle.py:339-345 / evp.py:359-365: rows (c,m,g), columns (a,n,e)."""
    r, r2 = L.shape[0], Rt.shape[0]
    m, n = A.shape[1], A.shape[2]
    M = td(L, A, axes=(1, 0))
    M = td(M, Rt, axes=(4, 1))
    return M.transpose([1, 2, 5, 0, 3, 4]).reshape(r * m * r2, r * n * r2)


def micro_matvec_als(L, A, Rt, v):
    """Matrix-free product with the matrix of micro_matrix_als (SURVEY.md a4'; not in the reference)."""
    t = td(L, v, axes=(0, 0))                       # [b, c, n, e]
    t = td(t, A, axes=([0, 2], [0, 2]))             # [c, e, m, d]
    return td(t, Rt, axes=([1, 3], [0, 1])).transpose(0, 1, 2)   # [c, m, g]


def micro_matrix_mals(L, A1, A2, Rt):
    """sle.py:381-388: rows (c,m,m2,g), columns (a,n,n2,e)."""
    r, r3 = L.shape[0], Rt.shape[0]
    m, n, m2, n2 = A1.shape[1], A1.shape[2], A2.shape[1], A2.shape[2]
    M = td(L, A1, axes=(1, 0))
    M = td(M, A2, axes=(4, 0))
    M = td(M, Rt, axes=(6, 1))
    return M.transpose([1, 2, 4, 7, 0, 3, 5, 6]).reshape(r * m * m2 * r3, r * n * n2 * r3)


def micro_matvec_mals(L, A1, A2, Rt, v):
    """Two-site matrix-free product; v [r, n, n2, r3] (not in the reference)."""
    t = td(L, v, axes=(0, 0))                       # [b, c, n, j, e]
    t = td(t, A1, axes=([0, 2], [0, 2]))            # [c, j, e, m, d]
    t = td(t, A2, axes=([1, 4], [2, 0]))            # [c, e, m, k, f]
    return td(t, Rt, axes=([1, 4], [0, 1]))         # [c, m, k, g]


def micro_rhs_als(bL, b, bR):
    """sle.py:424-428 / evp.py:377-379 (returned unflattened [r, m, r2])."""
    t = td(bL, b, axes=(0, 0))
    return td(t, bR, axes=(2, 0))


def micro_rhs_mals(bL, b1, b2, bR):
    """sle.py:464-470 (returned unflattened [r, m, m2, r3])."""
    t = td(bL, b1, axes=(0, 0))
    t = td(t, b2, axes=(2, 0))
    return td(t, bR, axes=(3, 0))
