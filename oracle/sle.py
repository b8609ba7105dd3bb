"""Oracle restatement of scikit_tt/solvers/sle.py (ALS / MALS for A x = b) on core lists."""
import numpy as np
import scipy.linalg as sla

from . import kernels as K
from .tt import copy_cores


def _x3(core):  # [r, n, 1, r2] -> [r, n, r2]
    return core[:, :, 0, :]


def _solve_als(M, f, solver):
    # sle.py:505-509
    if solver == 'solve':
        return np.linalg.solve(M, f)
    lu = sla.lu_factor(M, check_finite=False)
    return sla.lu_solve(lu, f, check_finite=False)


def _solve_mals(M, f, solver):
    # sle.py:588-594 (scipy.linalg.solve: structure detection)
    if solver == 'solve':
        return sla.solve(M, f, check_finite=False)
    lu = sla.lu_factor(M, check_finite=False)
    return sla.lu_solve(lu, f, check_finite=False)


def als(op, x0, b, repeats=1, solver='solve'):
    """sle.py:10-95.  op/x0/b: lists of 4-D cores; returns the list of solution cores."""
    x = copy_cores(x0)                                                    # sle.py:45
    d = len(op)
    one3, one2 = np.ones((1, 1, 1)), np.ones((1, 1))
    Lop, Lrhs, Rop, Rrhs = [None] * d, [None] * d, [None] * d, [None] * d
    for i in range(d - 1, -1, -1):                                        # sle.py:54-56
        Rop[i] = one3 if i == d - 1 else K.stack_right_op(Rop[i + 1], _x3(x[i + 1]), op[i + 1])
        Rrhs[i] = one2 if i == d - 1 else K.stack_right_rhs(Rrhs[i + 1], _x3(b[i + 1]), _x3(x[i + 1]))
    for _ in range(repeats):                                              # sle.py:62
        for i in range(d):                                                # forward, sle.py:65-77
            Lop[i] = one3 if i == 0 else K.stack_left_op(Lop[i - 1], _x3(x[i - 1]), op[i - 1])
            Lrhs[i] = one2 if i == 0 else K.stack_left_rhs(Lrhs[i - 1], _x3(b[i - 1]), _x3(x[i - 1]))
            if i < d - 1:
                r, n, r2 = Lop[i].shape[0], op[i].shape[2], Rop[i].shape[0]
                M = K.micro_matrix_als(Lop[i], op[i], Rop[i])
                f = K.micro_rhs_als(Lrhs[i], _x3(b[i]), Rrhs[i]).reshape(-1, 1)
                u = _solve_als(M, f, solver)
                q, _ = sla.qr(u.reshape(r * n, r2), mode='economic', check_finite=False)   # sle.py:517-525
                x[i] = q.reshape(r, n, 1, q.shape[1])
        for i in range(d - 1, -1, -1):                                    # backward, sle.py:80-90
            Rop[i] = one3 if i == d - 1 else K.stack_right_op(Rop[i + 1], _x3(x[i + 1]), op[i + 1])
            Rrhs[i] = one2 if i == d - 1 else K.stack_right_rhs(Rrhs[i + 1], _x3(b[i + 1]), _x3(x[i + 1]))
            r = Lop[i].shape[0]
            r2 = Rop[i].shape[0]
            n = op[i].shape[2]
            M = K.micro_matrix_als(Lop[i], op[i], Rop[i])
            f = K.micro_rhs_als(Lrhs[i], _x3(b[i]), Rrhs[i]).reshape(-1, 1)
            u = _solve_als(M, f, solver)
            if i > 0:
                _, q = sla.rq(u.reshape(r, n * r2), mode='economic', check_finite=False)   # sle.py:533-541
                x[i] = q.reshape(q.shape[0], n, 1, r2)
            else:
                x[i] = u.reshape(r, n, 1, r2)                                              # sle.py:546
    return x


def _split(u, r, n, n2, r3, threshold, max_rank):
    # sle.py:603-614 / :626-639
    U, s, V = sla.svd(u.reshape(r * n, n2 * r3), full_matrices=False, check_finite=False, lapack_driver='gesvd')
    if threshold != 0:
        keep = np.where(s / s[0] > threshold)[0]
        U, s, V = U[:, keep], s[keep], V[keep, :]
    if max_rank != np.inf:
        k = int(min(U.shape[1], max_rank))
        U, s, V = U[:, :k], s[:k], V[:k, :]
    return U, s, V


def mals(op, x0, b, repeats=1, solver='solve', threshold=1e-12, max_rank=np.inf):
    """sle.py:98-191."""
    x = copy_cores(x0)
    d = len(op)
    one3, one2 = np.ones((1, 1, 1)), np.ones((1, 1))
    Lop, Lrhs, Rop, Rrhs = [None] * d, [None] * d, [None] * d, [None] * d
    for i in range(d - 1, 0, -1):                                         # sle.py:148-151
        Rop[i] = one3 if i == d - 1 else K.stack_right_op(Rop[i + 1], _x3(x[i + 1]), op[i + 1])
        Rrhs[i] = one2 if i == d - 1 else K.stack_right_rhs(Rrhs[i + 1], _x3(b[i + 1]), _x3(x[i + 1]))

    def micro(i):
        M = K.micro_matrix_mals(Lop[i], op[i], op[i + 1], Rop[i + 1])
        f = K.micro_rhs_mals(Lrhs[i], _x3(b[i]), _x3(b[i + 1]), Rrhs[i + 1]).reshape(-1, 1)
        return _solve_mals(M, f, solver)

    for _ in range(repeats):
        for i in range(d - 1):                                            # forward, sle.py:160-172
            Lop[i] = one3 if i == 0 else K.stack_left_op(Lop[i - 1], _x3(x[i - 1]), op[i - 1])
            Lrhs[i] = one2 if i == 0 else K.stack_left_rhs(Lrhs[i - 1], _x3(b[i - 1]), _x3(x[i - 1]))
            if i < d - 2:
                r, n, n2, r3 = Lop[i].shape[0], op[i].shape[2], op[i + 1].shape[2], Rop[i + 1].shape[0]
                U, s, V = _split(micro(i), r, n, n2, r3, threshold, max_rank)
                x[i] = U.reshape(r, n, 1, s.shape[0])                    # sle.py:616-620
        for i in range(d - 2, -1, -1):                                    # backward, sle.py:175-186
            Rop[i + 1] = one3 if i + 1 == d - 1 else K.stack_right_op(Rop[i + 2], _x3(x[i + 2]), op[i + 2])
            Rrhs[i + 1] = one2 if i + 1 == d - 1 else K.stack_right_rhs(Rrhs[i + 2], _x3(b[i + 2]), _x3(x[i + 2]))
            r, n, n2, r3 = Lop[i].shape[0], op[i].shape[2], op[i + 1].shape[2], Rop[i + 1].shape[0]
            U, s, V = _split(micro(i), r, n, n2, r3, threshold, max_rank)
            x[i + 1] = V.reshape(s.shape[0], n2, 1, r3)                   # sle.py:645
            if i == 0:
                x[i] = (U * s[None, :]).reshape(r, n, 1, s.shape[0])     # sle.py:647-650
    return x


def residual(op, x, b):
    """|| A x - b || / || b ||  evaluated densely-in-TT (tensor_train.py:2035-2074 semantics)."""
    from . import tt as T
    return T.norm(T.sub(T.matmul(op, x), b)) / T.norm(b)
