"""TT helpers of the oracle on plain lists of 4-D numpy cores [r, m, n, r2]."""
import numpy as np
import scipy.linalg as sla


def ranks_of(cores):
    return [c.shape[0] for c in cores] + [cores[-1].shape[3]]


def copy_cores(cores):
    return [c.copy() for c in cores]


def full(cores):
    """Dense array of shape row_dims + col_dims (tensor_train.py:1029-1059 semantics)."""
    d = len(cores)
    t = cores[0]
    for c in cores[1:]:
        t = np.tensordot(t, c, axes=(t.ndim - 1, 0))
    t = t.reshape(t.shape[1:-1])                     # drop boundary ranks
    perm = [2 * i for i in range(d)] + [2 * i + 1 for i in range(d)]
    return t.transpose(perm)


def matricize(cores):
    rows = int(np.prod([c.shape[1] for c in cores]))
    cols = int(np.prod([c.shape[2] for c in cores]))
    return full(cores).reshape(rows, cols)


def _svd(mat):
    # tensor_train.py:1162-1171: gesdd first, gesvd as the fallback
    try:
        return sla.svd(mat, full_matrices=False, overwrite_a=False, check_finite=False)
    except Exception:
        return sla.svd(mat, full_matrices=False, overwrite_a=False, check_finite=False, lapack_driver='gesvd')


def _truncate(u, s, v, threshold, max_rank):
    # tensor_train.py:1173-1182 / sle.py:608-614: strict relative threshold, then max_rank
    if threshold != 0:
        keep = np.where(s / s[0] > threshold)[0]
        u, s, v = u[:, keep], s[keep], v[keep, :]
    if max_rank != np.inf:
        k = int(min(u.shape[1], max_rank))
        u, s, v = u[:, :k], s[:k], v[:k, :]
    return u, s, v


def _max_ranks(order, max_rank):
    if isinstance(max_rank, list):
        return max_rank
    return [1] + [max_rank] * (order - 1) + [1]


def ortho_left(cores, start_index=0, end_index=None, threshold=0.0, max_rank=np.inf):
    """tensor_train.py:1159-1196 (in place on the list, returns it)."""
    d = len(cores)
    if end_index is None:
        end_index = d - 2
    mr = _max_ranks(d, max_rank)
    for i in range(start_index, end_index + 1):
        r, m, n, r2 = cores[i].shape
        u, s, v = _svd(cores[i].reshape(r * m * n, r2))
        u, s, v = _truncate(u, s, v, threshold, mr[i + 1])
        k = u.shape[1]
        cores[i] = u.reshape(r, m, n, k)
        cores[i + 1] = np.tensordot(s[:, None] * v, cores[i + 1], axes=(1, 0))
    return cores


def ortho_right(cores, start_index=None, end_index=1, threshold=0.0, max_rank=np.inf):
    """tensor_train.py:1265-1301 (in place on the list, returns it)."""
    d = len(cores)
    if start_index is None:
        start_index = d - 1
    mr = _max_ranks(d, max_rank)
    for i in range(start_index, end_index - 1, -1):
        r, m, n, r2 = cores[i].shape
        u, s, v = _svd(cores[i].reshape(r, m * n * r2))
        u, s, v = _truncate(u, s, v, threshold, mr[i])
        k = v.shape[0]
        cores[i] = v.reshape(k, m, n, r2)
        p = cores[i - 1]
        cores[i - 1] = (p.reshape(-1, p.shape[3]) @ (u * s[None, :])).reshape(p.shape[0], p.shape[1], p.shape[2], k)
    return cores


def norm(cores, p=2):
    """tensor_train.py:1334-1430."""
    cs = copy_cores(cores)
    if p == 1:
        if all(c.shape[1] == 1 for c in cs):
            cs = [c.transpose(0, 2, 1, 3) for c in cs]
        cs = [c.sum(axis=1, keepdims=True) for c in cs]
        return np.max(matricize(cs))
    if p == 2:
        cs = [c.reshape(c.shape[0], c.shape[1] * c.shape[2], 1, c.shape[3]) for c in cs]
        cs = ortho_right(cs)
        return np.linalg.norm(cs[0].reshape(-1))
    raise ValueError('p must be 1 or 2.')


def scale(cores, alpha):
    """tensor_train.py:368-420: the scalar goes into the first core."""
    out = copy_cores(cores)
    out[0] = alpha * out[0]
    return out


def add(a, b):
    """tensor_train.py:282-344: block-diagonal concatenation of the cores."""
    d = len(a)
    out = []
    for i in range(d):
        ra, m, n, ra2 = a[i].shape
        rb, _, _, rb2 = b[i].shape
        dt = np.result_type(a[i], b[i])
        if d == 1:
            out.append((a[i] + b[i]).astype(dt))
        elif i == 0:
            out.append(np.concatenate([a[i], b[i]], axis=3).astype(dt))
        elif i == d - 1:
            out.append(np.concatenate([a[i], b[i]], axis=0).astype(dt))
        else:
            c = np.zeros((ra + rb, m, n, ra2 + rb2), dtype=dt)
            c[:ra, :, :, :ra2] = a[i]
            c[ra:, :, :, ra2:] = b[i]
            out.append(c)
    return out


def eye(dims):
    """tensor_train.py:1870-1894."""
    return [np.eye(k).reshape(1, k, k, 1) for k in dims]


def sub(a, b):
    return add(a, scale(b, -1.0))


def matmul(op, x):
    """tensor_train.py:422-503: core-wise contraction of the column index of op with the row index of x."""
    out = []
    for A, X in zip(op, x):
        R, m, n, R2 = A.shape
        r, n_, k, r2 = X.shape
        c = np.einsum('bmnd,anke->bamkde', A, X).reshape(R * r, m, k, R2 * r2)
        out.append(c)
    return out
