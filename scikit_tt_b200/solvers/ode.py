"""Time stepping on top of the GPU ALS/MALS solvers -- `implicit_euler`, `trapezoidal_rule` and `adaptive_step_size`
with the call surfaces of scikit_tt/solvers/ode.py:249-330, :366-450, :487-636 of PGelss/scikit_tt."""
import time as _time

import numpy as np

from .. import tensor_train as tt
from .. import utils as utl
from . import sle
from .sle import __construct_stack_right_op, __construct_stack_left_op, __construct_micro_matrix_als, \
    __construct_micro_matrix_mals  # noqa: F401  (the reference's ode.py:13 imports these names too)


def implicit_euler(operator, initial_value, initial_guess, step_sizes, repeats=1, tt_solver='als', threshold=1e-12,
                   max_rank=np.inf, micro_solver='solve', normalize=1, progress=True):
    """Implicit Euler for dx/dt = operator @ x: every step solves (I - h A) x_{k+1} = x_k with sle.als / sle.mals
    (ode.py:309-317), normalises in the p-norm `normalize` (ode.py:320-321) and appends a copy (ode.py:324).
    Returns [initial_value, x_1, x_2, ...]."""
    start = utl.progress('Running implicit Euler method', 0, show=progress)
    solution = [initial_value]
    cur = initial_guess
    n_steps = len(step_sizes)
    for i in range(n_steps):
        lhs = tt.eye(operator.row_dims) - step_sizes[i] * operator
        if tt_solver == 'als':
            cur = sle.als(lhs, cur, solution[i], solver=micro_solver, repeats=repeats)
        if tt_solver == 'mals':
            cur = sle.mals(lhs, cur, solution[i], solver=micro_solver, threshold=threshold, repeats=repeats,
                           max_rank=max_rank)
        if normalize > 0:
            cur = (1 / cur.norm(p=normalize)) * cur
        solution.append(cur.copy())
        utl.progress('Running implicit Euler method', 100 * (i + 1) / n_steps, show=progress,
                     cpu_time=_time.time() - start)
    return solution


def trapezoidal_rule(operator, initial_value, initial_guess, step_sizes, repeats=1, tt_solver='als', threshold=1e-12,
                     max_rank=np.inf, micro_solver='solve', normalize=1, progress=True):
    """Trapezoidal rule (scikit_tt/solvers/ode.py:366-450): every step solves
    (I - h/2 A) x_{k+1} = (I + h/2 A) x_k with sle.als / sle.mals, then normalises and appends a copy."""
    start = utl.progress('Running trapezoidal rule', 0, show=progress)
    solution = [initial_value]
    cur = initial_guess
    n_steps = len(step_sizes)
    for i in range(n_steps):
        lhs = tt.eye(operator.row_dims) - 0.5 * step_sizes[i] * operator
        rhs = (tt.eye(operator.row_dims) + 0.5 * step_sizes[i] * operator).dot(solution[i])   # ode.py:431-437
        if tt_solver == 'als':
            cur = sle.als(lhs, cur, rhs, solver=micro_solver, repeats=repeats)
        if tt_solver == 'mals':
            cur = sle.mals(lhs, cur, rhs, solver=micro_solver, repeats=repeats, threshold=threshold, max_rank=max_rank)
        if normalize > 0:
            cur = (1 / cur.norm(p=normalize)) * cur
        solution.append(cur.copy())
        utl.progress('Running trapezoidal rule', 100 * (i + 1) / n_steps, show=progress, cpu_time=_time.time() - start)
    return solution


def _shifted(operator, h):
    """I - h * operator as a TT operator."""
    return tt.eye(operator.row_dims) - h * operator


def _unit(t, p):
    return (1 / t.norm(p=p)) * t


def adaptive_step_size(operator, initial_value, initial_guess, time_end, step_size_first=1e-10, repeats=1, solver='solve',
                       error_tol=1e-1, closeness_tol=0.5, step_size_min=1e-14, step_size_max=10, closeness_min=1e-3,
                       factor_max=2, factor_safe=0.9, second_method='two_step_Euler', normalize=1, progress=True):
    """Step-size control with the call surface and decisions of scikit_tt/solvers/ode.py:487-636.

    Per trial step h: a first-order candidate (one implicit Euler step, 1-norm normalised) and a higher-order one (two
    Euler steps of h/2, or one trapezoidal step).  The trial is accepted when both the local error
    ||low - high|| / ||low|| and the relative change of the closeness ||A low|| stay within their tolerances; either way
    the next h is the current one scaled by min(factor_max, factor_safe * tolerance / measured).  The accepted state is the
    higher-order candidate; the next guess is the first-order one.  Returns (states, times)."""
    t0 = utl.progress('Running adaptive step size method', 0, show=progress)
    als = lambda lhs, guess, rhs: sle.als(lhs, guess.copy(), rhs, solver=solver, repeats=repeats)
    states, times = [initial_value], [0]
    guess, now, h = initial_guess, 0, step_size_first
    drift = operator.dot(initial_value).norm()                                   # closeness of the current state
    high = []
    while now < time_end and drift > closeness_min and h > step_size_min:        # ode.py:584
        last = states[-1]
        low = _unit(als(_shifted(operator, h), guess, last), 1)
        if second_method == 'two_step_Euler':
            half = _shifted(operator, 0.5 * h)
            high = als(half, als(half, guess, last), last)
        if second_method == 'trapezoidal_rule':
            high = als(_shifted(operator, 0.5 * h), guess, _shifted(operator, -0.5 * h).dot(last))
        if normalize > 0:
            high = _unit(high, normalize)
        drift_new = operator.dot(low).norm()
        room_error = error_tol / ((low - high).norm() / low.norm())
        room_drift = closeness_tol / np.abs((drift_new - drift) / drift)
        h_next = np.amin([factor_max, factor_safe * room_error, factor_safe * room_drift]) * h
        if room_error > 1 and room_drift > 1:                                    # accept (ode.py:618-630)
            now = np.min([now + h, time_end])
            h = np.amin([h_next, time_end - now, step_size_max])
            states.append(high.copy())
            times.append(now)
            guess, drift = low, drift_new
            utl.progress('Running adaptive step size method', 100 * now / time_end, show=progress,
                         cpu_time=_time.time() - t0)
        else:
            h = h_next
    return states, times


# ------------------------------------------------------------------------------------------------------------------
# Time-dependent variational principle: `tdvp1site` / `tdvp2site` with the call surfaces of
# scikit_tt/solvers/ode.py:1188-1395.  The sweep skeleton is the ALS one -- the same interface stacks and dense micro
# matrices (sle.py:194-390, imported by name there: ode.py:13) -- with the micro solve replaced by the action of a matrix
# exponential on the core (ode.py:1398-1614).  Everything between the upload of the initial value and the download of a
# finished time step runs on the device: stacks and micro matrices through the C-ABI entry points of the ALS sweep, the
# exponential action as a Krylov projection (matvecs and classical Gram-Schmidt through the contraction engine) whose small
# projected matrix is exponentiated by `sktt_expm_small`, QR / RQ / truncated SVD by the orthonormalisation kernels.
def _dense_matvec(dev, M, v):
    N = M.shape[0]
    return dev.matmul(M, v.reshape(N, 1)).reshape(-1)


def _expm_action(dev, M, v, c, tol=1e-14, m_max=64):
    """exp(c * M) v for a dense device matrix M [N, N] and vector v [N] (complex128), c a complex scalar -- what
    scipy.sparse.linalg.expm_multiply delivers at ode.py:1437-1508.  Arnoldi with CGS2; the Krylov dimension grows until the
    a-posteriori estimate |c| h_{m+1,m} |e_m^T exp(c H_m) e_1| drops below tol (a space of dimension N is exact); if m_max
    vectors do not suffice the step is halved (exp(cM) = exp(cM/2)^2)."""
    import torch
    N = v.numel()
    beta0 = dev.nrm2(v)
    if beta0 == 0.0:
        return v.clone()
    big = 1 << 40
    m_cap = min(N, m_max)
    V = dev.empty((m_cap + 1, N), torch.complex128)
    H = torch.zeros((m_cap + 1, m_cap), dtype=torch.complex128, device=dev.device)
    dev.axpby(1.0 / beta0, v, 0.0, v, out=V[0])
    check = 4
    for j in range(m_cap):
        w = _dense_matvec(dev, M, V[j])
        hsum = None
        for _ in range(2):
            h = dev.empty((j + 1,), torch.complex128)
            dev.gemm2(j + 1, 1, N, V, (big, 0, N), (big, 0, 1), w, (big, 0, 1), (big, 0, 0), h, (big, 0, 1), (big, 0, 0), conjA=1)
            dev.gemm2(N, 1, j + 1, V, (big, 0, 1), (big, 0, N), h, (big, 0, 1), (big, 0, 0), w, (big, 0, 1), (big, 0, 0),
                      alpha=(-1.0, 0.0), beta=(1.0, 0.0))
            hsum = h if hsum is None else dev.axpby(1.0, h, 1.0, hsum)
        H[: j + 1, j] = hsum
        hn = dev.nrm2(w)
        H[j + 1, j] = hn
        m = j + 1
        full = m == N or hn == 0.0
        if hn > 0.0:
            dev.axpby(1.0 / hn, w, 0.0, w, out=V[j + 1])
        if full or m == m_cap or m >= check:
            E = dev.expm_small(H[:m, :m].contiguous(), c)
            col = E[:, 0]
            err = abs(c) * hn * abs(complex(col[m - 1].item()))
            if full or err <= tol:
                coef = (col * beta0).contiguous()
                out = dev.empty((N,), torch.complex128)
                dev.gemm2(N, 1, m, V, (big, 0, 1), (big, 0, N), coef, (big, 0, 1), (big, 0, 0), out, (big, 0, 1), (big, 0, 0))
                return out
            check = m + 4
    half = _expm_action(dev, M, v, 0.5 * c, tol=tol, m_max=m_max)        # Krylov space exhausted: two half steps
    return _expm_action(dev, M, half, 0.5 * c, tol=tol, m_max=m_max)


def _local_krylov(dev, M, v, dimension, step_size):
    """ode.local_krylov (ode.py:1689-1757) on the device, arithmetic as there: `dimension` steps of the three-term Lanczos
    recurrence started from the UN-normalised core (krylov_tensors[0] = initial_value, ode.py:1727), tridiagonal T, and
    exp(-i T step_size) e_1 combined with the Krylov tensors."""
    import torch
    N = v.numel()
    big = 1 << 40
    K = dev.empty((dimension, N), torch.complex128)
    T = torch.zeros((dimension, dimension), dtype=torch.complex128, device=dev.device)
    K[0].copy_(v)
    w = _dense_matvec(dev, M, K[0])
    alpha = dev.dotc(w, K[0])                                             # conj(w)^T k  (ode.py:1729)
    T[0, 0] = alpha
    w = dev.axpby(1.0, w, -alpha, K[0])
    for i in range(1, dimension):
        beta = dev.nrm2(w)
        T[i, i - 1] = beta
        T[i - 1, i] = beta
        dev.axpby(1.0 / beta, w, 0.0, w, out=K[i])
        w = _dense_matvec(dev, M, K[i])
        alpha = dev.dotc(w, K[i])
        T[i, i] = alpha
        w = dev.axpby(1.0, w, -alpha, K[i])
        w = dev.axpby(1.0, w, -beta, K[i - 1])
    E = dev.expm_small(T, -1j * step_size)                                # ode.py:1751
    coef = E[:, 0].contiguous()
    out = dev.empty((N,), torch.complex128)
    dev.gemm2(N, 1, dimension, K, (big, 0, 1), (big, 0, N), coef, (big, 0, 1), (big, 0, 0), out, (big, 0, 1), (big, 0, 0))
    return out


class _Tdvp:
    """Device-resident state of a TDVP run: operator and solution cores (complex128), interface stacks."""

    def __init__(self, operator, initial_value, local_solver):
        import torch
        from .. import _device
        from . import _local
        self.dev = dev = _device.get_device()
        self.torch = torch
        self.d = operator.order
        self.A = _local.Uploaded(dev, operator, torch.complex128, vector=False).cores
        self.x = list(_local.Uploaded(dev, initial_value, torch.complex128, vector=True).cores)
        self.L, self.R = [None] * self.d, [None] * self.d
        self.one3 = torch.ones((1, 1, 1), dtype=torch.complex128, device=dev.device)
        if not local_solver or local_solver['method'] == 'exact':          # ode.py:1239-1243
            self.krylov_dim = None
        else:
            self.krylov_dim = local_solver.get('dimension') or 5

    def left(self, i):                                                     # sle.py:194-219
        self.L[i] = self.one3 if i == 0 else self.dev.stack_left_op(self.L[i - 1], self.x[i - 1], self.A[i - 1])

    def right(self, i):                                                    # sle.py:250-276
        self.R[i] = self.one3 if i == self.d - 1 else self.dev.stack_right_op(self.R[i + 1], self.x[i + 1], self.A[i + 1])

    def evolve(self, M, v, scale, step_size):
        """exp(-i * scale * step_size * M) v: exact (ode.py:1437: expm_multiply(-1j*step*0.5*M, .)) or local_krylov with
        the signed step (ode.py:1439: local_krylov(M, ., dim, 0.5*step))."""
        v = v.reshape(-1).contiguous()
        if self.krylov_dim is None:
            return _expm_action(self.dev, M, v, -1j * scale * step_size)
        return _local_krylov(self.dev, M, v, self.krylov_dim, scale * step_size)

    def project(self, M, Q):
        """Q^H M Q for an isometry Q [N, k] given as a dense device matrix (ode.py:1450-1451, :1492-1493)."""
        dev = self.dev
        return dev.matmul(Q, dev.matmul(M, Q), opa='C')

    def download(self):
        from . import _local
        from ..tensor_train import TT
        return TT(_local.download_vector_cores(self.x))


def _kron_left(torch, q, r2):
    """Q~[(a, n, a2), (k, a2')] = q[(a, n), k] delta(a2, a2')  (ode.py:1449-1450) -- data placement only."""
    P, k = q.shape
    out = torch.zeros((P, r2, k, r2), dtype=q.dtype, device=q.device)
    idx = torch.arange(r2, device=q.device)
    out[:, idx, :, idx] = q.unsqueeze(0).expand(r2, P, k)
    return out.reshape(P * r2, k * r2)


def _kron_right(torch, q, r1):
    """Q~[(a, n, a2), (a', k)] = delta(a, a') q[k, (n, a2)]  (ode.py:1491-1492)."""
    k, P = q.shape
    out = torch.zeros((r1, P, r1, k), dtype=q.dtype, device=q.device)
    idx = torch.arange(r1, device=q.device)
    out[idx, :, idx, :] = q.t().unsqueeze(0).expand(r1, P, k)
    return out.reshape(r1 * P, r1 * k)


def tdvp1site(operator, initial_value, step_size, number_of_steps, local_solver=None, normalize=0):
    """One-site TDVP (ode.py:1188-1287).  Returns [initial_value, x_1, ..., x_number_of_steps]."""
    st = _Tdvp(operator, initial_value, local_solver)
    dev, torch, d, x = st.dev, st.torch, st.d, st.x
    solution = [initial_value]
    for i in range(d - 1, -1, -1):
        st.right(i)
    for _ in range(number_of_steps):
        for i in range(d):                                                  # first half sweep (ode.py:1253-1263)
            st.left(i)
            M = dev.micro_matrix_als(st.L[i], st.A[i], st.R[i])
            r1, n, r2 = x[i].shape
            if i < d - 1:
                core = st.evolve(M, x[i], 0.5, step_size).reshape(r1 * n, r2)
                q, rfac = dev.qr(core, want_r=True)                         # ode.py:1442
                k = q.shape[1]
                x[i] = q.reshape(r1, n, k)
                Mk = st.project(M, _kron_left(torch, q, r2))
                rfac = st.evolve(Mk, rfac, -0.5, step_size).reshape(k, r2)  # ode.py:1454-1460: backwards in time
                nxt = x[i + 1]
                x[i + 1] = dev.matmul(rfac, nxt.reshape(r2, -1)).reshape(k, nxt.shape[1], nxt.shape[2])
            else:
                x[i] = st.evolve(M, x[i], 1.0, step_size).reshape(r1, n, r2)   # ode.py:1468-1472
        for i in range(d - 1, -1, -1):                                      # second half sweep (ode.py:1266-1275)
            st.right(i)
            M = dev.micro_matrix_als(st.L[i], st.A[i], st.R[i])
            r1, n, r2 = x[i].shape
            if i > 0:
                core = x[i]
                if i < d - 1:                                               # the backward half sweep of the reference always
                    core = _expm_action(dev, M, core.reshape(-1).contiguous(), -1j * 0.5 * step_size)   # uses expm_multiply (ode.py:1482)
                rfac, q = dev.rq(core.reshape(r1, n * r2).contiguous(), want_r=True)   # ode.py:1485
                k = q.shape[0]
                x[i] = q.reshape(k, n, r2)
                Mk = st.project(M, _kron_right(torch, q, r1))
                rfac = _expm_action(dev, Mk, rfac.reshape(-1).contiguous(), 1j * 0.5 * step_size).reshape(r1, k)   # ode.py:1499
                prv = x[i - 1]
                x[i - 1] = dev.matmul(prv.reshape(-1, r1), rfac).reshape(prv.shape[0], prv.shape[1], k)
            else:
                x[i] = _expm_action(dev, M, x[i].reshape(-1).contiguous(), -1j * 0.5 * step_size).reshape(r1, n, r2)   # ode.py:1508
        tmp = st.download()
        if normalize > 0:
            tmp = (1 / tmp.norm(p=normalize)) * tmp
            from . import _local
            st.x[:] = _local.Uploaded(dev, tmp, torch.complex128, vector=True).cores
            x = st.x
        solution.append(tmp)
    return solution


def tdvp2site(operator, initial_value, step_size, number_of_steps, local_solver=None, threshold=1e-12, max_rank=50,
              normalize=0):
    """Two-site TDVP with truncated-SVD rank adaption (ode.py:1290-1395, update :1512-1614)."""
    st = _Tdvp(operator, initial_value, local_solver)
    dev, torch, d, x = st.dev, st.torch, st.d, st.x
    solution = [initial_value]
    for i in range(d - 1, 0, -1):
        st.right(i)
    eye = lambda k: torch.eye(k, dtype=torch.complex128, device=dev.device)

    def two_site(i):
        M = dev.micro_matrix_mals(st.L[i], st.A[i], st.A[i + 1], st.R[i + 1])
        r1, n, _ = x[i].shape
        _, n2, r3 = x[i + 1].shape
        sup = dev.matmul(x[i].reshape(r1 * n, -1), x[i + 1].reshape(-1, n2 * r3))
        sup = st.evolve(M, sup, 0.5, step_size).reshape(r1 * n, n2 * r3)    # ode.py:1545-1549
        U, S, Vh, k = dev.svd(sup.contiguous(), threshold=threshold, max_rank=max_rank)   # utils.truncated_svd
        return M, sup, U[:, :k].contiguous(), Vh[:k, :].contiguous(), k, (r1, n, n2, r3)

    for _ in range(number_of_steps):
        for i in range(d - 1):                                              # ode.py:1351-1361
            st.left(i)
            M, sup, u, vh, k, (r1, n, n2, r3) = two_site(i)
            x[i] = u.reshape(r1, n, k)
            sv = dev.matmul(u, sup, opa='C')                                # diag(s) v restricted to the kept rank
            x[i + 1] = sv.reshape(k, n2, r3)
            if i < d - 2:                                                   # ode.py:1566-1576
                Q = _kron_left(torch, u, n2 * r3)
                Mk = st.project(M, Q)
                x[i + 1] = st.evolve(Mk, x[i + 1], -0.5, step_size).reshape(k, n2, r3)
        for i in range(d - 2, -1, -1):                                      # ode.py:1364-1374
            st.right(i + 1)
            M, sup, u, vh, k, (r1, n, n2, r3) = two_site(i)
            x[i + 1] = vh.reshape(k, n2, r3)
            us = dev.matmul(sup, vh, opb='C')                               # u diag(s)
            x[i] = us.reshape(r1, n, k)
            if i > 0:                                                       # ode.py:1604-1614
                # v~ = kron(eye(r1 n), v^T): [(a, n, j, a3), (a', n', k)] = delta * vh[k, (j, a3)]
                Q = _kron_right(torch, vh, r1 * n)
                Mk = st.project(M, Q)
                x[i] = st.evolve(Mk, x[i], -0.5, step_size).reshape(r1, n, k)
        tmp = st.download()
        if normalize > 0:
            tmp = (1 / tmp.norm(p=normalize)) * tmp
            from . import _local
            st.x[:] = _local.Uploaded(dev, tmp, torch.complex128, vector=True).cores
            x = st.x
        solution.append(tmp)
    return solution
