"""Alternating linear schemes for A x = b in TT format -- same call surface as
scikit_tt/solvers/sle.py of PGelss/scikit_tt (`als` :10, `mals` :98), executed on the GPU.

Per micro-step (reference lines in brackets): interface-stack updates [sle.py:194-305] ->
micro right-hand side [:393-472] -> micro system, either assembled densely and LU-factorised
exactly as the reference does [:308-390, :505-509, :588-594] or, where that matrix cannot exist,
applied matrix-free inside CG / GMRES -> Householder QR / RQ of the solved core [:517-541] (ALS) or
truncated SVD split of the solved super-core [:603-650] (MALS).  All of it runs through the C-ABI of
libsktt_b200.so; cores and stacks stay in HBM between micro-steps.
"""
import numpy as np
import torch

from .. import _device
from ..tensor_train import TT
from . import _local


class _State:
    def __init__(self, operator, initial_guess, right_hand_side):
        self.dev = dev = _device.get_device()
        cplx = _local.any_complex(operator, initial_guess, right_hand_side)
        self.dtype = torch.complex128 if cplx else torch.float64
        self.d = operator.order
        self.A = _local.Uploaded(dev, operator, self.dtype, vector=False)
        self.b = _local.Uploaded(dev, right_hand_side, self.dtype, vector=True)
        self.x = list(_local.Uploaded(dev, initial_guess, self.dtype, vector=True).cores)   # sle.py:45 (copy)
        d = self.d
        self.Lop, self.Rop = [None] * d, [None] * d
        self.Lrhs, self.Rrhs = [None] * d, [None] * d
        self.one3 = _local.ones(dev, (1, 1, 1), self.dtype)
        self.one2 = _local.ones(dev, (1, 1), self.dtype)
        self.cache = {}                                  # per-call verdicts shared by the micro solves (see _local.solve_micro)
        self.out = None                                  # streamed-out result cores (see stream_out)
        self.stream_results = False                      # set by the public entry points: copy final cores out early
        self.group = None                                # process group of a rank-sharded call (als(..., group=))
        self.px = None                                   # its peer-mapped exchange buffers (multi.PeerExchange)

    # sle.py:194-247
    def left(self, i):
        dev = self.dev
        if i == 0:
            self.Lop[i], self.Lrhs[i] = self.one3, self.one2
        else:
            with _local.phase(dev, 'stacks'):
                self.Lop[i] = dev.stack_left_op(self.Lop[i - 1], self.x[i - 1], self.A[i - 1])
                self.Lrhs[i] = dev.stack_left_rhs(self.Lrhs[i - 1], self.b[i - 1], self.x[i - 1])

    # sle.py:250-305
    def right(self, i):
        dev = self.dev
        if i == self.d - 1:
            self.Rop[i], self.Rrhs[i] = self.one3, self.one2
        else:
            with _local.phase(dev, 'stacks'):
                self.Rop[i] = dev.stack_right_op(self.Rop[i + 1], self.x[i + 1], self.A[i + 1])
                self.Rrhs[i] = dev.stack_right_rhs(self.Rrhs[i + 1], self.b[i + 1], self.x[i + 1])

    def reset(self, x_cores):
        """Restart from another set of device solution cores (bench: same problem, timed repeatedly)."""
        self.x[:] = list(x_cores)
        self.out = None

    # -- results leave the device while the sweep is still running: a core is final once the last backward half sweep
    #    has passed it (sle.py:541, :546), so its device-to-host copy is queued on a second stream right there and
    #    overlaps the remaining micro steps
    def stream_out(self, i):
        dev = self.dev
        if not self.stream_results:
            return
        if self.out is None:
            bound = [int(c.numel()) for c in self.x]               # ALS ranks never grow (sle.py:522, :538)
            total = sum(bound)
            if total * self.x[0].element_size() < dev.PIN_RESULT_MIN_BYTES:
                self.out = False
            else:
                offs = np.concatenate([[0], np.cumsum(bound)]).tolist()
                self.out = {'host': torch.empty(total, dtype=self.dtype, pin_memory=True), 'offs': offs, 'bound': bound,
                            'done': {}, 'stream': dev.copy_stream()}
        if self.out is False:
            return
        o = self.out
        c = self.x[i]
        n = int(c.numel())
        if n > o['bound'][i]:
            self.out = False                                         # cannot happen for ALS; fall back to the plain download
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(o['stream']):
            o['stream'].wait_event(ev)
            if n:
                o['host'][o['offs'][i]:o['offs'][i] + n].copy_(c.reshape(-1), non_blocking=True)
        o['done'][i] = (c, tuple(c.shape))                           # keeps the device core alive until the copy has run

    def result(self):
        o = self.out
        if isinstance(o, dict) and len(o['done']) == self.d and all(o['done'][i][0] is self.x[i] for i in range(self.d)):
            o['stream'].synchronize()
            hnp = o['host'].numpy()
            cores = []
            for i in range(self.d):
                r, n, r2 = o['done'][i][1]
                cores.append(hnp[o['offs'][i]:o['offs'][i] + r * n * r2].reshape(r, n, 1, r2))
            self.out = None
            return TT(cores)
        self.out = None
        return TT(_local.download_vector_cores(self.x))


def als(operator, initial_guess, right_hand_side, repeats=1, solver='solve', group=None):
    """ALS sweeps for operator @ x = right_hand_side (sle.py:10-95).

    solver: 'solve' / 'lu' (reference values; both an LU with partial pivoting here as there) pick the
    dense path while r*n*r' <= _local.DENSE_LIMIT and the matrix-free path above it; 'dense', 'cg',
    'gmres' force one.  Returns a new TT; inputs are not modified.

    group (not in the reference): a torch.distributed process group, one process per GPU of one NVSwitch domain, every
    rank calling with the same arguments (BASELINE config 4: one system at very large rank).  The matrix-free micro-matvec
    is then sharded over the output solution-rank index across the ranks, its all-gather fused into the last contraction's
    epilogue over peer memory (solvers/multi.py, csrc/peer.cu); everything else of the sweep is replicated, and every
    rank returns the same TT.
    """
    _local.reset_stats()
    st = _State(operator, initial_guess, right_hand_side)
    st.group = group
    if st.group is not None:
        _sweeps_als(st, repeats, solver)
        if st.px is not None:
            st.px.check()                                    # the exchange buffers stay mapped for the next call (multi._PX_CACHE)
            st.px = None
        return st.result()
    st.stream_results = True
    _run_als(st, repeats, solver)
    return st.result()


def _run_als(st, repeats, solver):
    """The device-resident part of `als`: everything between the upload of the inputs and the download of the result.

    The matrix-free micro solves of a sweep are queued without draining the GPU (their outcomes stay on the device,
    _local.Deferred) and are inspected once at the end; if one of them did not reach the accepted residual -- an
    indefinite or non-Hermitian micro system, an iteration limit -- the sweeps are redone from the same initial cores
    with a host decision after every solve (preconditioned continuation, GMRES), which is what raises the errors."""
    x_initial = list(st.x)
    sig = tuple(tuple(c.shape) for c in st.A.cores) + tuple(tuple(c.shape) for c in x_initial)
    if solver in ('solve', 'lu', 'krylov', 'cg') and st.dtype == torch.float64 and _DEFER_MISSES.get(sig, 0) < 2:
        st.cache['defer'] = _local.Deferred(st.dev, 2 * st.d)
        # the QR / RQ steps of the optimistic pass run the sketched CholeskyQR alone (no Householder launch behind its
        # failure flag); a failure shows up in the sticky word that is read with the deferred outcomes of the micro solves
        st.dev.set_qr_deferred(True)
        check = lambda: bool(st.cache['defer'].check()) & (st.dev.qr_deferred_failures() == 0)
        ok = False
        try:
            ok = _sweeps_als(st, repeats, solver, check=check)
        except (_device.SkttError, np.linalg.LinAlgError) as exc:
            if getattr(exc, 'status', 0) == 3:                            # SKTT_ERR_CUDA: a broken context cannot be redone
                raise
            ok = False                                                    # let the host-driven pass raise what there is to raise
        finally:
            st.cache['defer'] = None
            st.dev.set_qr_deferred(False)
            if not ok:
                st.dev.qr_deferred_failures()                             # leave the sticky word clear
        if ok:
            _DEFER_MISSES.pop(sig, None)
            return
        _DEFER_MISSES[sig] = _DEFER_MISSES.get(sig, 0) + 1                # two misses in a row: this problem family is not
        st.x[:] = x_initial                                               # queued optimistically again (time steppers)
    _sweeps_als(st, repeats, solver)


_DEFER_MISSES = {}


def _sweeps_als(st, repeats, solver, check=None):
    """`check` (deferred micro solves only) is called at the end of every half sweep; a False ends the pass.

    Warm starts of the matrix-free micro solves: the reference throws the triangular factor of every QR / RQ away
    (sle.py:525, :541) because its direct solver needs no starting point.  Here that factor -- recovered as Q^H u resp.
    u Q^H -- is pushed into the neighbouring core, which turns the current iterate of the sweep into the starting vector
    of the next micro system (the same tensor, gauge moved by one core).  The micro systems and their solutions are
    unchanged; CG merely starts from the sweep's current residual instead of from zero."""
    dev, d, x = st.dev, st.d, st.x
    for i in range(d - 1, -1, -1):                                        # sle.py:54-56
        st.right(i)
    carry = None                                                          # gauge factor towards the next core to be solved
    st.out = None
    for rep in range(repeats):                                            # sle.py:62
        for i in range(d):                                                # first half sweep, sle.py:65-77
            st.left(i)
            if i < d - 1:
                guess = _pushed(dev, carry, x[i], left=True)
                u, (r, n, r2) = _micro_als(st, i, solver, guess)
                u2 = u.reshape(r * n, r2)
                with _local.phase(dev, 'qr'):
                    q = dev.qr(u2)                                        # sle.py:517-525
                    carry = dev.gauge_factor(q, u2, tall=True) if _wants_guess(solver, u2.numel()) else None
                x[i] = q.reshape(r, n, q.shape[1])
        if check is not None and not check():
            return False
        for i in range(d - 1, -1, -1):                                    # second half sweep, sle.py:80-90
            st.right(i)
            guess = _pushed(dev, carry, x[i], left=(i == d - 1))
            u, (r, n, r2) = _micro_als(st, i, solver, guess)
            carry = None
            if i > 0:
                u2 = u.reshape(r, n * r2)
                with _local.phase(dev, 'qr'):
                    q = dev.rq(u2)                                        # sle.py:533-541
                    carry = dev.gauge_factor(q, u2, tall=False) if _wants_guess(solver, u2.numel()) else None
                x[i] = q.reshape(q.shape[0], n, r2)
            else:
                x[i] = u.reshape(r, n, r2)                                # sle.py:546
            if rep == repeats - 1:
                st.stream_out(i)
        if check is not None and not check():
            return False
    return True


def _wants_guess(solver, N):
    return solver in ('cg', 'gmres', 'krylov') or (solver in ('solve', 'lu') and N > _local.SMALL_DENSE_LIMIT)


def _pushed(dev, carry, core, left):
    """Starting vector of the next micro system: the core as it stands with the gauge factor of the neighbour that was
    just orthonormalised pushed into it (left: factor [k, r] times core [r, n, r2]; else core [r, n, r2] times factor
    [r2, k]).  None when there is no factor -- the core itself is the starting vector then."""
    if carry is None or core.dim() != 3:
        return None
    r, n, r2 = core.shape
    if left:
        if carry.shape[1] != r:
            return None
        return dev.gauge_push(carry, core.reshape(r, n * r2), left=True).reshape(carry.shape[0], n, r2)
    if carry.shape[0] != r2:
        return None
    return dev.gauge_push(carry, core.reshape(r * n, r2), left=False).reshape(r, n, carry.shape[1])


def _micro_als(st, i, solver, guess=None):
    dev = st.dev
    L, R, A = st.Lop[i], st.Rop[i], st.A[i]
    with _local.phase(dev, 'micro_rhs'):
        f = dev.micro_rhs_als(st.Lrhs[i], st.b[i], st.Rrhs[i])            # sle.py:424-428
    r, n, r2 = L.shape[0], A.shape[2], R.shape[0]
    if guess is None or tuple(guess.shape) != (r, n, r2):
        guess = st.x[i] if tuple(st.x[i].shape) == (r, n, r2) else None
    if (st.group is not None and _wants_guess(solver, r * n * r2) and solver != 'gmres' and r >= 2 * _group_size(st.group)
            and r * n * r2 >= SHARD_MIN_UNKNOWNS):
        return _solve_sharded(st, L, A, R, f, guess), (r, n, r2)
    op = dev.local_op(L, A, R)
    with _local.phase(dev, 'solve'):
        u = _local.solve_micro(dev, solver, lambda: dev.micro_matrix_als(L, A, R), op, f, guess, st.cache)
    return u, (r, n, r2)


# Micro systems below this many unknowns are solved by every rank of `group` on its own (replicated): measured on 2 / 4 B200,
# the exchange of the sharded matvec costs more than the compute it saves at r = 128, n = 16 (262 144 unknowns: 16.6 ->
# 16.8 -> 10.2 half-sweeps/s at 1 / 2 / 4 GPUs), and pays at r = 256 (1 048 576 unknowns: 5.0 -> 6.8 -> 7.4 at 1 / 2 / 8).
SHARD_MIN_UNKNOWNS = 1 << 20


def _group_size(group):
    import torch.distributed as dist
    return dist.get_world_size(group)


def _solve_sharded(st, L, A, R, f, guess):
    """One micro solve with the matvec sharded over the ranks of st.group (multi.solve_sharded); the exchange buffers are
    created on first use, sized for the largest micro system of the call, and closed by `als`."""
    from . import multi
    dev = st.dev
    if st.px is None:
        nmax = max(int(st.x[i].shape[0] * st.A[i].shape[1] * st.x[i].shape[2]) for i in range(st.d))
        st.px = multi.peer_exchange(dev, nmax, st.dtype, st.group)
    return multi.solve_sharded(dev, lambda v: st.px.matvec(L, A, R, v), f, guess)


def mals(operator, initial_guess, right_hand_side, repeats=1, solver='solve', threshold=1e-12, max_rank=np.inf):
    """MALS sweeps (two-site micro systems, truncated-SVD core splitting; sle.py:98-191)."""
    _local.reset_stats()
    st = _State(operator, initial_guess, right_hand_side)
    _run_mals(st, repeats, solver, threshold, max_rank)
    return st.result()


def _run_mals(st, repeats, solver, threshold, max_rank):
    """Warm starts as in _sweeps_als: the reference keeps only U (forward, sle.py:620) resp. V (backward, sle.py:645) of
    the truncated SVD and drops diag(s) V resp. U diag(s); here that factor -- recovered as U^H u resp. u V^H -- stands in
    for the neighbouring core when the starting vector of the next two-site system is formed, so the matrix-free solve
    starts from the sweep's current (truncated) iterate.  The micro systems and their solutions are unchanged."""
    dev, d, x = st.dev, st.d, st.x
    for i in range(d - 1, 0, -1):                                         # sle.py:148-151
        st.right(i)
    for _ in range(repeats):
        carry = None                                                      # diag(s) V of the last split, as core i + 1
        for i in range(d - 1):                                            # sle.py:160-172
            st.left(i)
            if i < d - 2:
                u, (r, n, n2, r3) = _micro_mals(st, i, solver, left=carry)
                mat = u.reshape(r * n, n2 * r3)
                U, S, Vh, k = dev.svd_truncated(mat, threshold=threshold, max_rank=max_rank)   # sle.py:603-614
                uk = U[:, :k].contiguous()
                x[i] = uk.reshape(r, n, k)                                # sle.py:616-620
                carry = dev.matmul(uk, mat, opa='C').reshape(k, n2, r3) if mat.numel() > _local.SMALL_DENSE_LIMIT else None
        left, right = carry, None
        for i in range(d - 2, -1, -1):                                    # sle.py:175-186
            st.right(i + 1)
            u, (r, n, n2, r3) = _micro_mals(st, i, solver, left=left, right=right)
            left = None
            mat = u.reshape(r * n, n2 * r3)
            U, S, Vh, k = dev.svd_truncated(mat, threshold=threshold, max_rank=max_rank)                             # sle.py:626-639
            vh = Vh[:k, :].contiguous()
            x[i + 1] = vh.reshape(k, n2, r3)                              # sle.py:645
            right = None
            if i == 0:
                x[i] = dev.matmul(mat, vh, opb='C').reshape(r, n, k)      # U diag(s), sle.py:647-650
            elif mat.numel() > _local.SMALL_DENSE_LIMIT:
                right = dev.matmul(mat, vh, opb='C').reshape(r, n, k)     # U diag(s) as core i of the next two-site system


def _micro_mals(st, i, solver, left=None, right=None):
    """`left` / `right` stand in for x[i] / x[i + 1] when the starting vector is formed (see _run_mals)."""
    dev = st.dev
    L, R, A1, A2 = st.Lop[i], st.Rop[i + 1], st.A[i], st.A[i + 1]
    f = dev.micro_rhs_mals(st.Lrhs[i], st.b[i], st.b[i + 1], st.Rrhs[i + 1])   # sle.py:464-470
    r, n, n2, r3 = L.shape[0], A1.shape[2], A2.shape[2], R.shape[0]
    op = dev.local_op(L, A1, R, A2=A2)
    guess = None
    xi = left if left is not None else st.x[i]
    xj = right if right is not None else st.x[i + 1]
    if xi.dim() == 3 and xj.dim() == 3 and xi.shape[0] == r and xj.shape[2] == r3 and xi.shape[2] == xj.shape[0] \
            and tuple(xi.shape[1:2]) == (n,) and tuple(xj.shape[1:2]) == (n2,) and r * n * n2 * r3 > _local.SMALL_DENSE_LIMIT:
        guess = dev.matmul(xi.reshape(r * n, xi.shape[2]), xj.reshape(xj.shape[0], n2 * r3))
    u = _local.solve_micro(dev, solver, lambda: dev.micro_matrix_mals(L, A1, A2, R), op, f.reshape(r, n, n2, r3), guess,
                           st.cache)
    return u, (r, n, n2, r3)


# ------------------------------------------------------------------------------------------------------------------
# The reference's private per-micro-step helpers under their own names and call conventions (sle.py:194-472; imported by
# name by its TDVP integrators, ode.py:13): host numpy arrays in the stack lists and TT objects in, numpy out, the
# arithmetic on the device through the same C-ABI entry points the sweeps use.  Stateless and therefore slower than the
# resident sweep (one upload / download per call); they exist for drop-in callers and for parity tests that read like the
# reference's.
def _up(dev, a, dtype):
    return dev.upload_many([np.asarray(a)], dtype)[0]


def _helper_dtype(*arrays):
    return torch.complex128 if any(np.iscomplexobj(a) for a in arrays) else torch.float64


def __construct_stack_left_op(i, stack_left_op, operator, solution):
    if i == 0:
        stack_left_op[i] = np.array([1], ndmin=3)                        # sle.py:213
        return
    dev = _device.get_device()
    x, A, L = solution.cores[i - 1][:, :, 0, :], operator.cores[i - 1], stack_left_op[i - 1]
    dt = _helper_dtype(x, A, L)
    stack_left_op[i] = dev.download(dev.stack_left_op(_up(dev, L, dt), _up(dev, x, dt), _up(dev, A, dt)))


def __construct_stack_right_op(i, stack_right_op, operator, solution):
    if i == operator.order - 1:
        stack_right_op[i] = np.array([1], ndmin=3)                       # sle.py:269
        return
    dev = _device.get_device()
    x, A, R = solution.cores[i + 1][:, :, 0, :], operator.cores[i + 1], stack_right_op[i + 1]
    dt = _helper_dtype(x, A, R)
    stack_right_op[i] = dev.download(dev.stack_right_op(_up(dev, R, dt), _up(dev, x, dt), _up(dev, A, dt)))


def __construct_stack_left_rhs(i, stack_left_rhs, right_hand_side, solution):
    if i == 0:
        stack_left_rhs[i] = np.array([1], ndmin=2)                       # sle.py:241
        return
    dev = _device.get_device()
    x, b, S = solution.cores[i - 1][:, :, 0, :], right_hand_side.cores[i - 1][:, :, 0, :], stack_left_rhs[i - 1]
    dt = _helper_dtype(x, b, S)
    stack_left_rhs[i] = dev.download(dev.stack_left_rhs(_up(dev, S, dt), _up(dev, b, dt), _up(dev, x, dt)))


def __construct_stack_right_rhs(i, stack_right_rhs, right_hand_side, solution):
    if i == right_hand_side.order - 1:
        stack_right_rhs[i] = np.array([1], ndmin=2)                      # sle.py:298
        return
    dev = _device.get_device()
    x, b, S = solution.cores[i + 1][:, :, 0, :], right_hand_side.cores[i + 1][:, :, 0, :], stack_right_rhs[i + 1]
    dt = _helper_dtype(x, b, S)
    stack_right_rhs[i] = dev.download(dev.stack_right_rhs(_up(dev, S, dt), _up(dev, b, dt), _up(dev, x, dt)))


def __construct_micro_matrix_als(i, stack_left_op, stack_right_op, operator, solution):
    dev = _device.get_device()
    L, A, R = stack_left_op[i], operator.cores[i], stack_right_op[i]
    dt = _helper_dtype(L, A, R)
    return dev.download(dev.micro_matrix_als(_up(dev, L, dt), _up(dev, A, dt), _up(dev, R, dt)))


def __construct_micro_matrix_mals(i, stack_left_op, stack_right_op, operator, solution):
    dev = _device.get_device()
    L, A1, A2, R = stack_left_op[i], operator.cores[i], operator.cores[i + 1], stack_right_op[i + 1]
    dt = _helper_dtype(L, A1, A2, R)
    return dev.download(dev.micro_matrix_mals(_up(dev, L, dt), _up(dev, A1, dt), _up(dev, A2, dt), _up(dev, R, dt)))


def __construct_micro_rhs_als(i, stack_left_rhs, stack_right_rhs, right_hand_side, solution):
    dev = _device.get_device()
    bL, b, bR = stack_left_rhs[i], right_hand_side.cores[i][:, :, 0, :], stack_right_rhs[i]
    dt = _helper_dtype(bL, b, bR)
    f = dev.download(dev.micro_rhs_als(_up(dev, bL, dt), _up(dev, b, dt), _up(dev, bR, dt)))
    return f.reshape(-1, 1)                                              # sle.py:428 (column vector)


def __construct_micro_rhs_mals(i, stack_left_rhs, stack_right_rhs, right_hand_side, solution):
    dev = _device.get_device()
    bL, b1, b2, bR = (stack_left_rhs[i], right_hand_side.cores[i][:, :, 0, :], right_hand_side.cores[i + 1][:, :, 0, :],
                      stack_right_rhs[i + 1])
    dt = _helper_dtype(bL, b1, b2, bR)
    f = dev.download(dev.micro_rhs_mals(_up(dev, bL, dt), _up(dev, b1, dt), _up(dev, b2, dt), _up(dev, bR, dt)))
    return f.reshape(-1, 1)                                              # sle.py:470
