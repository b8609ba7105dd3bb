"""Device context: one sktt_ctx per (process, CUDA device), torch tensors as the memory owner.

PyTorch is plumbing here (device memory, streams, host<->device copies); every arithmetic step of
the ALS/MALS path goes through the C-ABI of libsktt_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F64, C128, CONJ_ROW, CONJ_COL, Idx2, LocalOp, SkttError

_contexts = {}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dtype_code(t):
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.complex128:
        return C128
    raise TypeError(f"scikit_tt_b200 supports float64 and complex128 only, got {t.dtype}")


class _GuardedLib:
    """Proxy of the ctypes library that makes the context's device current around every C-ABI call.  Only installed when
    one process drives more than one GPU (get_device); the usual one-process-per-GPU layout calls the library directly."""

    def __init__(self, lib, index):
        self._lib, self._index = lib, index

    def __getattr__(self, name):
        fn, idx = getattr(self._lib, name), self._index

        def call(*args):
            if torch.cuda.current_device() != idx:
                with torch.cuda.device(idx):
                    return fn(*args)
            return fn(*args)
        setattr(self, name, call)
        return call


class Device:
    """Thin object wrapper over the C-ABI; methods mirror include/sktt_b200.h one to one."""

    def __init__(self, index=None):
        if not torch.cuda.is_available():
            raise RuntimeError("scikit_tt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.index = torch.cuda.current_device() if index is None else int(index)
        self.device = torch.device("cuda", self.index)
        with torch.cuda.device(self.index):
            self.stream = torch.cuda.current_stream()
        h = C.c_void_p()
        st = self.lib.sktt_ctx_create(self.index, C.c_void_p(self.stream.cuda_stream), C.byref(h))
        if st != 0:
            raise SkttError(st, "sktt_ctx_create failed (is this an sm_100 device?)")
        self.h = h
        self._work = {}
        import os
        if os.environ.get("SKTT_DEBUG"):                       # diagnostics: see sktt_ctx_set_debug
            self.lib.sktt_ctx_set_debug(self.h, int(os.environ["SKTT_DEBUG"]))

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self.lib.sktt_ctx_destroy(self.h)
            self.h = None

    def _check(self, st):
        if st != 0:
            msg = self.lib.sktt_last_error(self.h)
            raise SkttError(st, msg.decode() if msg else "")

    def launches(self):
        return int(self.lib.sktt_launch_count(self.h))

    def set_gemm_mode(self, mode):
        self._check(self.lib.sktt_ctx_set_gemm_mode(self.h, int(mode)))

    def sync(self):
        torch.cuda.synchronize(self.device)

    def set_qr_deferred(self, on):
        """Tall QR / RQ without the Householder launch behind the failure flag of the sketched CholeskyQR (see
        qr_deferred_failures)."""
        self._check(self.lib.sktt_ctx_set_qr_deferred(self.h, int(bool(on))))

    def qr_deferred_failures(self):
        """1 when a sketched CholeskyQR failed since the last call (synchronises, clears the sticky word), else 0."""
        out = C.c_int32(0)
        self._check(self.lib.sktt_qr_deferred_failures(self.h, C.byref(out)))
        return int(out.value)

    def set_debug(self, on):
        self._check(self.lib.sktt_ctx_set_debug(self.h, int(on)))

    def scratch_peek(self, byte_offset, count, ctype=C.c_uint64):
        """Diagnostics: `count` items of `ctype` from the scalar scratch area of the context (synchronises)."""
        buf = (ctype * count)()
        self._check(self.lib.sktt_scratch_peek(self.h, int(byte_offset), C.sizeof(buf), buf))
        return list(buf)

    def work(self, n, dtype, tag="w"):
        """Reusable scratch tensor of at least n elements."""
        key = (tag, dtype)
        t = self._work.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(max(int(n), 1024), dtype=dtype, device=self.device)
            self._work[key] = t
        return t

    def to_device(self, a, dtype=None):
        """numpy array -> contiguous device tensor (float64 / complex128)."""
        a = np.ascontiguousarray(a)
        if dtype is None:
            dtype = torch.complex128 if np.iscomplexobj(a) else torch.float64
        npdt = np.complex128 if dtype == torch.complex128 else np.float64
        return torch.from_numpy(a.astype(npdt, copy=False)).to(self.device)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # -- bulk transfers: DMA straight from / into page-locked host memory.  Inputs that already live in page-locked
    #    memory (e.g. the cores a previous solver call returned) are copied from where they are; pageable inputs are
    #    staged through one reusable pinned buffer by a few host threads (numpy releases the GIL inside copyto).
    PIN_RESULT_MIN_BYTES = 1 << 20

    def _pinned(self, nbytes):
        buf = getattr(self, "_pin", None)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8).pin_memory()
            self._pin = buf
        return buf

    def copy_stream(self):
        """Second stream for device-to-host copies that overlap the sweep."""
        s = getattr(self, "_copy_stream", None)
        if s is None:
            s = self._copy_stream = torch.cuda.Stream(device=self.device)
        return s

    def _pool(self):
        pool = getattr(self, "_threads", None)
        if pool is None:
            import os
            from concurrent.futures import ThreadPoolExecutor
            pool = self._threads = ThreadPoolExecutor(max_workers=max(1, min(8, len(os.sched_getaffinity(0)))))
        return pool

    @staticmethod
    def _page_locked(a, npdt):
        """numpy array that can be handed to the DMA engine as it is: right dtype, contiguous, in page-locked memory."""
        if a.dtype != npdt or not a.flags.c_contiguous or a.size == 0 or not a.flags.writeable:
            return None
        t = torch.from_numpy(a.reshape(-1))
        return t if t.is_pinned() else None

    def pinned_copies(self, arrays):
        """Copies of numpy arrays as views of ONE page-locked block (dtype kept: float64 / complex128 / anything numpy)."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        offs, total = [], 0
        for a in arrays:
            total = (total + 63) // 64 * 64
            offs.append(total)
            total += a.nbytes
        block = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True).numpy()
        out = []
        for a, o in zip(arrays, offs):
            v = block[o:o + a.nbytes].view(a.dtype).reshape(a.shape)
            np.copyto(v, a)
            out.append(v)
        return out

    def upload_many(self, arrays, dtype):
        """List of numpy arrays -> list of contiguous device tensors of `dtype` (float64 / complex128)."""
        npdt = np.complex128 if dtype == torch.complex128 else np.float64
        item = np.dtype(npdt).itemsize
        arrays = [np.asarray(a) for a in arrays]
        sizes = [int(a.size) for a in arrays]
        total = sum(sizes)
        if total == 0:
            return [self.empty(a.shape, dtype) for a in arrays]
        devbuf = torch.empty(total, dtype=dtype, device=self.device)
        offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
        direct = [self._page_locked(a, npdt) if n * item >= 65536 else None for a, n in zip(arrays, sizes)]
        staged = [k for k, (t, n) in enumerate(zip(direct, sizes)) if t is None and n > 0]
        for k, t in enumerate(direct):
            if t is not None:
                devbuf[offs[k]:offs[k + 1]].copy_(t if dtype == t.dtype else t.view(dtype), non_blocking=True)
        if staged:
            nst = sum(sizes[k] for k in staged)
            self.sync()                                        # the staging buffer may still feed an earlier copy
            host = self._pinned(nst * item)[: nst * item].view(dtype)
            hnp = host.numpy()
            soff, pos = {}, 0
            for k in staged:
                soff[k] = pos
                pos += sizes[k]

            def stage(k):
                np.copyto(hnp[soff[k]:soff[k] + sizes[k]].reshape(arrays[k].shape), arrays[k], casting="same_kind")
            if nst * item >= (4 << 20) and len(staged) > 1:
                list(self._pool().map(stage, staged))
            else:
                for k in staged:
                    stage(k)
            for k in staged:
                devbuf[offs[k]:offs[k + 1]].copy_(host[soff[k]:soff[k] + sizes[k]], non_blocking=True)
        return [devbuf[offs[k]:offs[k + 1]].view(arrays[k].shape) for k in range(len(arrays))]

    def download_many(self, tensors):
        """List of device tensors (one dtype) -> list of numpy arrays, one synchronisation.  Results of at least
        PIN_RESULT_MIN_BYTES are views of ONE page-locked block from torch's caching host allocator (no second host
        copy, no page faults; the block returns to the allocator when the last view dies), so they can be fed back
        to the next solver call without staging.  Smaller results are ordinary numpy arrays."""
        if not tensors:
            return []
        dtype = tensors[0].dtype
        item = tensors[0].element_size()
        sizes = [t.numel() for t in tensors]
        total = sum(sizes)
        big = total * item >= self.PIN_RESULT_MIN_BYTES
        host = torch.empty(max(total, 1), dtype=dtype, pin_memory=True) if big else \
            self._pinned(max(total, 1) * item)[: max(total, 1) * item].view(dtype)
        off = 0
        for t, n in zip(tensors, sizes):
            if n:
                host[off:off + n].copy_(t.contiguous().reshape(-1), non_blocking=True)
            off += n
        self.sync()
        hnp = host.numpy()
        out, off = [], 0
        for t, n in zip(tensors, sizes):
            v = hnp[off:off + n].reshape(tuple(t.shape))
            out.append(v if big else v.copy())
            off += n
        return out

    def download(self, t):
        return self.download_many([t])[0]

    # ------------------------------------------------------------------ stacks
    def stack_left_op(self, Lst, x, A, conj_mode=CONJ_ROW):
        r, n, r2 = x.shape
        R, m, n_, R2 = A.shape
        out = self.empty((r2, R2, r2), x.dtype)
        w = self.work(self.lib.sktt_stack_op_work(r, R, m, n, r2, R2), x.dtype)
        self._check(self.lib.sktt_stack_left_op(self.h, dtype_code(x), r, R, m, n_, r2, R2, _ptr(Lst), _ptr(x), _ptr(A),
                                                _ptr(out), _ptr(w), conj_mode))
        return out

    def stack_right_op(self, Rst, x, A):
        r, n, r2 = x.shape
        R, m, n_, R2 = A.shape
        out = self.empty((r, R, r), x.dtype)
        w = self.work(self.lib.sktt_stack_op_work(r, R, m, n, r2, R2), x.dtype)
        self._check(self.lib.sktt_stack_right_op(self.h, dtype_code(x), r, R, m, n_, r2, R2, _ptr(Rst), _ptr(x), _ptr(A),
                                                 _ptr(out), _ptr(w)))
        return out

    def stack_left_rhs(self, bL, b, x):
        p, m, p2 = b.shape
        r, _, r2 = x.shape
        out = self.empty((p2, r2), x.dtype)
        w = self.work(r * m * p2, x.dtype)
        self._check(self.lib.sktt_stack_left_rhs(self.h, dtype_code(x), p, r, m, p2, r2, _ptr(bL), _ptr(b), _ptr(x),
                                                 _ptr(out), _ptr(w)))
        return out

    def stack_right_rhs(self, bR, b, x):
        p, m, p2 = b.shape
        r, _, r2 = x.shape
        out = self.empty((p, r), x.dtype)
        w = self.work(r * m * p2, x.dtype)
        self._check(self.lib.sktt_stack_right_rhs(self.h, dtype_code(x), p, r, m, p2, r2, _ptr(bR), _ptr(b), _ptr(x),
                                                  _ptr(out), _ptr(w)))
        return out

    # ------------------------------------------------------------------ micro systems
    def micro_matrix_als(self, Lst, A, Rst):
        r = Lst.shape[0]
        r2 = Rst.shape[0]
        R, m, n, R2 = A.shape
        out = self.empty((r * m * r2, r * n * r2), A.dtype)
        w = self.work(R * m * n * r2 * r2, A.dtype)
        self._check(self.lib.sktt_micro_matrix_als(self.h, dtype_code(A), r, R, m, n, r2, R2, _ptr(Lst), _ptr(A),
                                                   _ptr(Rst), _ptr(out), _ptr(w)))
        return out

    def micro_matvec_als(self, Lst, A, Rst, v):
        r = Lst.shape[0]
        r2 = Rst.shape[0]
        R, m, n, R2 = A.shape
        y = self.empty((r, m, r2), v.dtype)
        w = self.work(self.lib.sktt_stack_op_work(r, R, m, n, r2, R2), v.dtype)
        self._check(self.lib.sktt_micro_matvec_als(self.h, dtype_code(v), r, R, m, n, r2, R2, _ptr(Lst), _ptr(A),
                                                   _ptr(Rst), _ptr(v), _ptr(y), _ptr(w)))
        return y

    def micro_matrix_mals(self, Lst, A1, A2, Rst):
        r = Lst.shape[0]
        r3 = Rst.shape[0]
        R, m, n, R2 = A1.shape
        _, m2, n2, R3 = A2.shape
        out = self.empty((r * m * m2 * r3, r * n * n2 * r3), A1.dtype)
        w = self.work(self.lib.sktt_micro_matrix_mals_work(r, R, m, n, R2, m2, n2, R3, r3), A1.dtype)
        self._check(self.lib.sktt_micro_matrix_mals(self.h, dtype_code(A1), r, R, m, n, R2, m2, n2, R3, r3, _ptr(Lst),
                                                    _ptr(A1), _ptr(A2), _ptr(Rst), _ptr(out), _ptr(w)))
        return out

    def micro_matvec_mals(self, Lst, A1, A2, Rst, v):
        r = Lst.shape[0]
        r3 = Rst.shape[0]
        R, m, n, R2 = A1.shape
        _, m2, n2, R3 = A2.shape
        y = self.empty((r, m, m2, r3), v.dtype)
        w = self.work(self.lib.sktt_micro_matvec_mals_work(r, R, m, n, R2, m2, n2, R3, r3), v.dtype)
        self._check(self.lib.sktt_micro_matvec_mals(self.h, dtype_code(v), r, R, m, n, R2, m2, n2, R3, r3, _ptr(Lst),
                                                    _ptr(A1), _ptr(A2), _ptr(Rst), _ptr(v), _ptr(y), _ptr(w)))
        return y

    def micro_rhs_als(self, bL, b, bR):
        p, m, p2 = b.shape
        r, r2 = bL.shape[1], bR.shape[1]
        f = self.empty((r, m, r2), b.dtype)
        w = self.work(r * m * p2, b.dtype)
        self._check(self.lib.sktt_micro_rhs_als(self.h, dtype_code(b), p, r, m, p2, r2, _ptr(bL), _ptr(b), _ptr(bR),
                                                _ptr(f), _ptr(w)))
        return f

    def micro_rhs_mals(self, bL, b1, b2, bR):
        p, m, p2 = b1.shape
        _, m2, p3 = b2.shape
        r, r3 = bL.shape[1], bR.shape[1]
        f = self.empty((r, m, m2, r3), b1.dtype)
        w = self.work(r * m * p2 + r * m * m2 * p3, b1.dtype)
        self._check(self.lib.sktt_micro_rhs_mals(self.h, dtype_code(b1), p, r, m, p2, m2, p3, r3, _ptr(bL), _ptr(b1),
                                                 _ptr(b2), _ptr(bR), _ptr(f), _ptr(w)))
        return f

    def rank1_update(self, M, t, shift):
        self._check(self.lib.sktt_rank1_update(self.h, dtype_code(M), M.shape[0], float(shift), _ptr(t), _ptr(M)))

    # ------------------------------------------------------------------ dense solves
    def lu_factor(self, M):
        """In-place LU of the square row-major matrix M; returns (ipiv tensor, info)."""
        N = M.shape[0]
        ipiv = torch.empty(2 * N, dtype=torch.int32, device=self.device)
        info = C.c_int(0)
        self._check(self.lib.sktt_lu_factor(self.h, dtype_code(M), N, _ptr(M), _ptr(ipiv), C.byref(info)))
        return ipiv, info.value

    def lu_solve(self, LU, ipiv, B):
        """Overwrites B ([N] or [N, nrhs]) with the solution."""
        N = LU.shape[0]
        nrhs = 1 if B.dim() == 1 else B.shape[1]
        self._check(self.lib.sktt_lu_solve(self.h, dtype_code(LU), N, nrhs, _ptr(LU), _ptr(ipiv), _ptr(B)))
        return B

    def solve_fused(self, M, f, want_pivots=False):
        """np.linalg.solve of one fp64 system in ONE cooperative launch (csrc/lu_fused.cu); M is overwritten by L \\ U.
        Returns the solution (and the 0-based pivot rows)."""
        N = M.shape[0]
        x = f.reshape(-1).clone()
        ipiv = torch.empty(N, dtype=torch.int32, device=self.device)
        info = torch.empty(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.sktt_lu_solve_fused(self.h, N, _ptr(M), _ptr(x), _ptr(ipiv), _ptr(info)))
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return (x, ipiv) if want_pivots else x

    def solve(self, M, f):
        """np.linalg.solve semantics (M destroyed): raises numpy.linalg.LinAlgError when singular."""
        if M.dtype == torch.float64 and M.shape[0] <= self.fused_lu_max_n() and f.numel() == M.shape[0]:
            return self.solve_fused(M, f)
        ipiv, info = self.lu_factor(M)
        if info != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        x = f.reshape(-1).clone()
        return self.lu_solve(M, ipiv, x)

    def fused_lu_max_n(self):
        n = getattr(self, "_fused_lu_max_n", None)
        if n is None:
            n = self._fused_lu_max_n = int(self.lib.sktt_lu_fused_max_n())
        return n

    def chol_factor(self, M):
        info = C.c_int(0)
        self._check(self.lib.sktt_chol_factor(self.h, dtype_code(M), M.shape[0], _ptr(M), C.byref(info)))
        return info.value

    def chol_solve(self, Lfac, B):
        N = Lfac.shape[0]
        nrhs = 1 if B.dim() == 1 else B.shape[1]
        self._check(self.lib.sktt_chol_solve(self.h, dtype_code(Lfac), N, nrhs, _ptr(Lfac), _ptr(B)))
        return B

    def chol_trsm(self, Lfac, B, backward=False):
        """L^-1 B (backward=False) or L^-H B (backward=True); returns a new tensor."""
        N = Lfac.shape[0]
        X = B.contiguous().clone()
        nrhs = 1 if X.dim() == 1 else X.shape[1]
        self._check(self.lib.sktt_chol_trsm(self.h, dtype_code(Lfac), N, nrhs, _ptr(Lfac), _ptr(X), int(bool(backward))))
        return X

    # ------------------------------------------------------------------ Krylov
    def local_op(self, Lst, A1, Rst, A2=None, prepare=False):
        op = LocalOp()
        op.sites = 1 if A2 is None else 2
        op.r = Lst.shape[0]
        op.R, op.m, op.n, op.R2 = A1.shape
        if A2 is None:
            op.m2 = op.n2 = 1
            op.R3 = op.R2
        else:
            _, op.m2, op.n2, op.R3 = A2.shape
        op.r3 = Rst.shape[0]
        op.Lst, op.A1, op.Rst = Lst.data_ptr(), A1.data_ptr(), Rst.data_ptr()
        op.A2 = A2.data_ptr() if A2 is not None else None
        op.image = None
        op._keep = (Lst, A1, A2, Rst)
        if prepare:
            self.prepare_local_op(op)
        return op

    def prepare_local_op(self, op):
        """Build the TMA tile images of a one-site local operator (no-op for shapes the fused matvec does not cover)."""
        if getattr(op, "_prepared", False):
            return op
        op._prepared = True
        dt = op._keep[0].dtype
        n = self.lib.sktt_local_op_image_size(self.h, dtype_code(op._keep[0]), C.byref(op))
        if n > 0:
            op._image = self.empty((n,), dt)
            self._check(self.lib.sktt_local_op_prepare(self.h, dtype_code(op._keep[0]), C.byref(op), _ptr(op._image)))
        return op

    def local_matvec(self, op, v):
        Lst, A1, A2, Rst = op._keep
        shape = (op.r, op.m, op.r3) if A2 is None else (op.r, op.m, op.m2, op.r3)
        y = self.empty(shape, v.dtype)
        w = self.work(self.lib.sktt_local_matvec_work(C.byref(op)), v.dtype)
        self._check(self.lib.sktt_local_matvec(self.h, dtype_code(v), C.byref(op), _ptr(v), _ptr(y), _ptr(w)))
        return y

    def local_matvec_tiled(self, op, vt, yt=None):
        """The Krylov inner step: y = M v on tiled-layout vectors of a prepared operator (two kernel launches)."""
        n = self.lib.sktt_local_op_tiled_len(self.h, dtype_code(vt), C.byref(op))
        if n <= 0:
            raise ValueError("operator is not prepared for the tiled matvec")
        if yt is None:
            yt = torch.zeros(n, dtype=vt.dtype, device=self.device)
        w = self.work(self.lib.sktt_local_matvec_work(C.byref(op)), vt.dtype)
        self._check(self.lib.sktt_local_matvec_tiled(self.h, dtype_code(vt), C.byref(op), _ptr(vt), _ptr(yt), _ptr(w)))
        return yt

    def local_matvec_tiled_repeat(self, op, vt, yt, reps):
        """`reps` matvecs of a prepared operator inside one persistent cooperative launch (roofline timing)."""
        w = self.work(self.lib.sktt_local_matvec_work(C.byref(op)), vt.dtype)
        self._check(self.lib.sktt_local_matvec_tiled_repeat(self.h, dtype_code(vt), C.byref(op), _ptr(vt), _ptr(yt), _ptr(w),
                                                            int(reps)))
        return yt

    def tiled_len(self, op):
        return int(self.lib.sktt_local_op_tiled_len(self.h, dtype_code(op._keep[0]), C.byref(op)))

    def krylov_solve(self, op, f, u, method="cg", tol=1e-13, max_iters=5000, restart=40):
        """Solves the micro system matrix-free; u (initial guess) is overwritten. Returns (iters, relres)."""
        meth = {"cg": 0, "gmres": 1}[method]
        nwork = self.lib.sktt_krylov_work(C.byref(op), meth, restart)
        w = self.work(nwork, f.dtype, tag="krylov")
        iters, relres = C.c_int(0), C.c_double(0.0)
        st = self.lib.sktt_krylov_solve(self.h, dtype_code(f), C.byref(op), meth, restart, _ptr(f), _ptr(u), float(tol),
                                        int(max_iters), _ptr(w), C.byref(iters), C.byref(relres))
        return st, iters.value, relres.value

    def krylov_solve_refined(self, op, f, u, tol=1e-14, max_iters=20000, max_cycles=5):
        """The whole matrix-free micro solve in one C call (prepared one-site f64 operators only): warm start, CG, true
        residual checks with restarts.  u is overwritten.  Returns (status, iterations, true relres, cycles)."""
        nwork = self.lib.sktt_krylov_work(C.byref(op), 0, 0)
        w = self.work(nwork, f.dtype, tag="krylov")
        iters, relres, cycles = C.c_int(0), C.c_double(0.0), C.c_int(0)
        st = self.lib.sktt_krylov_solve_refined(self.h, dtype_code(f), C.byref(op), _ptr(f), _ptr(u), float(tol),
                                                int(max_iters), int(max_cycles), _ptr(w), C.byref(iters), C.byref(relres),
                                                C.byref(cycles))
        return st, iters.value, relres.value, cycles.value

    def krylov_solve_refined_async(self, op, f, u, result, tol=1e-14, max_cycles=5):
        """Deferred form of krylov_solve_refined: queued without a host synchronisation; `result` (device, 4 doubles)
        receives (iterations, true relres, cycles, status).  Returns False when the operator is not covered."""
        nwork = self.lib.sktt_krylov_work(C.byref(op), 0, 0)
        w = self.work(nwork, f.dtype, tag="krylov")
        st = self.lib.sktt_krylov_solve_refined_async(self.h, dtype_code(f), C.byref(op), _ptr(f), _ptr(u), float(tol),
                                                      int(max_cycles), _ptr(w), _ptr(result))
        if st == -1:                                           # SKTT_ERR_ARG: operator not covered by the persistent kernel
            return False
        self._check(st)
        return True

    # ------------------------------------------------------------------ orthonormalisation
    # Tall matrices with 64 < n <= 512 columns (the unfoldings of BASELINE config 4: 4096 x 128 / 256): the sketched
    # CholeskyQR kernel takes at most 64 columns and the Householder kernel behind it pays one grid barrier and a
    # G x n reduction per reflector (measured 3.4 / 8.1 ms at 128 / 256 columns: half of a config-4 sweep).  Block classical
    # Gram-Schmidt with re-orthogonalisation (BCGS2) over blocks of 64 columns instead: every block is projected against
    # the finished ones by two contractions and orthonormalised by the 64-column kernel, twice -- the second pass restores
    # orthogonality to rounding level, and an accurate intra-block QR (the sketched CholeskyQR is Householder-grade, DESIGN.md
    # 4.4) is what BCGS2 needs.  The economic Q of a full-rank matrix is unique up to the sign of each column, so the
    # result equals Householder's up to signs (an exact symmetry of the sweeps); R is not formed (the sweeps discard it).
    QR_BLOCK = 64
    QR_BLOCKED_MAX_N = 512
    blocked_qr = True

    def _qr_blocked(self, A):
        m, n = A.shape
        big = 1 << 40
        Q = self.empty((m, n), A.dtype)
        for c0 in range(0, n, self.QR_BLOCK):
            nb = min(self.QR_BLOCK, n - c0)
            Xb = A[:, c0:c0 + nb].contiguous()
            for _ in range(2):
                if c0 > 0:
                    H = self.empty((c0, nb), A.dtype)                     # H = Q_done^T X_b
                    self.gemm2(c0, nb, m, Q, (big, 0, 1), (big, 0, n), Xb, (big, 0, nb), (big, 0, 1), H, (big, 0, nb), (big, 0, 1))
                    self.gemm2(m, nb, c0, Q, (big, 0, n), (big, 0, 1), H, (big, 0, nb), (big, 0, 1), Xb, (big, 0, nb), (big, 0, 1),
                               alpha=(-1.0, 0.0), beta=(1.0, 0.0))        # X_b -= Q_done H
                Qb = self.empty((m, nb), A.dtype)
                self._check(self.lib.sktt_qr_left(self.h, dtype_code(A), m, nb, _ptr(Xb), _ptr(Qb), C.c_void_p(0), C.c_void_p(0)))
                Xb = Qb
            Q[:, c0:c0 + nb].copy_(Xb)
        return Q

    def _wants_blocked_qr(self, m, n, dtype, want_r):
        return (self.blocked_qr and not want_r and dtype == torch.float64 and self.QR_BLOCK < n <= self.QR_BLOCKED_MAX_N
                and m >= 8 * self.QR_BLOCK and m >= 2 * n)

    def qr(self, A, want_r=False):
        m, n = A.shape
        if self._wants_blocked_qr(m, n, A.dtype, want_r):
            return self._qr_blocked(A)
        k = min(m, n)
        Q = self.empty((m, k), A.dtype)
        R = self.empty((k, n), A.dtype) if want_r else None
        self._check(self.lib.sktt_qr_left(self.h, dtype_code(A), m, n, _ptr(A), _ptr(Q), _ptr(R), C.c_void_p(0)))
        return (Q, R) if want_r else Q

    def rq(self, A, want_r=False):
        m, n = A.shape
        if self._wants_blocked_qr(n, m, A.dtype, want_r):
            # A = R Q with the conventions of LAPACK's gerqf: C = (J A J)^T = Qc Rc  =>  Q = J Qc^T J (csrc/qr.cu, sktt_rq_right)
            Cm = torch.flip(A, dims=[0, 1]).t().contiguous()
            return torch.flip(self._qr_blocked(Cm).t(), dims=[0, 1]).contiguous()
        k = min(m, n)
        Q = self.empty((k, n), A.dtype)
        R = self.empty((m, k), A.dtype) if want_r else None
        self._check(self.lib.sktt_rq_right(self.h, dtype_code(A), m, n, _ptr(A), _ptr(Q), _ptr(R), C.c_void_p(0)))
        return (R, Q) if want_r else Q

    def svd(self, A, threshold=0.0, max_rank=0):
        """Economic SVD + the reference's rank rule.  Returns (U, S, Vh, rank) untruncated + kept rank."""
        m, n = A.shape
        k = min(m, n)
        U = self.empty((m, k), A.dtype)
        S = self.empty((k,), torch.float64)
        Vh = self.empty((k, n), A.dtype)
        w = self.work(self.lib.sktt_svd_work(m, n), A.dtype, tag="svd")
        rank, sweeps = C.c_int(0), C.c_int(0)
        mr = 0 if (max_rank is None or max_rank == np.inf or max_rank <= 0) else int(max_rank)
        self._check(self.lib.sktt_svd_truncate(self.h, dtype_code(A), m, n, _ptr(A), _ptr(U), _ptr(S), _ptr(Vh),
                                               float(threshold), mr, _ptr(w), C.byref(rank), C.byref(sweeps)))
        self.last_svd_sweeps = sweeps.value
        return U, S, Vh, rank.value

    SVD_DIRECT_LIMIT = 1024     # min(m, n) up to which the full Jacobi SVD is the route whatever max_rank says
    svd_topk_stats = {"calls": 0, "iterations": 0, "fallbacks": 0}

    def svd_truncated(self, A, threshold=0.0, max_rank=0):
        """The truncated SVD of sle.py:603-614 / :626-639 for a super-core whose full SVD is out of reach (C3 / C4 two-site
        blocks: 4096 x 4096) but of which only `max_rank` << min(m, n) singular triplets are kept: subspace iteration for the
        dominant left subspace of dimension max_rank + oversampling (products through the DMMA contraction engine,
        orthonormalisation by the QR kernels), then the accurate one-sided Jacobi SVD of the small projected matrix
        Q^H A.  Iterated until the kept singular values stop moving (1e-12 relative); falls back to the full SVD when they do
        not settle (flat spectrum at the cut).  Same return convention as `svd`: (U, S, Vh, rank) with at least `rank`
        columns / rows."""
        m, n = A.shape
        mr = 0 if (max_rank is None or max_rank == np.inf or max_rank <= 0) else int(max_rank)
        k = min(m, n)
        if mr == 0 or k <= self.SVD_DIRECT_LIMIT or 4 * mr > k:
            return self.svd(A, threshold=threshold, max_rank=max_rank)
        self.svd_topk_stats["calls"] += 1
        l = min(k, mr + max(16, mr // 2))
        g = torch.Generator(device=self.device).manual_seed(20240229)
        Om = torch.randn((n, l), dtype=torch.float64, device=self.device, generator=g)
        if A.dtype == torch.complex128:
            Om = self.widen(Om)
        Q = self.qr(self.matmul(A, Om))                            # [m, l]
        S_prev = None
        for it in range(12):
            Z = self.qr(self.matmul(A, Q, opa='C'))                # [n, l]: A^H Q, orthonormalised
            Q = self.qr(self.matmul(A, Z))                         # [m, l]
            B = self.matmul(Q, A, opa='C')                         # [l, n] = Q^H A
            Ub, S, Vh, rank = self.svd(B, threshold=threshold, max_rank=mr)
            self.svd_topk_stats["iterations"] += 1
            Sh = S[:mr].cpu().numpy()
            if S_prev is not None and np.all(np.abs(Sh - S_prev) <= 1e-12 * Sh[0]):
                U = self.matmul(Q, Ub)                             # [m, l]
                return U, S, Vh, rank
            S_prev = Sh
        self.svd_topk_stats["fallbacks"] += 1
        return self.svd(A, threshold=threshold, max_rank=max_rank)

    # ------------------------------------------------------------------ eigen
    def eigh(self, M):
        N = M.shape[0]
        W = self.empty((N,), torch.float64)
        V = self.empty((N, N), M.dtype)
        sweeps = C.c_int(0)
        self._check(self.lib.sktt_eigh_jacobi(self.h, dtype_code(M), N, _ptr(M), _ptr(W), _ptr(V), C.byref(sweeps)))
        return W, V

    def eig_shift_invert(self, M, sigma, k, B=None, ncv=None, tol=1e-13, max_restarts=20):
        """k eigenpairs of (M, B) closest to sigma; M is destroyed. Returns (lam[k], vecs[N,k]) complex128.

        The reference's `lin.eig` always returns exact pairs and `splin.eigs` raises ArpackNoConvergence; here a run that
        ends with fewer than k converged Ritz pairs (clustered spectrum near sigma, exhausted Krylov space) is repeated
        once with the largest Krylov dimension and more restarts, and raises numpy.linalg.LinAlgError if that fails too --
        unconverged pairs are never handed to the sweep.  Up to EIG_EXACT_MAX_N unknowns the repeat is the exact one: a
        Krylov space of full dimension."""
        N = M.shape[0]
        if ncv is None:
            ncv = min(N, max(2 * k + 1, 20))
        ncv = min(int(ncv), 64, N)
        keep = M.clone() if ncv < N else None
        lam = self.empty((k,), torch.complex128)
        vecs = self.empty((N, k), torch.complex128)
        nconv = C.c_int(0)

        def run(mat, ncv, max_restarts):
            w = self.work(self.lib.sktt_eig_si_work(N, k, ncv), torch.complex128, tag="eig")
            self._check(self.lib.sktt_eig_shift_invert(self.h, dtype_code(mat), N, _ptr(mat), _ptr(B), float(sigma), k, ncv,
                                                       float(tol), int(max_restarts), _ptr(lam), _ptr(vecs), _ptr(w),
                                                       C.byref(nconv)))
        run(M, ncv, max_restarts)
        if nconv.value < k and keep is not None and N <= self.EIG_EXACT_MAX_N:
            # Many eigenvalues at (nearly) the same distance from sigma -- restarted Arnoldi crawls.  The Krylov space of
            # dimension N is the whole space: the projected Hessenberg matrix then carries the full spectrum, as lin.eig does
            self.eig_exact_fallbacks = getattr(self, "eig_exact_fallbacks", 0) + 1
            run(keep, N, 0)
        elif nconv.value < k and keep is not None:
            run(keep, min(64, N), 5 * max_restarts)
        self.last_eig_nconv = nconv.value
        if nconv.value < k:
            raise np.linalg.LinAlgError(f"shift-invert Arnoldi: {nconv.value} of {k} eigenpairs near sigma = {sigma} converged "
                                        f"(N = {N})")
        return lam, vecs

    EIG_EXACT_MAX_N = 1024

    # ------------------------------------------------------------------ batched small systems (leading batch dimension)
    def batch_stack_left_op(self, Lst, x, A, conj_mode=CONJ_ROW):
        """Lst [B, r, R, r], x [B, r, n, r2], A [B, R, m, n, R2] -> [B, r2, R2, r2]."""
        B, r, n, r2 = x.shape
        _, R, m, n_, R2 = A.shape
        out = self.empty((B, r2, R2, r2), x.dtype)
        w = self.work(B * (R * r * n * r2 + r * m * r2 * R2), x.dtype, tag="bstack")
        self._check(self.lib.sktt_batch_stack_left_op(self.h, dtype_code(x), B, r, R, m, n_, r2, R2, _ptr(Lst), _ptr(x),
                                                      _ptr(A), _ptr(out), _ptr(w), conj_mode))
        return out

    def batch_stack_right_op(self, Rst, x, A):
        """Rst [B, r2, R2, r2], x [B, r, n, r2], A [B, R, m, n, R2] -> [B, r, R, r]."""
        B, r, n, r2 = x.shape
        _, R, m, n_, R2 = A.shape
        out = self.empty((B, r, R, r), x.dtype)
        w = self.work(B * (R * r * n * r2 + r * m * r2 * R2), x.dtype, tag="bstack")
        self._check(self.lib.sktt_batch_stack_right_op(self.h, dtype_code(x), B, r, R, m, n_, r2, R2, _ptr(Rst), _ptr(x),
                                                       _ptr(A), _ptr(out), _ptr(w)))
        return out

    def batch_micro_matrix_als(self, Lst, A, Rst):
        """-> [B, r m r2, r n r2] dense micro matrices."""
        B, r = Lst.shape[0], Lst.shape[1]
        r2 = Rst.shape[1]
        _, R, m, n, R2 = A.shape
        out = self.empty((B, r * m * r2, r * n * r2), A.dtype)
        w = self.work(B * R * m * n * r2 * r2, A.dtype, tag="bmicro")
        self._check(self.lib.sktt_batch_micro_matrix_als(self.h, dtype_code(A), B, r, R, m, n, r2, R2, _ptr(Lst), _ptr(A),
                                                         _ptr(Rst), _ptr(out), _ptr(w)))
        return out

    BATCH_EIG_MAX_N = 1024

    def batch_eig_shift_invert(self, M, sigma, k, ncv=None, tol=1e-12, max_restarts=5):
        """k eigenpairs closest to sigma of each M[b] ([B, N, N], float64 or complex128), one CTA per system, one launch.
        Returns (lam [B, k], vecs [B, N, k], status [B, 3] float64: converged pairs, info, worst relative residual estimate)
        -- all on the device, nothing is synchronised."""
        B, N, _ = M.shape
        if ncv is None:
            ncv = min(N, max(2 * k + 1, 20))
        ncv = min(int(ncv), 32, N)
        lam = self.empty((B, k), torch.complex128)
        vecs = self.empty((B, N, k), torch.complex128)
        status = torch.empty((B, 2), dtype=torch.int32, device=self.device)
        relres = torch.empty((B,), dtype=torch.float64, device=self.device)
        w = self.work(self.lib.sktt_batch_eig_work(B, N, k, ncv), torch.complex128, tag="beig")
        self._check(self.lib.sktt_batch_eig_shift_invert(self.h, dtype_code(M), B, N, _ptr(M), float(sigma), k, ncv,
                                                         float(tol), int(max_restarts), _ptr(lam), _ptr(vecs), _ptr(w),
                                                         _ptr(status), _ptr(relres)))
        return lam, vecs, torch.cat([status.to(torch.float64), relres[:, None]], dim=1)

    def batch_svd_left(self, src, P, Q, keep, fi, fj, conj_in, out, so_i, so_t, conj_out):
        """First `keep` left singular vectors of the P x Q blocks F(i, j) = op(src[b].flat[fi(i) + fj(j)]), written to
        out[b].flat[i * so_i + t * so_t]; src / out are [B, ...] contiguous complex128 tensors."""
        B = src.shape[0]
        mk = lambda t: Idx2(*t)
        self._check(self.lib.sktt_batch_svd_left(self.h, B, P, Q, keep, _ptr(src), src[0].numel(), mk(fi), mk(fj),
                                                 int(conj_in), _ptr(out), out[0].numel(), so_i, so_t, int(conj_out)))
        return out

    def tt_matmul_core(self, A, B):
        """A [P, m, K, Q] (x) B [S, K, n, T] -> [P * S, m, n, Q * T]: one core of TT.__matmul__ (tensor_train.py:422-503)."""
        A, B = common_dtype(A.contiguous(), B.contiguous())
        P, m, K, Q = A.shape
        S, K2, n, T = B.shape
        if K != K2:
            raise ValueError("Dimensions do not match.")
        out = self.empty((P * S, m, n, Q * T), A.dtype)
        self._check(self.lib.sktt_tt_matmul_core(self.h, dtype_code(A), P, m, K, Q, S, n, T, _ptr(A), _ptr(B), _ptr(out)))
        return out

    # ------------------------------------------------------------------ alternating ridge regression (fp64)
    def arr_stack_left(self, L, Phi, core):
        r, n, r2 = core.shape
        m = Phi.shape[1]
        out = self.empty((r2, m), torch.float64)
        self._check(self.lib.sktt_arr_stack(self.h, 0, r, n, r2, m, _ptr(L), _ptr(Phi), _ptr(core), _ptr(out)))
        return out

    def arr_stack_right(self, R, Phi, core):
        r, n, r2 = core.shape
        m = Phi.shape[1]
        out = self.empty((r, m), torch.float64)
        self._check(self.lib.sktt_arr_stack(self.h, 1, r, n, r2, m, _ptr(R), _ptr(Phi), _ptr(core), _ptr(out)))
        return out

    def arr_micro_matrix(self, L, Phi, R):
        r, n, r2, m = L.shape[0], Phi.shape[0], R.shape[0], Phi.shape[1]
        out = self.empty((r * n * r2, m), torch.float64)
        self._check(self.lib.sktt_arr_micro_matrix(self.h, r, n, r2, m, _ptr(L), _ptr(Phi), _ptr(R), _ptr(out)))
        return out

    def lstsq_gelss(self, A, b, rcond):
        """Minimum-norm least-squares solution of A x = b (A [m, N], b [m]) with the singular values below rcond * s_0
        dropped -- scipy.linalg.lstsq(A, b, cond=rcond, lapack_driver='gelss') (regression.py:419) on the device: one-sided
        Jacobi SVD, two thin products, the cut in between."""
        U, S, Vh, _ = self.svd(A.contiguous())
        t = self.matmul(U, b.reshape(-1, 1), opa='C').reshape(-1)
        self._check(self.lib.sktt_pinv_scale(self.h, S.numel(), _ptr(S), float(rcond), _ptr(t)))
        return self.matmul(Vh, t.reshape(-1, 1), opa='C').reshape(-1)

    def expm_small(self, H, c):
        """exp(c * H) of a small dense complex128 matrix (m <= 64) on the device."""
        m = H.shape[0]
        H = H.contiguous() if H.dtype == torch.complex128 else self.widen(H.contiguous())
        E = self.empty((m, m), torch.complex128)
        c = complex(c)
        self._check(self.lib.sktt_expm_small(self.h, m, _ptr(H), c.real, c.imag, _ptr(E)))
        return E

    # ------------------------------------------------------------------ misc
    def gemm2(self, M, N, K, A, am, ak, B, bk, bn, Cmat, cm, cn, conjA=0, conjB=0, alpha=(1.0, 0.0), beta=(0.0, 0.0)):
        al = (C.c_double * 2)(*alpha)
        be = (C.c_double * 2)(*beta)
        mk = lambda t: Idx2(*t)
        self._check(self.lib.sktt_gemm2(self.h, dtype_code(A), M, N, K, al, _ptr(A), mk(am), mk(ak), conjA, _ptr(B),
                                        mk(bk), mk(bn), conjB, be, _ptr(Cmat), mk(cm), mk(cn)))

    def matmul(self, A, B, opa='N', opb='N', out=None):
        """op(A) @ op(B) for 2-D row-major device tensors; op in {'N', 'T', 'C'} resolved in the tile loaders."""
        big = 1 << 40
        am, ak = ((big, 0, A.shape[1]), (big, 0, 1)) if opa == 'N' else ((big, 0, 1), (big, 0, A.shape[1]))
        bk, bn = ((big, 0, B.shape[1]), (big, 0, 1)) if opb == 'N' else ((big, 0, 1), (big, 0, B.shape[1]))
        M, K = (A.shape[0], A.shape[1]) if opa == 'N' else (A.shape[1], A.shape[0])
        K2, N = (B.shape[0], B.shape[1]) if opb == 'N' else (B.shape[1], B.shape[0])
        if K != K2:
            raise ValueError(f"matmul: inner extents differ ({K} vs {K2})")
        if out is None:
            out = self.empty((M, N), A.dtype)
        self.gemm2(M, N, K, A, am, ak, B, bk, bn, out, (big, 0, N), (big, 0, 1), conjA=int(opa == 'C'), conjB=int(opb == 'C'))
        return out

    # ------------------------------------------------------------------ gauge products of the sweep (warm starts)
    def gauge_factor(self, q, u, tall):
        """Triangular factor the reference throws away (sle.py:525, :541), recovered from the orthonormal factor:
        tall: q [L, k], u [L, r2] -> q^H u [k, r2];  else: q [k, L], u [r, L] -> u q^H [r, k]."""
        if q.dtype != torch.float64 or max(q.shape[1 if tall else 0], u.shape[1 if tall else 0]) > 64:
            return self.matmul(q, u, opa='C') if tall else self.matmul(u, q, opb='C')
        if tall:
            L, k = q.shape
            r2 = u.shape[1]
            out = self.empty((k, r2), q.dtype)
            self._check(self.lib.sktt_gauge_factor(self.h, F64, L, k, r2, _ptr(q), k, 1, _ptr(u), r2, 1, _ptr(out), r2, 1))
        else:
            k, L = q.shape
            r = u.shape[0]
            out = self.empty((r, k), q.dtype)
            self._check(self.lib.sktt_gauge_factor(self.h, F64, L, r, k, _ptr(u), 1, L, _ptr(q), 1, L, _ptr(out), k, 1))
        return out

    def gauge_push(self, carry, core, left):
        """left: carry [k, r] @ core [r, L] -> [k, L];  else: core [L, r2] @ carry [r2, k] -> [L, k]."""
        if core.dtype != torch.float64 or carry.dtype != torch.float64 or max(carry.shape) > 64:
            return self.matmul(carry, core) if left else self.matmul(core, carry)
        if left:
            k, r = carry.shape
            L = core.shape[1]
            out = self.empty((k, L), core.dtype)
            self._check(self.lib.sktt_gauge_push(self.h, F64, L, k, r, _ptr(carry), r, 1, _ptr(core), 1, L, _ptr(out), 1, L))
        else:
            r2, k = carry.shape
            L = core.shape[0]
            out = self.empty((L, k), core.dtype)
            self._check(self.lib.sktt_gauge_push(self.h, F64, L, k, r2, _ptr(carry), 1, k, _ptr(core), r2, 1, _ptr(out), k, 1))
        return out

    def widen(self, x):
        """float64 -> complex128 copy on the device."""
        if x.dtype == torch.complex128:
            return x
        out = self.empty(x.shape, torch.complex128)
        self._check(self.lib.sktt_widen(self.h, x.numel(), _ptr(x.contiguous()), _ptr(out)))
        return out

    def axpby(self, alpha, x, beta, y, out=None):
        """out = alpha * x + beta * y (elementwise; complex scalars are allowed for complex128 operands)."""
        if out is None:
            out = torch.empty_like(x)
        alpha, beta = complex(alpha), complex(beta)
        al = (C.c_double * 2)(alpha.real, alpha.imag)
        be = (C.c_double * 2)(beta.real, beta.imag)
        self._check(self.lib.sktt_axpby(self.h, dtype_code(x), x.numel(), al, _ptr(x), be, _ptr(y), _ptr(out)))
        return out

    def dotc(self, x, y):
        """sum conj(x) * y (host scalar)."""
        out = (C.c_double * 2)(0.0, 0.0)
        self._check(self.lib.sktt_dotc(self.h, dtype_code(x), x.numel(), _ptr(x), _ptr(y), out))
        return complex(out[0], out[1]) if x.dtype == torch.complex128 else out[0]

    def nrm2(self, x):
        out = C.c_double(0.0)
        self._check(self.lib.sktt_nrm2(self.h, dtype_code(x), x.numel(), _ptr(x), C.byref(out)))
        return out.value


def common_dtype(*tensors):
    """Promote a group of device tensors to complex128 if any of them is complex."""
    if any(t.dtype == torch.complex128 for t in tensors):
        dev = get_device()
        return tuple(dev.widen(t) for t in tensors)
    return tensors


def get_device(index=None):
    """Per-process cache of Device objects keyed by CUDA device index."""
    if not torch.cuda.is_available():
        raise RuntimeError("scikit_tt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if index is None else int(index)
    dev = _contexts.get(idx)
    if dev is None:
        dev = Device(idx)
        _contexts[idx] = dev
        if len(_contexts) > 1:                                 # several GPUs in one process: guard every call
            for other in _contexts.values():
                if not isinstance(other.lib, _GuardedLib):
                    other.lib = _GuardedLib(other.lib, other.index)
    else:
        # follow torch's current stream
        with torch.cuda.device(idx):
            s = torch.cuda.current_stream()
        if s.cuda_stream != dev.stream.cuda_stream:
            dev.stream = s
            dev.lib.sktt_ctx_set_stream(dev.h, C.c_void_p(s.cuda_stream))
    return dev
