"""Multi-GPU forms of the ALS hot path, one process per GPU under torch.distributed.

Two partitions, and only these two, because the sweep itself is sequential in the core index and in time:

1. Independent systems (BASELINE config 5: a sweep over co_oxidation CO pressures; several right-hand sides; separate
   trajectories).  `map_sharded` gives every rank a contiguous block of the work list, runs the unchanged single-GPU
   solver on it (no data-path collective) and gathers the results once at the end.  `evp_als_batch` / `sle_als_batch`
   are the two front ends; the reference has no counterpart (it would loop: examples/co_oxidation.py:100-104).

2. One system at very large rank (config 4): the micro-matvec
       y[c,m,c2] = sum L[a,b,c] v[a,n,a2] A[b,m,n,b2] Rt[a2,b2,c2]
   is sharded over the OUTPUT solution-rank index `c` (sharding the contracted index `a` would only shrink the first of
   the three contractions: the partial T1 is full-size).  Rank g computes y[c_g] with the strided GEMM engine (no slice
   copies: the shard is an offset and an extent of the left stack), then one collective over NVLink (NCCL all_gather,
   or all_reduce for ragged blocks) assembles y everywhere.  `sharded_micro_matvec` is that step; `cg_sharded` iterates it.

Both work with any initialised process group; the CPU tests run them over gloo with world size 2 (tests/test_multi.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from ..tensor_train import TT


# ------------------------------------------------------------------------------------------------ partition 1
def shard_bounds(n_items, world, rank):
    """Contiguous block partition: the first n_items % world ranks get one extra item.  Returns (lo, hi)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def map_sharded(fn, items, group=None, gather=True):
    """results[i] = fn(items[i]) with the items block-partitioned over the ranks of `group`.

    No communication while the items are processed; one all_gather_object of the (picklable, host-side) results at the
    end.  With gather=False every rank only returns its own block [(index, result), ...]."""
    world, rank = _world(group)
    lo, hi = shard_bounds(len(items), world, rank)
    mine = [(i, fn(items[i])) for i in range(lo, hi)]
    if not gather:
        return mine
    if world == 1:
        return [r for _, r in mine]
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    out = [None] * len(items)
    for part in parts:
        for i, r in part:
            out[i] = r
    return out


def _pack(t):
    return [np.ascontiguousarray(c) for c in t.cores]


def evp_als_batch(operators, initial_guesses, group=None, **kwargs):
    """evp.als on a list of independent operators: one contiguous block of the list per GPU (no data-path collective),
    and inside each GPU the block runs through the batched kernels of evp.als_batch (the systems are a grid dimension).
    `initial_guesses` is one TT (shared) or a list.  Returns a list of (eigenvalue(s), eigentensor(s), iterations) in
    input order on every rank; one host-side gather of the results at the end."""
    from . import evp
    guesses = initial_guesses if isinstance(initial_guesses, (list, tuple)) else [initial_guesses] * len(operators)
    world, rank = _world(group)
    lo, hi = shard_bounds(len(operators), world, rank)
    mine = []
    if hi > lo:
        for j, (lam, vec, it) in enumerate(evp.als_batch(operators[lo:hi], guesses[lo:hi], **kwargs)):
            vec = [_pack(v) for v in vec] if isinstance(vec, list) else _pack(vec)
            mine.append((lo + j, (lam, vec, it)))
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
        mine = [item for part in parts for item in part]
    res = [None] * len(operators)
    for i, r in mine:
        res[i] = r
    out = []
    for lam, vec, it in res:
        tens = [TT(v) for v in vec] if vec and isinstance(vec[0], list) else TT(vec)
        out.append((lam, tens, it))
    return out


def sle_als_batch(operators, initial_guesses, right_hand_sides, group=None, mals=False, **kwargs):
    """sle.als (or sle.mals) on independent systems; `operators` / `initial_guesses` may be single TTs shared by all
    right-hand sides.  Returns the list of solution TTs in input order on every rank."""
    from . import sle
    n = len(right_hand_sides)
    ops = operators if isinstance(operators, (list, tuple)) else [operators] * n
    guesses = initial_guesses if isinstance(initial_guesses, (list, tuple)) else [initial_guesses] * n
    solver = sle.mals if mals else sle.als
    res = map_sharded(lambda i: _pack(solver(ops[i], guesses[i], right_hand_sides[i], **kwargs)), list(range(n)), group=group)
    return [TT(c) for c in res]


# ------------------------------------------------------------------------------------------------ partition 2
def matvec_rows_device(dev, L, A, Rt, v, lo, hi):
    """y[lo:hi, :, :] of the micro-matvec through three strided GEMMs of the contraction engine (device tensors).

    Only the rows c in [lo, hi) of the left stack enter, so all three contractions shrink with the shard:
      T1[b,cs,n,a2]  = sum_a  L[a,b,lo+cs] v[a,n,a2]
      T2[cs,m,a2,b2] = sum_bn A[b,m,n,b2] T1[b,cs,n,a2]
      y[cs,m,c2]     = sum    T2[cs,m,a2,b2] Rt[a2,b2,c2]
    The shard of L is an offset into the same buffer (two-level index map), never a copy."""
    big = 1 << 40
    r, R, _ = L.shape
    _, m, n, R2 = A.shape
    r2 = Rt.shape[0]
    rc = hi - lo
    T1 = dev.empty((R * rc, n * r2), v.dtype)
    dev.gemm2(R * rc, n * r2, r, L.reshape(-1)[lo:], (rc, r, 1), (big, 0, R * r), v, (big, 0, n * r2), (big, 0, 1),
              T1, (big, 0, n * r2), (big, 0, 1))
    T2 = dev.empty((rc * m, r2 * R2), v.dtype)
    dev.gemm2(rc * r2, m * R2, R * n, T1, (r2, n * r2, 1), (n, rc * n * r2, r2), A, (n, m * n * R2, R2), (R2, n * R2, 1),
              T2, (r2, m * r2 * R2, R2), (R2, r2 * R2, 1))
    y = dev.empty((rc, m, r2), v.dtype)
    dev.gemm2(rc * m, r2, r2 * R2, T2, (big, 0, r2 * R2), (big, 0, 1), Rt, (big, 0, r2), (big, 0, 1),
              y, (big, 0, r2), (big, 0, 1))
    return y


def sharded_micro_matvec(L, A, Rt, v, group=None, rows=None, dev=None):
    """y = M v with the OUTPUT solution-rank index c sharded over `group`: every rank holds the (small) operands and the
    full vector v, computes the rows [lo, hi) of y, and one collective over NVLink assembles y on every rank
    (all_gather when the blocks are equal, all_reduce of the zero-padded blocks otherwise).

    `rows(L, A, Rt, v, lo, hi)` computes one block; it defaults to the CUDA contraction engine on `dev`."""
    world, rank = _world(group)
    r = L.shape[0]
    lo, hi = shard_bounds(r, world, rank)
    if rows is None:
        if dev is None:
            from .. import _device
            dev = _device.get_device()
        rows = lambda *a: matvec_rows_device(dev, *a)
    mine = rows(L, A, Rt, v, lo, hi) if hi > lo else None
    if world == 1:
        return mine
    as_tensor = (lambda t: t) if torch.is_tensor(v) else torch.from_numpy
    m, r2 = A.shape[1], Rt.shape[0]
    ref = as_tensor(v)
    y = torch.zeros((r, m, r2), dtype=ref.dtype, device=ref.device)
    if r % world == 0:
        dist.all_gather_into_tensor(y, as_tensor(mine).contiguous(), group=group)
    else:
        if mine is not None:
            y[lo:hi] = as_tensor(mine)
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y if torch.is_tensor(v) else y.numpy()


def cg_sharded(L, A, Rt, f, u0=None, tol=1e-13, max_iters=2000, group=None, dev=None):
    """CG on the micro system with the sharded matvec; every rank carries the full Krylov vectors and runs identical
    scalar recurrences (the assembled y is bit-identical on all ranks), so the only collective is the one inside the
    matvec.  Device tensors; returns (u, iterations, relative residual)."""
    if dev is None:
        from .. import _device
        dev = _device.get_device()
    mv = lambda x: sharded_micro_matvec(L, A, Rt, x.reshape(f.shape), group=group, dev=dev).reshape(-1)
    fv = f.reshape(-1)
    u = torch.zeros_like(fv) if u0 is None else u0.reshape(-1).clone()
    res = dev.axpby(-1.0, mv(u), 1.0, fv) if u0 is not None else fv.clone()
    p = res.clone()
    rr = dev.dotc(res, res).real if f.dtype == torch.complex128 else dev.dotc(res, res)
    f2 = dev.dotc(fv, fv).real if f.dtype == torch.complex128 else dev.dotc(fv, fv)
    it = 0
    while it < max_iters and rr > tol * tol * f2:
        Ap = mv(p)
        pAp = dev.dotc(p, Ap)
        pAp = pAp.real if isinstance(pAp, complex) else pAp
        if not pAp > 0.0:
            raise np.linalg.LinAlgError("cg_sharded: operator is not Hermitian positive definite")
        alpha = rr / pAp
        dev.axpby(alpha, p, 1.0, u, out=u)
        dev.axpby(-alpha, Ap, 1.0, res, out=res)
        rr_new = dev.dotc(res, res)
        rr_new = rr_new.real if isinstance(rr_new, complex) else rr_new
        dev.axpby(rr_new / rr, p, 1.0, res, out=p)
        rr = rr_new
        it += 1
    return u.reshape(f.shape), it, float(np.sqrt(rr / f2)) if f2 > 0 else 0.0


# ------------------------------------------------------------------------------------------------ partition 2, fused form
class _RawCuda:
    """Zero-copy torch view of device memory that was not allocated by torch (peer-exchange buffers come from cudaMalloc
    through the C-ABI so that they can be exported as CUDA IPC handles)."""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (int(nelem),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """The y buffers of the rank-sharded micro-matvec, mapped into every process of `group` (one process per GPU of one
    NVSwitch domain): two alternating buffers of `nelem` elements per rank plus one slot array for the flag barrier.  The
    last contraction of the matvec stores its rows of y into all of them from its epilogue (csrc/peer.cu); the barrier
    kernel orders those remote stores before the readers.  Two buffers: a rank may already be writing matvec t + 1 into a
    peer that is still reading y of matvec t; matvec t + 2 reuses the first buffer only after the barrier of t + 1, which a
    rank passes after it has finished reading y of matvec t (stream order)."""

    def __init__(self, dev, nelem, dtype, group=None):
        import ctypes as C
        self.dev, self.group, self.dtype = dev, group, dtype
        self.world, self.rank = _world(group)
        self.nelem = int(nelem)
        item = 16 if dtype == torch.complex128 else 8
        self._C = C
        self.local = []                                    # raw pointers: y0, y1, flags
        for nbytes in (self.nelem * item, self.nelem * item, 64 * 8):
            p = C.c_void_p()
            dev._check(dev.lib.sktt_peer_alloc(dev.h, int(nbytes), C.byref(p)))
            self.local.append(p.value)
        handles = []
        for p in self.local:
            buf = C.create_string_buffer(64)
            dev._check(dev.lib.sktt_peer_export(dev.h, C.c_void_p(p), buf))
            handles.append(buf.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self.opened = []
        self.ptrs = []                                     # ptrs[g] = [y0, y1, flags] of rank g as seen from this process
        for g in range(self.world):
            if g == self.rank:
                self.ptrs.append(list(self.local))
                continue
            mine = []
            for h in gathered[g]:
                p = C.c_void_p()
                dev._check(dev.lib.sktt_peer_open(dev.h, h, C.byref(p)))
                mine.append(p.value)
                self.opened.append(p.value)
            self.ptrs.append(mine)
        typestr = "<c16" if dtype == torch.complex128 else "<f8"
        self.y = [torch.as_tensor(_RawCuda(self.local[k], self.nelem, typestr), device=dev.device) for k in (0, 1)]
        self.timeout = torch.zeros(1, dtype=torch.int32, device=dev.device)
        self.epoch = 0
        self.turn = 0
        dist.barrier(group=group)                          # every mapping exists before the first remote store

    def barrier(self):
        C = self._C
        self.epoch += 1
        arr = (C.c_void_p * self.world)(*[self.ptrs[g][2] for g in range(self.world)])
        self.dev._check(self.dev.lib.sktt_peer_barrier(self.dev.h, self.world, self.rank, self.epoch, arr,
                                                       C.c_void_p(self.timeout.data_ptr())))

    def matvec(self, L, A, Rt, v):
        """y = M v, rows sharded over the ranks, all-gather fused into the last contraction.  Returns a view of the local
        exchange buffer shaped like v (valid until the matvec after the next one)."""
        C = self._C
        dev = self.dev
        r, R = L.shape[0], L.shape[1]
        _, m, n, R2 = A.shape
        r2 = Rt.shape[0]
        N = r * m * r2
        if N > self.nelem:
            raise ValueError("PeerExchange buffer too small for this micro system")
        lo, hi = shard_bounds(r, self.world, self.rank)
        k = self.turn
        self.turn ^= 1
        peers = [self.ptrs[g][k] for g in range(self.world) if g != self.rank]
        arr = (C.c_void_p * max(len(peers), 1))(*peers)
        w = dev.work(dev.lib.sktt_sharded_matvec_work(r, R, m, n, r2, R2, max(hi - lo, 1)), v.dtype, tag="shard")
        code = 1 if v.dtype == torch.complex128 else 0
        dev._check(dev.lib.sktt_sharded_matvec(dev.h, code, r, R, m, n, r2, R2, C.c_void_p(L.data_ptr()),
                                               C.c_void_p(A.data_ptr()), C.c_void_p(Rt.data_ptr()),
                                               C.c_void_p(v.data_ptr()), lo, hi, C.c_void_p(self.local[k]), len(peers), arr,
                                               C.c_void_p(w.data_ptr())))
        self.barrier()
        return self.y[k][:N].view(r, m, r2)

    def check(self):
        if int(self.timeout.item()) != 0:
            raise RuntimeError("peer barrier timed out: a rank of the group did not reach the exchange step")

    def close(self):
        dev = self.dev
        if getattr(self, "local", None) is None:
            return
        dev.sync()
        try:
            dist.barrier(group=self.group)                 # nobody unmaps while a peer may still store
        except Exception:
            pass
        self.y = None
        for p in self.opened:
            dev.lib.sktt_peer_close(dev.h, self._C.c_void_p(p))
        for p in self.local:
            dev.lib.sktt_peer_free(dev.h, self._C.c_void_p(p))
        self.local = None


# Exchange buffers are kept between calls: creating them costs three cudaMalloc, an all_gather_object and 3 (world - 1)
# cudaIpcOpenMemHandle per rank (tens of milliseconds at 8 GPUs, with large jitter: one sle.als(group=) sweep at rank 256
# measured between 4.9 and 12.8 half-sweeps/s on 4 GPUs depending on it).  One set per (process group, dtype, device), grown
# when a call needs more; released by close_peer_exchanges() or at interpreter exit.
_PX_CACHE = {}


def peer_exchange(dev, nelem, dtype, group=None):
    key = (id(group) if group is not None else None, str(dtype), dev.index)
    px = _PX_CACHE.get(key)
    if px is not None and (px.local is None or px.nelem < int(nelem)):
        px.close()
        px = None
    if px is None:
        px = _PX_CACHE[key] = PeerExchange(dev, nelem, dtype, group)
    return px


def close_peer_exchanges():
    for px in list(_PX_CACHE.values()):
        try:
            px.close()
        except Exception:
            pass
    _PX_CACHE.clear()


SHARDED_TOL = 1e-14
SHARDED_ACCEPT = 1e-12
SHARDED_MAX_ITERS = 5000
SHARDED_MAX_CYCLES = 5
sharded_stats = {"solves": 0, "matvecs": 0, "worst_relres": 0.0}


def solve_sharded(dev, matvec, f, guess=None, tol=SHARDED_TOL, max_iters=SHARDED_MAX_ITERS):
    """CG on the Hermitian positive definite micro system with a matvec that is sharded over the ranks (`matvec(v)` returns
    the assembled y on every rank, bit-identical everywhere, so every rank runs the same scalar recurrences and takes the
    same decisions: the exchange inside the matvec is the only communication).  Residual replacement as the one-GPU path:
    after every CG run the TRUE residual f - M u is recomputed and the correction equation solved again until it is below
    `tol` or stops improving.  Returns the flat solution; raises numpy.linalg.LinAlgError above SHARDED_ACCEPT."""
    shape = tuple(f.shape)
    fv = f.reshape(-1)
    real = f.dtype == torch.float64
    dot = (lambda a, b: dev.dotc(a, b)) if real else (lambda a, b: dev.dotc(a, b).real)
    fnorm = np.sqrt(dot(fv, fv))
    u = guess.reshape(-1).clone() if guess is not None and guess.numel() == fv.numel() else torch.zeros_like(fv)
    if fnorm == 0.0:
        return u.zero_()
    mv = lambda x: matvec(x.reshape(shape)).reshape(-1)
    prev, relres, used = np.inf, None, 0
    for cycle in range(SHARDED_MAX_CYCLES):
        res = dev.axpby(-1.0, mv(u), 1.0, fv)
        used += 1
        rr = dot(res, res)
        relres = np.sqrt(rr) / fnorm
        if cycle == 0 and not relres < 1.0:                      # a warm start worse than zero is dropped
            u.zero_()
            res, relres, rr = fv.clone(), 1.0, fnorm * fnorm
        if relres <= tol or relres > 0.5 * prev:
            break
        prev = relres
        target2 = (0.5 * tol * fnorm) ** 2
        e = torch.zeros_like(fv)
        p = res.clone()
        r_ = res
        for _ in range(max_iters):
            if rr <= target2:
                break
            Ap = mv(p)
            used += 1
            pAp = dot(p, Ap)
            if not pAp > 0.0:
                raise np.linalg.LinAlgError("sharded CG: the micro operator is not Hermitian positive definite")
            alpha = rr / pAp
            dev.axpby(alpha, p, 1.0, e, out=e)
            r_ = dev.axpby(-alpha, Ap, 1.0, r_)
            rr_new = dot(r_, r_)
            dev.axpby(rr_new / rr, p, 1.0, r_, out=p)
            rr = rr_new
        dev.axpby(1.0, e, 1.0, u, out=u)
    sharded_stats["solves"] += 1
    sharded_stats["matvecs"] += used
    sharded_stats["worst_relres"] = max(sharded_stats["worst_relres"], float(relres))
    if not relres <= SHARDED_ACCEPT:
        raise np.linalg.LinAlgError(f"sharded CG micro solve did not converge (true relative residual {relres:.2e})")
    return u
