"""Device-resident state shared by the ALS-type sweeps (sle.als / sle.mals / evp.als).

All cores, interface stacks and micro systems live in HBM for the whole solver call; numpy cores are
uploaded once on entry and downloaded once on exit.  Every arithmetic step is a C-ABI call
(include/sktt_b200.h) issued through scikit_tt_b200._device.Device.
"""
import os

import numpy as np
import torch

from .. import _device

# The reference always assembles the dense micro matrix (sle.py:339-345) and LU-factorises it.  For the reference's own
# solver names ('solve', 'lu') we do exactly that for as long as the matrix is comfortably held in HBM -- DENSE_LIMIT
# unknowns: 2 GiB in fp64, the largest size the reference itself factorises in bounded time -- so that
# ill-conditioned but nonsingular micro systems (I - hA of a stiff generator, near-singular shifts of power_method) are
# solved by the same backward-stable algorithm as there.  Beyond it the same micro system is solved matrix-free (CG when
# the local operator is Hermitian, GMRES otherwise) to KRYLOV_TOL true relative residual (residual replacement); the
# reference cannot run at those sizes at all.  A matrix-free solve that stagnates above KRYLOV_ACCEPT
# falls back to the dense LU while the matrix still fits (DENSE_FALLBACK_LIMIT), and raises only beyond that.
# 'dense' / 'cg' / 'gmres' as `solver` force one route.
DENSE_LIMIT = 16384
SMALL_DENSE_LIMIT = 2048    # up to here dense LU is also the FASTEST route; between SMALL_DENSE_LIMIT and DENSE_LIMIT a micro
                            # system that the persistent fused CG kernel covers (fp64, Hermitian, bench-type shapes: the edge
                            # cores of C3 have 4096 unknowns) goes matrix-free FIRST -- same answer to 1e-12 or the dense LU
                            # takes over -- everything else is factorised as the reference does
DENSE_FALLBACK_LIMIT = 32768
KRYLOV_TOL = 1e-14          # target TRUE relative residual of the matrix-free micro solves
KRYLOV_ACCEPT = 1e-12       # a solve that stagnates above this is redone by dense LU where that fits
KRYLOV_ACCEPT_LARGE = 1e-10  # ... and is accepted with a warning (recorded in `stats`) up to here where it cannot fit
KRYLOV_MAX_ITERS = 20000
KRYLOV_MAX_CYCLES = 5
GMRES_RESTART = 60
_TRACE = bool(int(os.environ.get("SKTT_TRACE", "0")))
PROFILE = None              # set to a dict to accumulate synchronised wall time per sweep phase (diagnostics only)


# Outcome of the matrix-free micro solves of the most recent solver call (sle.als / sle.mals): how many there were, their
# CG / GMRES iterations and the WORST accepted true relative residual || f - M u || / || f ||.  Diagnostics for callers and
# tests (at C3 size nothing else can detect a drifting micro solve: no reference result exists).
stats = {"krylov_solves": 0, "krylov_iterations": 0, "worst_relres": 0.0, "dense_fallbacks": 0}


def reset_stats():
    stats.update(krylov_solves=0, krylov_iterations=0, worst_relres=0.0, dense_fallbacks=0)


def _record(relres, iters, count=1):
    stats["krylov_solves"] += count
    stats["krylov_iterations"] += int(iters)
    if not relres <= stats["worst_relres"]:                 # NaN-propagating max
        stats["worst_relres"] = float(relres)


def _accept_limit(N):
    return KRYLOV_ACCEPT if N <= DENSE_FALLBACK_LIMIT else KRYLOV_ACCEPT_LARGE


class phase:
    """with phase(dev, 'qr'): ...  -- accumulates synchronised wall time into PROFILE when profiling is on."""

    def __init__(self, dev, name):
        self.dev, self.name = dev, name

    def __enter__(self):
        if PROFILE is not None:
            import time
            self.dev.sync()
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            import time
            self.dev.sync()
            PROFILE[self.name] = PROFILE.get(self.name, 0.0) + time.perf_counter() - self.t0
            PROFILE[self.name + "#"] = PROFILE.get(self.name + "#", 0) + 1


def any_complex(*trains):
    return any(np.iscomplexobj(c) for t in trains if t is not None for c in t.cores)


class Uploaded:
    """Device copies of the cores of a TT: operators as [R, m, n, R2], vectors as [r, n, r2]."""

    def __init__(self, dev, tt, dtype, vector):
        self.cores = dev.upload_many([c[:, :, 0, :] if vector else c for c in tt.cores], dtype)

    def __getitem__(self, i):
        return self.cores[i]

    def __len__(self):
        return len(self.cores)


def ones(dev, shape, dtype):
    return torch.ones(shape, dtype=dtype, device=dev.device)


def is_hermitian_local(dev, op, shape, dtype, tol=1e-11):
    """Randomised probe <u, M v> == <M u, v> of the matrix-free micro operator (two matvecs)."""
    g = torch.Generator(device=dev.device).manual_seed(1234)
    u = torch.randn(shape, dtype=dtype, device=dev.device, generator=g)
    v = torch.randn(shape, dtype=dtype, device=dev.device, generator=g)
    Mu, Mv = dev.local_matvec(op, u), dev.local_matvec(op, v)
    a = dev.dotc(u, Mv)
    b = dev.dotc(Mu, v)
    scale = dev.nrm2(u) * dev.nrm2(Mv)
    return abs(a - b) <= tol * max(scale, 1e-300)


def _krylov_refined(dev, op, f, u, method):
    """Krylov solve with residual replacement: after every Krylov cycle the TRUE residual f - M u is recomputed and
    the correction equation is solved again, until the true relative residual is below KRYLOV_TOL or stops
    improving (the floor eps * cond of any backward-stable solver, LU included).  Returns (true relres, iterations)."""
    fnorm = dev.nrm2(f)
    if fnorm == 0.0:
        u.zero_()
        return 0.0, 0
    shape = tuple(f.shape)
    total, relres, prev = 0, None, np.inf
    e = None
    for cycle in range(KRYLOV_MAX_CYCLES):
        res = dev.axpby(-1.0, dev.local_matvec(op, u.reshape(shape)).reshape(-1), 1.0, f.reshape(-1))
        relres = dev.nrm2(res) / fnorm
        if _TRACE:
            print(f"    [krylov] {method} cycle {cycle}: true relres {relres:.3e} after {total} iterations", flush=True)
        if cycle == 0 and not relres < 1.0:
            # a warm start that is worse than the zero vector (e.g. a guess core carrying the norm of a random tensor,
            # 1e57 at the bench shape) would make the Krylov solver spend its accuracy on cancelling the guess
            u.zero_()
            relres, res = 1.0, f.reshape(-1)
        if relres <= KRYLOV_TOL or relres > 0.5 * prev:
            break
        prev = relres
        if cycle == 0:
            target, rhs, x = KRYLOV_TOL, f.reshape(-1), u              # first cycle: the system itself, warm start
        else:
            if e is None:
                e = torch.zeros_like(u)
            e.zero_()
            target, rhs, x = min(0.5, KRYLOV_TOL / relres), res, e      # later cycles: correction equation M e = res
        st, iters, rr = dev.krylov_solve(op, rhs.reshape(shape), x, method=method, tol=0.5 * target,
                                         max_iters=KRYLOV_MAX_ITERS, restart=GMRES_RESTART)
        total += iters
        if st not in (0, 2):          # 2 = no convergence / CG breakdown: judged by the true residual of the next cycle
            raise _device.SkttError(st, "krylov_solve failed")
        if x is not u:
            dev.axpby(1.0, e, 1.0, u, out=u)
    return relres, total


class Deferred:
    """Outcome slots of micro solves that were queued without a host synchronisation (one row of 4 doubles per solve:
    iterations, true relative residual, cycles, status)."""

    def __init__(self, dev, capacity):
        self.slots = torch.zeros((max(int(capacity), 1), 4), dtype=torch.float64, device=dev.device)
        self.used = 0
        self.limit = KRYLOV_ACCEPT_LARGE

    def solve(self, dev, op, f, u):
        if self.used >= self.slots.shape[0]:
            return False
        if not dev.krylov_solve_refined_async(op, f, u, self.slots[self.used], tol=KRYLOV_TOL, max_cycles=KRYLOV_MAX_CYCLES):
            return False
        self.used += 1
        self.limit = min(self.limit, _accept_limit(f.numel()))
        return True

    def check(self):
        """One synchronisation for the whole sweep: True when every queued solve ended with status 0 and an acceptable
        true residual."""
        if self.used == 0:
            return True
        out = self.slots[: self.used].cpu().numpy()
        self.iterations = int(out[:, 0].sum())
        self.used = 0
        ok = bool(np.all(out[:, 3] == 0.0) and np.all(out[:, 1] <= self.limit))
        if ok:
            _record(float(out[:, 1].max()), self.iterations, count=out.shape[0])
        return ok


def solve_micro(dev, solver, dense_builder, op, f, guess, cache=None):
    """Solve the micro system M u = f.  `dense_builder()` returns the dense matrix (destroyed by the LU),
    `op` is the matrix-free description of the same M, `f` / `guess` have the unknown's tensor shape.
    Returns the flat solution [N]."""
    N = f.numel()
    mode = solver
    if solver in ('solve', 'lu'):
        mode = 'dense' if N <= DENSE_LIMIT else 'krylov'
        if SMALL_DENSE_LIMIT < N <= DENSE_LIMIT and f.dtype == torch.float64 and op is not None and op.sites == 1:
            dev.prepare_local_op(op)
            if dev.tiled_len(op) > 0 and (cache is None or cache.get('hermitian') is not False):
                mode = 'krylov'
    if mode == 'dense':
        M = dense_builder()
        if _TRACE:
            print(f"  [micro] N={N} dense LU", flush=True)
        return dev.solve(M, f)
    u = guess.reshape(-1).clone() if (guess is not None and guess.numel() == N) else torch.zeros(N, dtype=f.dtype, device=dev.device)
    dev.prepare_local_op(op)                                  # TMA tile images of (A, Rt) for the fused matvec, once per solve
    method = mode
    if mode == 'krylov':
        # Hermitian-ness of the local operators is inherited from the global operator (the stacks are built with the
        # conjugate core on the row side): probe the first Krylov micro system of a solver call, reuse the verdict
        herm = cache.get('hermitian') if cache is not None else None
        if herm is None:
            with phase(dev, 'probe'):
                herm = is_hermitian_local(dev, op, tuple(f.shape), f.dtype)
            if cache is not None:
                cache['hermitian'] = herm
        if not herm and solver in ('solve', 'lu') and N <= DENSE_LIMIT:
            return dev.solve(dense_builder(), f)              # not the fused CG's business: factorise as the reference does
        method = 'cg' if herm else 'gmres'
    if _TRACE:
        print(f"  [micro] N={N} {method}", flush=True)
    if method == 'cg' and f.dtype == torch.float64 and dev.tiled_len(op) > 0 and cache is not None \
            and cache.get('defer') is not None and not _TRACE:
        # deferred outcome: the solve is queued as one cooperative launch and the sweep goes on queueing behind it; the
        # caller looks at all outcomes once (Deferred.check) and redoes the sweep synchronously if one of them is bad
        if cache['defer'].solve(dev, op, f, u):
            return u
    if method == 'cg' and f.dtype == torch.float64 and dev.tiled_len(op) > 0:
        # prepared operator: the whole solve (warm start, CG, true-residual restarts) is one C call
        st, iters, relres, cycles = dev.krylov_solve_refined(op, f, u, tol=KRYLOV_TOL, max_iters=KRYLOV_MAX_ITERS,
                                                             max_cycles=KRYLOV_MAX_CYCLES)
        if st != 0:
            raise _device.SkttError(st, "krylov_solve_refined failed")
        if _TRACE:
            print(f"    [krylov] cg (one call): true relres {relres:.3e} after {iters} iterations, {cycles} cycles", flush=True)
    else:
        relres, iters = _krylov_refined(dev, op, f, u, method)
    limit = _accept_limit(N)
    if method == 'cg' and mode == 'krylov' and not relres <= limit:
        u.zero_()                                             # Hermitian but not definite: CG broke down, use GMRES
        if cache is not None:
            cache['hermitian'] = False
        relres, iters = _krylov_refined(dev, op, f, u, 'gmres')
    if not relres <= limit:
        if solver in ('solve', 'lu') and N <= DENSE_FALLBACK_LIMIT:
            # the reference's own algorithm while its matrix still fits: assemble + LU with partial pivoting
            stats["dense_fallbacks"] += 1
            return dev.solve(dense_builder(), f)
        raise np.linalg.LinAlgError(f"{method} micro solve did not converge (relative residual {relres:.2e} after {iters} iterations)")
    if relres > KRYLOV_ACCEPT:
        import warnings
        warnings.warn(f"matrix-free micro solve of {N} unknowns accepted at true relative residual {relres:.2e} "
                      f"(target {KRYLOV_TOL:.0e}; a dense factorisation does not fit at this size)", RuntimeWarning)
    _record(relres, iters)
    return u


def download_vector_cores(cores):
    """[r, n, r2] device tensors -> list of host [r, n, 1, r2] numpy cores."""
    if not cores:
        return []
    dev = _device.get_device()
    return [h.reshape(h.shape[0], h.shape[1], 1, h.shape[2]) for h in dev.download_many([c.detach() for c in cores])]


# ------------------------------------------------------------------------------------------------------------------
# Matrix-free Hermitian local eigen-solver: thick-restart Lanczos on the micro-matvec (north_star: "a Lanczos/LOBPCG for
# evp").  The reference's micro eigen-solve for solver='eigh' is scipy.linalg.eigh of the dense micro matrix, of which it
# keeps the `k` LARGEST eigenvalues (evp.py:434-439); beyond EIGH_DENSE_LIMIT unknowns that matrix cannot be formed (C3 shape:
# 512 GiB), so the same eigenpairs are computed from matvecs alone.  Every arithmetic step is a C-ABI call: the fused /
# strided-GEMM micro-matvec, the contraction engine for the (re)orthogonalisation against the Krylov basis (classical
# Gram-Schmidt, twice), cyclic Jacobi for the projected matrix.
EIGH_DENSE_LIMIT = 4096     # Jacobi eigh of the dense micro matrix up to here (sktt_eigh_jacobi), Lanczos above
LANCZOS_TOL = 1e-14         # residual ||M v - theta v|| <= tol * max |theta| for the k wanted pairs ...
LANCZOS_FLOOR = 1e-11       # ... or, once restarts stop improving it (rounding floor eps * cond), at most this
LANCZOS_MAX_RESTARTS = 200
lanczos_stats = {"solves": 0, "matvecs": 0, "restarts": 0, "worst_residual": 0.0}


def eigh_matrix_free(dev, matvec, shape, dtype, k, v0=None, tol=LANCZOS_TOL, ncv=None, max_restarts=LANCZOS_MAX_RESTARTS):
    """The k largest eigenpairs of the Hermitian operator `matvec` (device tensor of `shape` -> same shape).
    Returns (theta [k] float64 descending, vecs [N, k]) on the device; raises numpy.linalg.LinAlgError if the wanted pairs
    do not converge.  v0: starting vector (the sweep's current core is a far better start than a constant)."""
    N = int(np.prod(shape))
    k = int(k)
    m = int(ncv) if ncv is not None else max(2 * k + 24, 40)
    m = min(m, N)
    if k > m:
        raise ValueError("eigh_matrix_free: k exceeds the Krylov dimension")
    big = 1 << 40
    V = dev.empty((m + 1, N), dtype)
    T = torch.zeros((m, m), dtype=dtype, device=dev.device)
    w0 = v0.reshape(-1).clone() if v0 is not None and v0.numel() == N else torch.ones(N, dtype=dtype, device=dev.device)
    nrm = dev.nrm2(w0)
    if not nrm > 0.0:
        w0 = torch.ones(N, dtype=dtype, device=dev.device)
        nrm = dev.nrm2(w0)
    dev.axpby(1.0 / nrm, w0, 0.0, w0, out=V[0])
    q = 0                                        # number of locked-in Ritz vectors at the head of V
    keep = min(m - 1, k + max(4, k))             # thick restart: the wanted pairs plus a few neighbours
    matvecs = 0
    prev_worst = np.inf
    for restart in range(max_restarts + 1):
        beta = 0.0
        for j in range(q, m):
            w = matvec(V[j].reshape(shape)).reshape(-1)
            matvecs += 1
            hsum = None
            for _ in range(2):                   # CGS2 against V[:j+1]; the sum of both passes is the column of T
                h = dev.empty((j + 1,), dtype)
                dev.gemm2(j + 1, 1, N, V, (big, 0, N), (big, 0, 1), w, (big, 0, 1), (big, 0, 0), h, (big, 0, 1), (big, 0, 0),
                          conjA=1)
                dev.gemm2(N, 1, j + 1, V, (big, 0, 1), (big, 0, N), h, (big, 0, 1), (big, 0, 0), w, (big, 0, 1), (big, 0, 0),
                          alpha=(-1.0, 0.0), beta=(1.0, 0.0))
                hsum = h if hsum is None else dev.axpby(1.0, h, 1.0, hsum)
            T[: j + 1, j] = hsum                 # (plumbing: a strided device copy)
            beta = dev.nrm2(w)
            if beta > 0.0:
                dev.axpby(1.0 / beta, w, 0.0, w, out=V[j + 1])
            else:
                V[j + 1].zero_()                 # invariant subspace: the Ritz pairs of this block are exact
            if j + 1 < m:
                T[j + 1, j] = beta
        # Hermitian projected matrix from its upper triangle, eigen-decomposition by cyclic Jacobi (ascending)
        Tu = torch.triu(T)
        Th = Tu + torch.triu(T, 1).mH
        W, Y = dev.eigh(Th.contiguous())
        Wh = W.cpu().numpy()
        last = Y[m - 1, :].cpu().numpy()
        scale = max(np.abs(Wh).max(), 1e-300)
        res = np.abs(beta * last[m - k:])        # residual norms of the k largest Ritz pairs
        worst = float(res.max() / scale)
        stalled = restart >= 2 and worst > 0.5 * prev_worst and worst <= LANCZOS_FLOOR
        prev_worst = worst
        done = bool(np.all(res <= tol * scale)) or m >= N or beta == 0.0 or stalled
        sel = slice(m - k, m) if done else slice(m - keep, m)
        Ys = Y[:, sel].contiguous()
        cnt = Ys.shape[1]
        X = dev.empty((cnt, N), dtype)           # Ritz vectors as rows: X = Ys^T V[:m]
        dev.gemm2(cnt, N, m, Ys, (big, 0, 1), (big, 0, cnt), V, (big, 0, N), (big, 0, 1), X, (big, 0, N), (big, 0, 1))
        if done:
            lanczos_stats["solves"] += 1
            lanczos_stats["matvecs"] += matvecs
            lanczos_stats["restarts"] += restart
            lanczos_stats["worst_residual"] = max(lanczos_stats["worst_residual"], float(res.max() / scale))
            theta = torch.flip(W[m - k:], dims=[0])
            vecs = torch.flip(X, dims=[0]).t().contiguous()       # [N, k], largest first (evp.py:438-439)
            return theta, vecs
        if restart == max_restarts:
            break
        # thick restart: V <- [Ritz vectors | v_{m+1}], T <- diag(theta) (the coupling row is rebuilt by the next CGS pass)
        vlast = V[m].clone()
        V[:cnt].copy_(X)
        V[cnt].copy_(vlast)
        T.zero_()
        T[:cnt, :cnt] = torch.diag(W[sel].to(dtype))
        q = cnt
    raise np.linalg.LinAlgError(f"thick-restart Lanczos: the {k} largest eigenpairs of the {N}-dimensional micro operator did "
                                f"not converge in {max_restarts} restarts (residual {float(res.max() / scale):.2e})")
