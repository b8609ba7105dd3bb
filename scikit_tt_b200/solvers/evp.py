"""ALS for (generalised) eigenvalue problems in TT format -- same call surface as
scikit_tt/solvers/evp.py of PGelss/scikit_tt (`als` :17-179, `power_method` :182-250), on the GPU.

Per micro-step: left / right interface stacks for the operator, the optional right-hand operator of
the pencil and the deflation vectors [evp.py:253-334] -> dense micro matrices plus the rank-one
deflation terms [:337-383] -> local eigen-solve [:417-443] -> SVD re-orthonormalisation of the
eigenvector block [:452-493].

Local eigen-solvers (all on the device):
  'eigh' : Hermitian micro matrix, the `number_ev` LARGEST eigenvalues (sigma ignored, as evp.py:434-439);
           cyclic Jacobi.  With operator_gevp the pencil is reduced through a Cholesky factor of B.
  'eig'  : general micro matrix, the `number_ev` eigenvalues closest to sigma, ascending |lambda - sigma|
           (evp.py:424-432); shift-invert Arnoldi on the LU factors of (M - sigma B).
  'eigs' : same eigenpairs, returned in DESCENDING |lambda - sigma| order as evp.py:417-422 does.

Conjugation: the left stacks conjugate the column-side copy of the solution core exactly as
evp.py:281-283 does; for real cores this is indistinguishable from the
Hermitian-consistent sle convention.  `hermitian_stacks=True` switches to the latter.
"""
import numpy as np
import torch

from .. import _device
from .. import tensor_train as tt
from ..tensor_train import TT
from . import _local, sle


def als(operator, initial_guess, previous=[], shift=0, operator_gevp=None, number_ev=1, repeats=1, conv_eps=1e-10,
        solver='eig', sigma=1, real=True, hermitian_stacks=False):
    dev = _device.get_device()
    d, k = operator.order, number_ev
    cplx = solver in ('eig', 'eigs') or _local.any_complex(operator, initial_guess, operator_gevp, *previous)
    dtype = torch.complex128 if cplx else torch.float64
    A = _local.Uploaded(dev, operator, dtype, vector=False)
    G = _local.Uploaded(dev, operator_gevp, dtype, vector=False) if operator_gevp is not None else None
    P = [_local.Uploaded(dev, t, dtype, vector=True) for t in previous]
    x = list(_local.Uploaded(dev, initial_guess, dtype, vector=True).cores)            # evp.py:90 (copy)
    one3, one2 = _local.ones(dev, (1, 1, 1), dtype), _local.ones(dev, (1, 1), dtype)
    Lop, Rop, Lg, Rg = [None] * d, [None] * d, [None] * d, [None] * d
    Lp = [[None] * d for _ in previous]
    Rp = [[None] * d for _ in previous]
    conj_mode = _device.CONJ_ROW if hermitian_stacks else _device.CONJ_COL

    def right(i):                                                                      # evp.py:295-334
        if i == d - 1:
            Rop[i], Rg[i] = one3, one3
            for j in range(len(P)):
                Rp[j][i] = one2
            return
        Rop[i] = dev.stack_right_op(Rop[i + 1], x[i + 1], A[i + 1])
        if G is not None:
            Rg[i] = dev.stack_right_op(Rg[i + 1], x[i + 1], G[i + 1])
        for j in range(len(P)):
            Rp[j][i] = dev.stack_right_rhs(Rp[j][i + 1], P[j][i + 1], x[i + 1])

    def left(i):                                                                       # evp.py:253-292
        if i == 0:
            Lop[i], Lg[i] = one3, one3
            for j in range(len(P)):
                Lp[j][i] = one2
            return
        Lop[i] = dev.stack_left_op(Lop[i - 1], x[i - 1], A[i - 1], conj_mode)
        if G is not None:
            Lg[i] = dev.stack_left_op(Lg[i - 1], x[i - 1], G[i - 1], conj_mode)
        for j in range(len(P)):
            Lp[j][i] = dev.stack_left_rhs(Lp[j][i - 1], P[j][i - 1], x[i - 1])

    def update(i, direction):                                                          # evp.py:337-495
        r, n, r2 = Lop[i].shape[0], A[i].shape[2], Rop[i].shape[0]
        if solver == 'eigh' and r * n * r2 > _local.EIGH_DENSE_LIMIT:
            lam, vec = _matrix_free_eigh(i, r, n, r2)
        else:
            M = dev.micro_matrix_als(Lop[i], A[i], Rop[i])
            B = dev.micro_matrix_als(Lg[i], G[i], Rg[i]) if G is not None else None
            for j in range(len(P)):
                t = dev.micro_rhs_als(Lp[j][i], P[j][i], Rp[j][i])
                dev.rank1_update(M, t, shift)                                          # evp.py:381
            lam, vec = _local_eig(dev, M, B, k, solver, sigma)                         # vec [N, k]
        if direction == 'forward':
            U, _, _, _ = dev.svd(vec.reshape(r * n, r2 * k))                           # evp.py:452-454
            rr = min(r2, U.shape[1])
            x[i] = U[:, :rr].contiguous().reshape(r, n, rr)
        elif i > 0:
            _, _, Vh, _ = dev.svd(vec.t().contiguous().reshape(k * r, n * r2))         # evp.py:469-477
            rr = min(r, Vh.shape[0])
            x[i] = Vh[:rr, :].contiguous().reshape(rr, n, r2)
        else:
            x[i] = vec.reshape(r, n, r2, k)                                            # evp.py:492-493
        lam = lam.detach().cpu().numpy()
        return np.real(lam) if real else lam                                           # evp.py:441-443

    def _matrix_free_eigh(i, r, n, r2):
        """evp.py:434-439 where the dense micro matrix cannot exist: the k largest eigenpairs of the Hermitian local
        operator by thick-restart Lanczos on the micro-matvec (plus the rank-one deflation terms of evp.py:376-381),
        started from the sweep's current core."""
        if G is not None:
            raise NotImplementedError("matrix-free 'eigh' micro solves do not cover operator_gevp")
        if dtype == torch.complex128 and not hermitian_stacks:
            raise NotImplementedError("matrix-free 'eigh' on complex cores needs hermitian_stacks=True (the reference's "
                                      "left-stack conjugation, evp.py:281-283, makes the projected operator non-Hermitian)")
        op = dev.local_op(Lop[i], A[i], Rop[i], prepare=True)
        ts = [dev.micro_rhs_als(Lp[j][i], P[j][i], Rp[j][i]).reshape(-1) for j in range(len(P))]

        def matvec(v):
            y = dev.local_matvec(op, v)
            for t in ts:                                                               # + shift * t (t^H v)
                c = dev.dotc(t, v.reshape(-1))
                yf = y.reshape(-1)
                dev.axpby(1.0, yf, shift * c, t, out=yf)
            return y
        x0 = x[i] if x[i].dim() == 3 and tuple(x[i].shape) == (r, n, r2) else None
        theta, vec = _local.eigh_matrix_free(dev, matvec, (r, n, r2), dtype, k, v0=x0)
        return theta, vec

    for i in range(d - 1, -1, -1):                                                     # evp.py:103-104
        right(i)
    it = 1
    pre = np.array([np.inf] * k)[None, :]                                              # evp.py:110
    conv = False
    lam_opt, x_opt, lam = np.inf, None, None
    while it <= repeats and not conv:                                                  # evp.py:118
        for i in range(d):
            left(i)
            if i < d - 1:
                lam = update(i, 'forward')
        for i in range(d - 1, -1, -1):
            right(i)
            lam = update(i, 'backward')
        it += 1
        if k == 1 and np.abs(lam[0] - sigma) < np.abs(lam_opt - sigma):               # evp.py:151-155
            lam_opt = lam[0].copy()
            x_opt = [x[0][:, :, :, 0].contiguous()] + list(x[1:])
        last = pre[-min(3, pre.shape[0]):, :]                                          # evp.py:158-165
        if np.amax(np.abs(last - lam)) < conv_eps:
            conv = True
        pre = np.vstack((pre, lam))
    if k == 1:
        return lam_opt, TT(_local.download_vector_cores(x_opt)), it - 1
    tensors = [TT(_local.download_vector_cores([x[0][:, :, :, j].contiguous()] + list(x[1:]))) for j in range(k)]
    return lam, tensors, it - 1


# Outcome of the batched local eigen-solves of the most recent als_batch call: how many there were, how many ended without
# reaching the Arnoldi tolerance, and the worst relative residual estimate among those that were accepted anyway.
batch_stats = {"eig_solves": 0, "unconverged": 0, "worst_relres": 0.0, "redone_systems": 0}


def als_batch(operators, initial_guesses, previous=[], shift=0, operator_gevp=None, number_ev=1, repeats=1, conv_eps=1e-10,
              solver='eig', sigma=1, real=True, hermitian_stacks=False, on_unconverged=None):
    """evp.als for a batch of independent operators with identical shapes (BASELINE config 5: the CO-pressure sweep the
    reference loops over, examples/co_oxidation.py:100-104) -- the loop over the systems is a grid dimension of every
    kernel: per micro step 3 + 2 launches for the stack update and the dense micro matrices of ALL systems, ONE launch
    for all local eigen-solves (csrc/batch.cu: LU, shift-invert Arnoldi, Hessenberg QR, Ritz vectors inside one CTA per
    system) and ONE for all SVD re-orthonormalisations.  `initial_guesses`: one TT shared by all systems, or a list.

    Returns a list of (eigenvalue(s), eigentensor(s), iterations), one entry per operator, each exactly what `als` returns
    for that operator.  Falls back to a loop over `als` for anything the batched kernels do not cover (deflation, pencils,
    'eigh', micro matrices above 1024 unknowns, ragged shapes).

    on_unconverged: what happens to a system one of whose local eigen-solves ended without reaching the Arnoldi tolerance
    inside the batched kernel (restarted Arnoldi, Krylov dimension 20, relative residual 1e-12, at most 5 restarts, leaving as
    soon as a restart cycle gains less than a factor of ten): 'redo' runs that system again through
    the host-driven path, whose last resort is the exact full-space solve (what `lin.eig` delivers) -- the default for
    solver='eig'; 'accept' keeps the best Ritz pair and reports it (batch_stats, one RuntimeWarning per call) -- the default
    for the iterative solver='eigs', where the reference's ARPACK is restarted Arnoldi too.  Zero pivots and failures of the
    projected eigen-solve are always redone."""
    ops = list(operators)
    B = len(ops)
    guesses = list(initial_guesses) if isinstance(initial_guesses, (list, tuple)) else [initial_guesses] * B
    kw = dict(previous=previous, shift=shift, operator_gevp=operator_gevp, number_ev=number_ev, repeats=repeats,
              conv_eps=conv_eps, solver=solver, sigma=sigma, real=real, hermitian_stacks=hermitian_stacks)
    if on_unconverged is None:
        on_unconverged = 'redo' if solver == 'eig' else 'accept'
    if on_unconverged not in ('redo', 'accept'):
        raise ValueError("on_unconverged must be 'redo' or 'accept'")
    batch_stats.update(eig_solves=0, unconverged=0, worst_relres=0.0, redone_systems=0)
    loop = lambda idx: [als(ops[j], guesses[j], **kw) for j in idx]
    if B == 0:
        return []
    d, k = ops[0].order, number_ev
    same = all(o.ranks == ops[0].ranks and o.row_dims == ops[0].row_dims and o.col_dims == ops[0].col_dims for o in ops) \
        and all(g.ranks == guesses[0].ranks and g.row_dims == guesses[0].row_dims for g in guesses)
    dev = _device.get_device()
    xr, nd = guesses[0].ranks, ops[0].row_dims
    n_max = max(xr[i] * nd[i] * xr[i + 1] for i in range(d))
    if (not same or previous or operator_gevp is not None or solver not in ('eig', 'eigs') or k > 8
            or n_max > dev.BATCH_EIG_MAX_N or n_max < k or d < 2):
        return loop(range(B))
    dtype = torch.complex128
    A = [dev.upload_many([np.stack([o.cores[i] for o in ops])], dtype)[0] for i in range(d)]            # [B, R, m, n, R2]
    x = [dev.upload_many([np.stack([g.cores[i][:, :, 0, :] for g in guesses])], dtype)[0] for i in range(d)]  # [B, r, n, r2]
    one3 = torch.ones((B, 1, 1, 1), dtype=dtype, device=dev.device)
    Lop, Rop = [None] * d, [None] * d
    conj_mode = _device.CONJ_ROW if hermitian_stacks else _device.CONJ_COL
    big = 1 << 40
    checks = []                                                          # device status of every batched eigen-solve

    def right(i):                                                        # evp.py:295-330
        Rop[i] = one3 if i == d - 1 else dev.batch_stack_right_op(Rop[i + 1], x[i + 1], A[i + 1])

    def left(i):                                                         # evp.py:253-288
        Lop[i] = one3 if i == 0 else dev.batch_stack_left_op(Lop[i - 1], x[i - 1], A[i - 1], conj_mode)

    def update(i, direction):                                            # evp.py:337-495
        r, n, r2 = Lop[i].shape[1], A[i].shape[3], Rop[i].shape[1]
        M = dev.batch_micro_matrix_als(Lop[i], A[i], Rop[i])
        lam, vec, status = dev.batch_eig_shift_invert(M, sigma, k)       # lam [B, k], vec [B, N, k]
        checks.append(status)
        if solver == 'eigs' and k > 1:                                   # evp.py:421: descending |lambda - sigma|
            lam, vec = torch.flip(lam, dims=[1]), torch.flip(vec, dims=[2]).contiguous()
        if direction == 'forward':                                       # evp.py:452-464
            P, Q = r * n, r2 * k
            rr = min(r2, P, Q)
            out = dev.empty((B, r, n, rr), dtype)
            dev.batch_svd_left(vec, P, Q, rr, (big, 0, Q), (big, 0, 1), 0, out, rr, 1, 0)
            x[i] = out
        elif i > 0:                                                      # evp.py:469-487
            P, Q = n * r2, k * r
            rr = min(r, P, Q)
            out = dev.empty((B, rr, n, r2), dtype)
            dev.batch_svd_left(vec, P, Q, rr, (big, 0, k), (r, 1, n * r2 * k), 1, out, 1, n * r2, 1)
            x[i] = out
        else:
            x[i] = vec.reshape(B, r, n, r2, k)                           # evp.py:492-493
        return lam

    for i in range(d - 1, -1, -1):                                       # evp.py:103-104
        right(i)
    it = 1
    pre = np.full((1, B, k), np.inf)                                     # evp.py:110, per system
    done = np.zeros(B, dtype=bool)
    iters = np.zeros(B, dtype=int)
    lam_opt = np.full(B, np.inf, dtype=complex)
    x_opt = [None] * B
    lam_last = np.zeros((B, k), dtype=complex)
    x_last = [None] * B
    bad = np.zeros(B, dtype=bool)
    while it <= repeats and not done.all():                              # evp.py:118, every unconverged system
        for i in range(d):
            left(i)
            if i < d - 1:
                update(i, 'forward')
        for i in range(d - 1, -1, -1):
            right(i)
            lam_dev = update(i, 'backward')
        # one synchronisation per sweep: eigenvalues of the last micro step, outcome flags of all of them
        st = torch.stack(checks).cpu().numpy()               # [micro steps, B, 3]: converged pairs, info, relative residual
        checks.clear()
        unconv = st[:, :, 0] < k
        batch_stats["eig_solves"] += int(st.shape[0] * B)
        batch_stats["unconverged"] += int(unconv.sum())
        if unconv.any():
            batch_stats["worst_relres"] = max(batch_stats["worst_relres"], float(st[:, :, 2][unconv].max()))
        bad |= (st[:, :, 1] != 0).any(axis=0)
        if on_unconverged == 'redo':
            bad |= unconv.any(axis=0)
        lam_h = lam_dev.cpu().numpy()
        lam_h = np.real(lam_h).astype(complex) if real else lam_h       # evp.py:441-443
        cores_h = None
        for j in range(B):
            if done[j]:
                continue
            iters[j] = it
            lam_last[j] = lam_h[j]
            if k == 1 and np.abs(lam_h[j, 0] - sigma) < np.abs(lam_opt[j] - sigma):       # evp.py:151-155
                if cores_h is None:
                    cores_h = [c.cpu().numpy() for c in x]
                lam_opt[j] = lam_h[j, 0]
                x_opt[j] = [cores_h[0][j][:, :, :, 0].copy()] + [c[j].copy() for c in cores_h[1:]]
            last = pre[-min(3, pre.shape[0]):, j, :]                     # evp.py:158-165
            if np.amax(np.abs(last - lam_h[j])) < conv_eps:
                done[j] = True
                if k > 1:
                    if cores_h is None:
                        cores_h = [c.cpu().numpy() for c in x]
                    x_last[j] = [c[j].copy() for c in cores_h]
        pre = np.vstack((pre, lam_h[None, :, :]))
        it += 1
    if k > 1:
        cores_h = [c.cpu().numpy() for c in x]
        for j in range(B):
            if x_last[j] is None:
                x_last[j] = [c[j].copy() for c in cores_h]
    results = []
    as4 = lambda c: c.reshape(c.shape[0], c.shape[1], 1, c.shape[2])
    for j in range(B):
        if bad[j]:
            results.append(None)
        elif k == 1:
            lam = lam_opt[j].real if real else lam_opt[j]
            results.append((lam, TT([as4(c) for c in x_opt[j]]), int(iters[j])))
        else:
            lam = np.real(lam_last[j]) if real else lam_last[j]
            tensors = [TT([as4(np.ascontiguousarray(x_last[j][0][:, :, :, s]))] + [as4(c) for c in x_last[j][1:]])
                       for s in range(k)]
            results.append((lam, tensors, int(iters[j])))
    redo = [j for j in range(B) if results[j] is None]                   # the host-driven path decides (retries, exact
    batch_stats["redone_systems"] = len(redo)                            # fallback, raises) for those systems
    for j, res in zip(redo, loop(redo)):
        results[j] = res
    if on_unconverged == 'accept' and batch_stats["unconverged"]:
        import warnings
        warnings.warn(f"evp.als_batch: {batch_stats['unconverged']} of {batch_stats['eig_solves']} local eigen-solves ended "
                      f"above the Arnoldi tolerance (worst relative residual estimate {batch_stats['worst_relres']:.1e}); their "
                      f"best Ritz pairs were used -- pass on_unconverged='redo' for the exact host-driven path", RuntimeWarning)
    return results


def _local_eig(dev, M, B, k, solver, sigma):
    """The micro eigen-solve of evp.py:417-439 on the device; returns (lam [k], vec [N, k])."""
    N = M.shape[0]
    if solver == 'eigh':
        if B is None:
            W, V = dev.eigh(M)
        else:
            # pencil (M, B), B Hermitian positive definite: B = L L^H, C = L^-1 M L^-H, v = L^-H y
            if dev.chol_factor(B) != 0:
                raise np.linalg.LinAlgError("the micro matrix of operator_gevp is not positive definite")
            Y = dev.chol_trsm(B, M, backward=False)
            Z = dev.chol_trsm(B, Y.mH.contiguous(), backward=False)
            W, Yv = dev.eigh(Z)
            V = dev.chol_trsm(B, Yv, backward=True)
        kk = min(k, N)
        lam = torch.flip(W[N - kk:], dims=[0])                                        # largest first, evp.py:438-439
        vec = torch.flip(V[:, N - kk:], dims=[1]).contiguous()
        return lam, vec
    if solver in ('eig', 'eigs'):
        if M.dtype != torch.complex128:
            M = dev.widen(M)
            B = dev.widen(B) if B is not None else None
        lam, vec = dev.eig_shift_invert(M, sigma, k, B=B)
        order = torch.argsort(torch.abs(lam - sigma), descending=(solver == 'eigs'), stable=True)
        return lam[order], vec[:, order].contiguous()
    raise ValueError("solver must be 'eig', 'eigs' or 'eigh'")


def power_method(operator, initial_guess, operator_gevp=None, repeats=10, sigma=0.999):
    """Inverse power iteration on top of sle.als (evp.py:182-250)."""
    if operator_gevp is None:
        shifted = operator - sigma * tt.eye(operator.row_dims)
    else:
        shifted = operator - sigma * operator_gevp
    eigenvalue, eigentensor = 0, initial_guess
    for _ in range(repeats):
        rhs = eigentensor if operator_gevp is None else operator_gevp.dot(eigentensor)
        eigentensor = sle.als(shifted, eigentensor, rhs)
        eigentensor = eigentensor * (1 / eigentensor.norm())
        eigenvalue = eigentensor.transpose().dot(operator).dot(eigentensor)
        if operator_gevp is not None:
            eigenvalue *= 1 / (eigentensor.transpose().dot(operator_gevp).dot(eigentensor))
    return eigenvalue, eigentensor
