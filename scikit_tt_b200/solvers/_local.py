"""Device-resident state shared by the ALS-type sweeps (sle.als / sle.mals / evp.als).

All cores, interface stacks and micro systems live in HBM for the whole solver call; numpy cores are
uploaded once on entry and downloaded once on exit.  Every arithmetic step is a C-ABI call
(include/sktt_b200.h) issued through scikit_tt_b200._device.Device.
"""
import numpy as np
import torch

from .. import _device

# The reference always assembles the dense micro matrix (sle.py:339-345) and LU-factorises it.  We do
# exactly that while the matrix is small enough to be worth it; beyond DENSE_LIMIT unknowns the same
# micro system is solved matrix-free (CG when the local operator is Hermitian, GMRES otherwise) to
# KRYLOV_TOL relative residual.  The reference cannot run in that regime at all.
DENSE_LIMIT = 8192
KRYLOV_TOL = 1e-13
KRYLOV_MAX_ITERS = 20000
GMRES_RESTART = 60


def any_complex(*trains):
    return any(np.iscomplexobj(c) for t in trains if t is not None for c in t.cores)


class Uploaded:
    """Device copies of the cores of a TT: operators as [R, m, n, R2], vectors as [r, n, r2]."""

    def __init__(self, dev, tt, dtype, vector):
        self.cores = []
        for c in tt.cores:
            t = dev.to_device(c[:, :, 0, :] if vector else c, dtype)
            self.cores.append(t)

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


def solve_micro(dev, solver, dense_builder, op, f, guess):
    """Solve the micro system M u = f.  `dense_builder()` returns the dense matrix (destroyed by the LU),
    `op` is the matrix-free description of the same M, `f` / `guess` have the unknown's tensor shape.
    Returns the flat solution [N]."""
    N = f.numel()
    mode = solver
    if solver in ('solve', 'lu'):
        mode = 'dense' if N <= DENSE_LIMIT else 'krylov'
    if mode == 'dense':
        M = dense_builder()
        return dev.solve(M, f)
    u = guess.reshape(-1).clone() if (guess is not None and guess.numel() == N) else torch.zeros(N, dtype=f.dtype, device=dev.device)
    method = mode
    if mode == 'krylov':
        method = 'cg' if is_hermitian_local(dev, op, tuple(f.shape), f.dtype) else 'gmres'
    if method == 'cg':
        st, iters, relres = dev.krylov_solve(op, f, u, method='cg', tol=KRYLOV_TOL, max_iters=KRYLOV_MAX_ITERS)
        if st == 0:
            return u
        if mode == 'cg':
            raise np.linalg.LinAlgError(f"cg micro solve failed (status {st}, relres {relres:.2e} after {iters} iterations)")
        u.zero_()                                             # not positive definite after all: fall through to GMRES
    st, iters, relres = dev.krylov_solve(op, f, u, method='gmres', tol=KRYLOV_TOL, max_iters=KRYLOV_MAX_ITERS,
                                         restart=GMRES_RESTART)
    if st != 0:
        raise np.linalg.LinAlgError(f"gmres micro solve failed (status {st}, relres {relres:.2e} after {iters} iterations)")
    return u


def download_vector_cores(cores):
    """[r, n, r2] device tensors -> list of host [r, n, 1, r2] numpy cores."""
    out = []
    for c in cores:
        h = c.detach().cpu().numpy()
        out.append(np.ascontiguousarray(h.reshape(h.shape[0], h.shape[1], 1, h.shape[2])))
    return out
