"""Alternating ridge regression on transformed data tensors -- `arr` with the call surface of
scikit_tt/data_driven/regression.py:15-142 of PGelss/scikit_tt, on the GPU.

The sweep is the ALS one with sample-indexed interface stacks [regression.py:297-357], a micro matrix of shape
(r n r') x m [:360-393] and a least-squares micro solve with a singular-value cut [:419], followed by the QR / RQ gauge
step of sle.als [:421-440].  The basis functions are arbitrary host callables (`basis_list[i][k](x_column)`, as there): they
are evaluated once per mode on the host; everything after that -- stacks, micro matrices, the SVD-based least-squares
solves, QR / RQ -- runs through the C-ABI on device-resident data.
"""
import time as _time

import numpy as np
import torch

from .. import _device
from .. import utils as utl
from ..tensor_train import TT


def _basis_values(basis_list, x_data):
    """Phi_i[k, j] = basis_list[i][k](x_data[:, j])  (regression.py:323, :354, :388 evaluate this again in every micro step)."""
    m = x_data.shape[1]
    return [np.array([[f(x_data[:, j]) for j in range(m)] for f in mode], dtype=float) for mode in basis_list]


def arr(x_data, y_data, basis_list, initial_guess, repeats=1, rcond=10 ** -2, string='ARR', progress=True):
    """Alternating ridge regression (regression.py:15-142): one TT of coefficients per row of y_data.
    `initial_guess` is one TT (copied for every row) or a list of TTs; returns the list of solution TTs."""
    start_time = utl.progress(string, 0, show=progress)
    dev = _device.get_device()
    order = len(basis_list)
    rows = y_data.shape[0]
    total = rows * repeats * (2 * order - 1)
    counter = 0
    guesses = initial_guess if isinstance(initial_guess, list) else [initial_guess.copy() for _ in range(rows)]
    Phi = [dev.to_device(p) for p in _basis_values(basis_list, x_data)]                 # [n_i, m] each
    m = x_data.shape[1]
    one = torch.ones((1, m), dtype=torch.float64, device=dev.device)                    # regression.py:318 / :349 broadcast
    solution = []
    for k in range(rows):
        rhs = dev.to_device(np.asarray(y_data[k, :], dtype=float))
        x = dev.upload_many([c[:, :, 0, :] for c in guesses[k].cores], torch.float64)
        x = [c.contiguous() for c in x]
        left, right = [None] * order, [None] * order

        def build_right(i):
            right[i] = one if i == order - 1 else dev.arr_stack_right(right[i + 1], Phi[i + 1], x[i + 1])

        def build_left(i):
            left[i] = one if i == 0 else dev.arr_stack_left(left[i - 1], Phi[i - 1], x[i - 1])

        def update(i, direction):
            r, n, r2 = x[i].shape
            M = dev.arr_micro_matrix(left[i], Phi[i], right[i])                          # [(r n r2), m]
            core = dev.lstsq_gelss(M.t().contiguous(), rhs, rcond)                       # regression.py:419
            if direction == 'forward':
                q = dev.qr(core.reshape(r * n, r2).contiguous())                         # regression.py:421-427
                x[i] = q.reshape(r, n, q.shape[1]).contiguous()
            elif i > 0:
                q = dev.rq(core.reshape(r, n * r2).contiguous())                         # regression.py:430-436
                x[i] = q.reshape(q.shape[0], n, r2).contiguous()
            else:
                x[i] = core.reshape(r, n, r2).contiguous()

        for i in range(order - 1, -1, -1):
            build_right(i)
        for _ in range(repeats):
            for i in range(order):
                build_left(i)
                if i < order - 1:
                    update(i, 'forward')
                    counter += 1
            utl.progress(string, 100 * counter / total, cpu_time=_time.time() - start_time, show=progress)
            for i in range(order - 1, -1, -1):
                build_right(i)
                update(i, 'backward')
                counter += 1
            utl.progress(string, 100 * counter / total, cpu_time=_time.time() - start_time, show=progress)
        host = dev.download_many(x)
        solution.append(TT([h.reshape(h.shape[0], h.shape[1], 1, h.shape[2]) for h in host]))
    return solution
