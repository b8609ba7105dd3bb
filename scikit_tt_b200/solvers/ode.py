"""Time stepping on top of the GPU ALS/MALS solvers -- `implicit_euler` with the call surface of
scikit_tt/solvers/ode.py:249-330 of PGelss/scikit_tt."""
import time as _time

import numpy as np

from .. import tensor_train as tt
from .. import utils as utl
from . import sle


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
