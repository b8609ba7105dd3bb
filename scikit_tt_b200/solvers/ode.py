"""Time stepping on top of the GPU ALS/MALS solvers -- `implicit_euler`, `trapezoidal_rule` and `adaptive_step_size`
with the call surfaces of scikit_tt/solvers/ode.py:249-330, :366-450, :487-636 of PGelss/scikit_tt."""
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
