"""Oracle restatement of scikit_tt/solvers/ode.py:249-330 (implicit Euler driven by sle.als/mals)."""
import numpy as np

from . import sle, tt


def implicit_euler(op, x_init, guess, step_sizes, repeats=1, tt_solver='als', threshold=1e-12, max_rank=np.inf,
                   micro_solver='solve', normalize=1):
    """Returns the list of core lists [x_0, x_1, ...] (ode.py:301-330)."""
    sol = [x_init]
    cur = guess
    dims = [c.shape[1] for c in op]
    for i, h in enumerate(step_sizes):
        lhs = tt.sub(tt.eye(dims), tt.scale(op, h))                     # ode.py:313: I - h A
        if tt_solver == 'als':
            cur = sle.als(lhs, cur, sol[i], repeats=repeats, solver=micro_solver)
        else:
            cur = sle.mals(lhs, cur, sol[i], repeats=repeats, solver=micro_solver, threshold=threshold,
                           max_rank=max_rank)
        if normalize > 0:                                                # ode.py:320-321
            cur = tt.scale(cur, 1.0 / tt.norm(cur, p=normalize))
        sol.append(tt.copy_cores(cur))
    return sol


def trapezoidal_rule(op, x_init, guess, step_sizes, repeats=1, normalize=1):
    """ode.py:366-450 on core lists (ALS variant)."""
    sol = [x_init]
    cur = guess
    dims = [c.shape[1] for c in op]
    for i, h in enumerate(step_sizes):
        lhs = tt.sub(tt.eye(dims), tt.scale(op, 0.5 * h))
        rhs = tt.matmul(tt.add(tt.eye(dims), tt.scale(op, 0.5 * h)), sol[i])              # ode.py:431-437
        cur = sle.als(lhs, cur, rhs, repeats=repeats)
        if normalize > 0:
            cur = tt.scale(cur, 1.0 / tt.norm(cur, p=normalize))
        sol.append(tt.copy_cores(cur))
    return sol
