"""Golden vectors for data_driven.regression.arr from the LIVE reference.  Build container only:
    OPENBLAS_NUM_THREADS=1 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_arr_golden.py"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def basis(d):
    """Four functions per coordinate: 1, x_i, x_i^2, sin(x_i) (plain callables of a sample column, as regression.py:323)."""
    return [[(lambda t: 1.0), (lambda t, i=i: t[i]), (lambda t, i=i: t[i] ** 2), (lambda t, i=i: np.sin(t[i]))] for i in range(d)]


if __name__ == "__main__":
    import scikit_tt.data_driven.regression as reg
    from scikit_tt.tensor_train import TT
    rng = np.random.default_rng(21)
    d, m = 3, 80
    x = rng.uniform(-1, 1, (d, m))
    # two targets that are exactly representable (rank <= 2) plus noise-free structure
    y = np.stack([x[0] * x[1] ** 2 + np.sin(x[2]) + 0.5, x[0] * np.sin(x[1]) * x[2] ** 2 - x[1]])
    ranks = [1, 3, 3, 1]
    guess = TT([rng.standard_normal((ranks[i], 4, 1, ranks[i + 1])) for i in range(d)])
    out = {"x": x, "y": y}
    for i, c in enumerate(guess.cores):
        out[f"guess/{i}"] = c
    for reps in (1, 3):
        sol = reg.arr(x, y, basis(d), guess, repeats=reps, rcond=1e-10, progress=False)
        for k, t in enumerate(sol):
            for i, c in enumerate(t.cores):
                out[f"rep{reps}/row{k}/{i}"] = c
            print(reps, k, t.ranks)
    np.savez_compressed(os.path.join(HERE, "arr.npz"), **out)
