"""Synthetic inputs of the BASELINE.json configurations, shared by bench.py, the tests and the fixture
generators.  Plain numpy; every function returns lists of 4-D cores [r, m, n, r'] as the reference's TT class holds them.

Nothing here is on the measured path: these are the operator families the hot path is run on.
  C1  signaling cascade, d = 20 (the three distinct cores of models.signaling_cascade, from tests/golden/euler_cascade.npz)
  C2 / C5  CO oxidation on RuO2, d = 20 (SLIM cores of models.co_oxidation, affine in the CO adsorption rate:
      tests/golden/co_oxidation_slim.npz holds them at k = 0 and k = 1)
  C3  rank-3 Laplacian-type operator (symmetric positive definite), d = 32, n = 64
  C4  random operator d = 10, n = 16, R = 8: throughput variant (dense Gaussian cores) and the SPD parity / solve variant
      (rank-diagonal cores, blocks I + 0.1 sym(randn))
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden")


# ---------------------------------------------------------------------------------------------- C3
def laplace_cores(d, n, c=1e-3):
    S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = np.sqrt(c) * 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1))
    I, Z = np.eye(n), np.zeros((n, n))

    def core(rows):
        out = np.zeros((len(rows), n, n, len(rows[0])))
        for i, row in enumerate(rows):
            for j, blk in enumerate(row):
                out[i, :, :, j] = blk
        return out
    first = core([[S, D, I]])
    mid = core([[I, Z, Z], [D, Z, Z], [S, D, I]])
    last = core([[I], [D], [S]])
    return [first] + [mid.copy() for _ in range(d - 2)] + [last]


def capped_ranks(d, n, r):
    """[1, r, ..., r, 1] with no rank above what the unfoldings on either side support."""
    ranks = [1] + [r] * (d - 1) + [1]
    for i in range(1, d):
        ranks[i] = min(ranks[i], ranks[i - 1] * n)
    for i in range(d - 1, 0, -1):
        ranks[i] = min(ranks[i], ranks[i + 1] * n)
    return ranks


def rank1_rhs(d, n, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((1, n, 1, 1)) for _ in range(d)]


def random_guess(d, n, r, seed=1):
    rng = np.random.default_rng(seed)
    ranks = capped_ranks(d, n, r)
    return [rng.standard_normal((ranks[i], n, 1, ranks[i + 1])) for i in range(d)]


def workload_cores(d, n, r):
    """C3 family: (operator, right-hand side, un-orthonormalised guess)."""
    return laplace_cores(d, n), rank1_rhs(d, n, 0), random_guess(d, n, r, 1)


# ---------------------------------------------------------------------------------------------- C4
def c4_spd_cores(d, n=16, R=8, eps=0.1, seed=2):
    """Sum of R Kronecker products of SPD factors I + eps sym(randn) as a TT operator with rank-diagonal cores."""
    rng = np.random.default_rng(seed)
    sym = lambda a: 0.5 * (a + a.T)
    blocks = [[np.eye(n) + eps * sym(rng.standard_normal((n, n))) for _ in range(R)] for _ in range(d)]
    cores = []
    for k in range(d):
        Rl, Rr = (1 if k == 0 else R), (1 if k == d - 1 else R)
        c = np.zeros((Rl, n, n, Rr))
        for b in range(R):
            c[0 if k == 0 else b, :, :, 0 if k == d - 1 else b] = blocks[k][b]
        cores.append(c)
    return cores


def c4_random_cores(d, n=16, R=8, seed=2):
    """Throughput variant: dense Gaussian cores."""
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((1 if k == 0 else R, n, n, 1 if k == d - 1 else R)) for k in range(d)]


# ---------------------------------------------------------------------------------------------- C1
def cascade_cores(d):
    """models.signaling_cascade(d): the cascade repeats its middle core (checked against the live reference for d = 20 by
    tests/golden/make_config_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "euler_cascade.npz"))
    return [z["op/first"].copy()] + [z["op/mid"].copy() for _ in range(d - 2)] + [z["op/last"].copy()]


def cascade_initial_value(d, n=64):
    iv = [np.zeros((1, n, 1, 1)) for _ in range(d)]
    for c in iv:
        c[0, 0, 0, 0] = 1.0
    return iv


# ---------------------------------------------------------------------------------------------- C2 / C5
def co_oxidation_cores(d, k_ad_co):
    """models.co_oxidation(d, k_ad_co) (cyclic) for d >= 3: SLIM cores, affine in the CO adsorption rate."""
    z = np.load(os.path.join(GOLDEN, "co_oxidation_slim.npz"))
    at = lambda name: z[f"k0/{name}"] + k_ad_co * (z[f"k1/{name}"] - z[f"k0/{name}"])
    return [at("first")] + [at("mid") for _ in range(d - 2)] + [at("last")]


def c5_pressures(count=64):
    """CO pressures of the sweep (examples/co_oxidation.py:77 uses 10**(8+p)); SURVEY.md 8d: p = linspace(-4, 2, 64)."""
    return [10.0 ** (8 + p) for p in np.linspace(-4, 2, count)]


def add_identity(cores):
    """tt.eye(row_dims) + TT(cores) in the reference's block layout (tensor_train.py:282-344)."""
    d = len(cores)
    out = []
    for i, c in enumerate(cores):
        R, m, n, R2 = c.shape
        e = np.eye(m, n)
        if d == 1:
            out.append(c + e.reshape(1, m, n, 1))
        elif i == 0:
            out.append(np.concatenate([e.reshape(1, m, n, 1), c], axis=3))
        elif i == d - 1:
            out.append(np.concatenate([e.reshape(1, m, n, 1), c], axis=0))
        else:
            blk = np.zeros((R + 1, m, n, R2 + 1))
            blk[0, :, :, 0] = e
            blk[1:, :, :, 1:] = c
            out.append(blk)
    return out
