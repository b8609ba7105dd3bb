"""Golden vectors for the BASELINE.json configurations at the sizes the reference can still run, from the LIVE reference
(PGelss/scikit_tt).  Build container only (the reference does not exist on the GPU box):

    OPENBLAS_NUM_THREADS=8 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_config_golden.py [case ...]

Cases (one .npz each):
  c1_full        config 1 at FULL size: models.signaling_cascade(20), implicit Euler via sle.als, solution rank 4 -- the three
                 distinct operator cores are already in euler_cascade.npz (the cascade repeats its middle core; checked
                 here), this file adds the rank-4 guess and the reference's states after steps 1..3.
  c4_parity      config 4 parity variant (SPD, rank-diagonal cores, n=16, R=8): sle.als at r=32 (dense
                 16 384^2 micro systems, the largest the reference factorises in bounded time) and sle.mals at r<=8
                 (two-site systems of 16 384 unknowns).
  c2_first_step  config 2: guesses of rank 16 / 24 / 32 and the eigenvalue the reference's FIRST micro step selects
                 (`lin.eig` of the 768^2 / 1728^2 / 3072^2 micro matrix); later micro steps are chaotic (test_c2_noise.py).
  co_oxidation_slim  the three distinct SLIM cores of models.co_oxidation(20, k) at k = 0 and k = 1; the cores are affine
                 in k (checked here against k = 1e4, 10**5.5), so every CO pressure of config 5 is first + k * (second - first).
"""
import os
import sys

import numpy as np

import scikit_tt.tensor_train as tt
from scikit_tt.tensor_train import TT
import scikit_tt.solvers.sle as sle
import scikit_tt.solvers.evp as evp
import scikit_tt.solvers.ode as ode
import scikit_tt.models as mdl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import workloads  # noqa: E402  (repo root: the synthetic operator families shared by bench.py and the tests)


def pack(prefix, t, out):
    cores = t.cores if isinstance(t, TT) else t
    out[prefix + "/n"] = np.array(len(cores))
    for i, c in enumerate(cores):
        out[f"{prefix}/{i}"] = np.asarray(c)


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}.npz: {os.path.getsize(path) / 1024:.1f} KiB, {len(d)} arrays", flush=True)


def case_c1_full():
    z = np.load(os.path.join(HERE, "euler_cascade.npz"))
    d = 20
    op = mdl.signaling_cascade(d)
    assert np.array_equal(op.cores[0], z["op/first"]) and np.array_equal(op.cores[-1], z["op/last"])
    assert all(np.array_equal(op.cores[i], z["op/mid"]) for i in range(1, d - 1))
    iv = tt.zeros(op.col_dims, [1] * d)
    for p in range(d):
        iv.cores[p][0, 0, 0, 0] = 1
    guess = tt.ones(op.col_dims, [1] * d, ranks=4).ortho_right()          # examples/signaling_cascade.py:73-85
    out = {"d": np.array(d)}
    pack("guess", guess, out)
    sol = ode.implicit_euler(op, iv, guess, [1.0] * 3, repeats=1, tt_solver='als', progress=False)
    for k in range(1, 4):
        pack(f"step{k}", sol[k], out)
    save("c1_full", out)


def case_c4_parity():
    out = {}
    d, n, R, r = 6, 16, 8, 32
    opc = workloads.c4_spd_cores(d, n, R)
    rhsc = workloads.rank1_rhs(d, n)
    op, rhs = TT([c.copy() for c in opc]), TT([c.copy() for c in rhsc])
    x0 = TT(workloads.random_guess(d, n, r, seed=1)).ortho_right()
    out["als/d"], out["als/r"] = np.array(d), np.array(r)
    pack("als/x0", x0, out)
    sol = sle.als(op, x0, rhs, repeats=1)
    pack("als/x", sol, out)
    out["als/residual"] = np.array((op.dot(sol) - rhs).norm() / rhs.norm())
    print("c4 als ranks", sol.ranks, "residual", float(out["als/residual"]), flush=True)
    d2, r2 = 5, 8
    op2, rhs2 = TT(workloads.c4_spd_cores(d2, n, R)), TT(workloads.rank1_rhs(d2, n))
    x02 = TT(workloads.random_guess(d2, n, r2, seed=3)).ortho_right()
    out["mals/d"], out["mals/r"] = np.array(d2), np.array(r2)
    pack("mals/x0", x02, out)
    sol2 = sle.mals(op2, x02, rhs2, repeats=1, threshold=1e-12, max_rank=r2)
    pack("mals/x", sol2, out)
    out["mals/residual"] = np.array((op2.dot(sol2) - rhs2).norm() / rhs2.norm())
    print("c4 mals ranks", sol2.ranks, "residual", float(out["mals/residual"]), flush=True)
    save("c4_parity", out)


class _Stop(Exception):
    pass


def case_c2_first_step():
    z = np.load(os.path.join(HERE, "c2_cooxidation20.npz"))
    op = TT([z[f"op/{i}"] for i in range(int(z["op/n"]))])                 # I + co_oxidation(20, 1e4), orthonormalised
    out = {}
    orig = evp.lin.eig
    # micro steps recorded per guess rank: the whole forward half sweep at r = 16 / 24, at r = 32 up to the second
    # full-size (3072^2) micro matrix (one dgeev of that size takes about a minute on the host)
    for r, limit in ((16, 19), (24, 19), (32, 5)):
        guess = tt.ones(op.row_dims, [1] * op.order, ranks=r).ortho_left().ortho_right()
        pack(f"r{r}/x0", guess, out)
        seen = []

        def spy(M, *a, **kw):
            w, v = orig(M, *a, **kw)
            sel = w[np.argsort(np.abs(w - 1))[0]]                          # evp.py:427-432 with sigma = 1
            seen.append((M.shape[0], sel, np.sort(np.abs(w - sel))[1], np.abs(M).max()))
            if len(seen) >= limit:
                raise _Stop()
            return w, v
        evp.lin.eig = spy
        try:
            evp.als(op, guess, repeats=1, conv_eps=0, solver='eig')
        except _Stop:
            pass
        finally:
            evp.lin.eig = orig
        out[f"r{r}/N"] = np.array([s[0] for s in seen])
        out[f"r{r}/lam"] = np.array([s[1] for s in seen])
        out[f"r{r}/gap"] = np.array([s[2] for s in seen])                  # distance to the nearest other eigenvalue
        out[f"r{r}/absmax"] = np.array([s[3] for s in seen])               # max |M_ij|: eigenvalues are defined to ~eps * this
        for s in seen:
            print(f"c2 r={r}: N={s[0]} lambda={s[1]} gap={s[2]:.3e} max|M|={s[3]:.3e}", flush=True)
    save("c2_first_step", out)


def case_co_oxidation_slim():
    a, b = mdl.co_oxidation(20, 0.0), mdl.co_oxidation(20, 1.0)
    for k in (1e4, 10 ** 5.5, 10 ** 10):
        c = mdl.co_oxidation(20, k)
        for i in range(20):
            assert np.array_equal(a.cores[i] + k * (b.cores[i] - a.cores[i]), c.cores[i]), (k, i)
    assert all(np.array_equal(a.cores[1], a.cores[i]) and np.array_equal(b.cores[1], b.cores[i]) for i in range(1, 19))
    out = {}
    for tag, t in (("k0", a), ("k1", b)):
        out[f"{tag}/first"], out[f"{tag}/mid"], out[f"{tag}/last"] = t.cores[0], t.cores[1], t.cores[-1]
    save("co_oxidation_slim", out)


if __name__ == "__main__":
    np.seterr(all="ignore")
    cases = {"c1_full": case_c1_full, "c4_parity": case_c4_parity, "c2_first_step": case_c2_first_step,
             "co_oxidation_slim": case_co_oxidation_slim}
    for name in (sys.argv[1:] or list(cases)):
        cases[name]()
