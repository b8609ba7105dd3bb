"""Generate golden vectors for the oracle from the LIVE reference (PGelss/scikit_tt).

Run in the build container only (the reference is not present on the GPU box):

    OPENBLAS_NUM_THREADS=1 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference \
        python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs) and versions.json.  Inputs are seeded;
every case feeds bit-identical initial-guess cores to the reference and, later, to the oracle and
the CUDA path (one-repeat ALS is sensitive to the null-space completion of the guess).
"""
import json
import os
import sys

import numpy as np
import scipy
import scipy.linalg

import scikit_tt.tensor_train as tt
from scikit_tt.tensor_train import TT
import scikit_tt.solvers.sle as sle
import scikit_tt.solvers.evp as evp
import scikit_tt.solvers.ode as ode
import scikit_tt.models as mdl

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(prefix, t, out):
    cores = t.cores if isinstance(t, TT) else t
    out[prefix + "/n"] = np.array(len(cores))
    for i, c in enumerate(cores):
        out[f"{prefix}/{i}"] = np.asarray(c)


def save(name, d):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    size = os.path.getsize(os.path.join(HERE, name + ".npz"))
    print(f"{name}.npz: {size / 1024:.1f} KiB, {len(d)} arrays")


def laplace_like(d, n, c=1e-3):
    """rank-3 SLIM-form operator  sum_i S_i + sum_i L_i M_{i+1}."""
    S = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    D = 0.5 * (np.eye(n, k=1) - np.eye(n, k=-1))
    L = M = np.sqrt(c) * D
    I = np.eye(n)
    Z = np.zeros((n, n))
    first = tt.build_core([[S, L, I]])
    mid = tt.build_core([[I, Z, Z], [M, Z, Z], [S, L, I]])
    last = tt.build_core([[I], [M], [S]])
    return TT([first] + [mid.copy() for _ in range(d - 2)] + [last])


def rand_tt(rng, row_dims, ranks, cplx=False):
    cores = []
    for i, n in enumerate(row_dims):
        c = rng.standard_normal((ranks[i], n, 1, ranks[i + 1]))
        if cplx:
            c = c + 1j * rng.standard_normal(c.shape)
        cores.append(c)
    return TT(cores)


def case_kernels():
    """Private per-micro-step functions of the reference on random (complex) inputs."""
    out = {}
    f = sle.__dict__
    g = evp.__dict__
    for tag, cplx in (("real", False), ("cplx", True)):
        rng = np.random.default_rng(11 if cplx else 7)
        d, n = 4, 3
        R = [1, 2, 3, 2, 1]
        r = [1, 3, 4, 2, 1]
        p = [1, 2, 2, 3, 1]
        mk = (lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)) if cplx else rng.standard_normal
        op = TT([mk((R[i], n, n, R[i + 1])) for i in range(d)])
        x = TT([mk((r[i], n, 1, r[i + 1])) for i in range(d)])
        b = TT([mk((p[i], n, 1, p[i + 1])) for i in range(d)])
        pack(f"{tag}/op", op, out); pack(f"{tag}/x", x, out); pack(f"{tag}/b", b, out)
        Lop, Rop, Lrhs, Rrhs = [None] * d, [None] * d, [None] * d, [None] * d
        for i in range(d):
            f["__construct_stack_left_op"](i, Lop, op, x)
            f["__construct_stack_left_rhs"](i, Lrhs, b, x)
        for i in range(d - 1, -1, -1):
            f["__construct_stack_right_op"](i, Rop, op, x)
            f["__construct_stack_right_rhs"](i, Rrhs, b, x)
        for i in range(d):
            out[f"{tag}/Lop/{i}"], out[f"{tag}/Rop/{i}"] = Lop[i], Rop[i]
            out[f"{tag}/Lrhs/{i}"], out[f"{tag}/Rrhs/{i}"] = Lrhs[i], Rrhs[i]
            out[f"{tag}/M1/{i}"] = f["__construct_micro_matrix_als"](i, Lop, Rop, op, x)
            out[f"{tag}/f1/{i}"] = f["__construct_micro_rhs_als"](i, Lrhs, Rrhs, b, x)
        for i in range(d - 1):
            out[f"{tag}/M2/{i}"] = f["__construct_micro_matrix_mals"](i, Lop, Rop, op, x)
            out[f"{tag}/f2/{i}"] = f["__construct_micro_rhs_mals"](i, Lrhs, Rrhs, b, x)
        # evp left stacks (conjugation on the column-side core, evp.py:281-283)
        class O:  # noqa
            pass
        trains, stacks = O(), O()
        trains.operator, trains.operator_gevp, trains.solution, trains.previous = op, None, x, []
        stacks.op_left, stacks.op_right = [None] * d, [None] * d
        stacks.op_gevp_left, stacks.op_gevp_right = [None] * d, [None] * d
        stacks.previous_left, stacks.previous_right = [], []
        for i in range(d):
            g["__construct_left_stacks"](i, trains, stacks)
            out[f"{tag}/evpL/{i}"] = stacks.op_left[i]
    save("kernels", out)


def case_sle_toeplitz():
    out = {}
    order = 10
    mat = scipy.linalg.toeplitz(np.arange(1, 2 ** order + 1), np.arange(1, 2 ** order + 1))
    # tests/test_sle.py:20-27 builds TT(mat.reshape(...)) without truncation (ranks up to 1024, 33 MB cores);
    # the fixture truncates the numerically-zero directions so that it stays small.
    op = TT(mat.reshape([2] * 2 * order), threshold=1e-14)
    rhs = tt.ones(op.row_dims, [1] * order)
    x0 = tt.ones(op.row_dims, [1] * order, ranks=5).ortho_right()
    pack("op", op, out); pack("rhs", rhs, out); pack("x0", x0, out)
    for solver in ("solve", "lu"):
        pack(f"als_{solver}", sle.als(op, x0, rhs, repeats=1, solver=solver), out)
        pack(f"mals_{solver}", sle.mals(op, x0, rhs, repeats=1, solver=solver, threshold=1e-14, max_rank=10), out)
    sol = TT([out[f"als_solve/{i}"] for i in range(order)])
    out["als_solve_residual"] = np.array((op.dot(sol) - rhs).norm() / rhs.norm())
    out["dense_solution"] = np.linalg.solve(mat, np.ones(mat.shape[0]))
    save("sle_toeplitz", out)


def case_sle_laplace():
    out = {}
    d, n, r = 6, 8, 4
    op = laplace_like(d, n)
    rhs = TT([np.random.default_rng(0).standard_normal((1, n, 1, 1)) for _ in range(d)])
    ranks = [1] + [r] * (d - 1) + [1]
    x0 = rand_tt(np.random.default_rng(1), [n] * d, ranks).ortho_right()
    pack("op", op, out); pack("rhs", rhs, out); pack("x0", x0, out)
    for rep in (1, 2):
        sol = sle.als(op, x0, rhs, repeats=rep)
        pack(f"als_rep{rep}", sol, out)
        out[f"als_rep{rep}_residual"] = np.array((op.dot(sol) - rhs).norm() / rhs.norm())
    sol = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=r)
    pack("mals", sol, out)
    out["mals_residual"] = np.array((op.dot(sol) - rhs).norm() / rhs.norm())
    save("sle_laplace", out)


def case_sle_random_spd():
    """SURVEY.md 8d C4 parity variant (rank-diagonal SPD cores), down-scaled."""
    out = {}
    d, n, R, r = 5, 6, 3, 5
    rng = np.random.default_rng(2)
    blocks = []
    for k in range(d):
        row = []
        for b in range(R):
            g = rng.standard_normal((n, n))
            row.append(np.eye(n) + 0.1 * 0.5 * (g + g.T))
        blocks.append(row)
    cores = [np.zeros((1, n, n, R))]
    for b in range(R):
        cores[0][0, :, :, b] = blocks[0][b]
    for k in range(1, d - 1):
        c = np.zeros((R, n, n, R))
        for b in range(R):
            c[b, :, :, b] = blocks[k][b]
        cores.append(c)
    c = np.zeros((R, n, n, 1))
    for b in range(R):
        c[b, :, :, 0] = blocks[d - 1][b]
    cores.append(c)
    op = TT(cores)
    rhs = TT([np.random.default_rng(3).standard_normal((1, n, 1, 1)) for _ in range(d)])
    x0 = rand_tt(np.random.default_rng(4), [n] * d, [1] + [r] * (d - 1) + [1]).ortho_right()
    pack("op", op, out); pack("rhs", rhs, out); pack("x0", x0, out)
    sol = sle.als(op, x0, rhs, repeats=2)
    pack("als_rep2", sol, out)
    out["als_rep2_residual"] = np.array((op.dot(sol) - rhs).norm() / rhs.norm())
    sol = sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=r)
    pack("mals", sol, out)
    save("sle_random_spd", out)


def case_sle_complex():
    out = {}
    d, n, R, r = 4, 4, 2, 3
    rng = np.random.default_rng(5)
    mk = lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    cores = [mk((1 if i == 0 else R, n, n, 1 if i == d - 1 else R)) for i in range(d)]
    for c in cores:  # keep the operator comfortably non-singular
        for a in range(c.shape[0]):
            for b in range(c.shape[3]):
                c[a, :, :, b] += 3 * np.eye(n) if a == b or c.shape[0] == 1 or c.shape[3] == 1 else 0
    op = TT(cores)
    rhs = TT([mk((1, n, 1, 1)) for _ in range(d)])
    x0 = TT([mk((1 if i == 0 else r, n, 1, 1 if i == d - 1 else r)) for i in range(d)]).ortho_right()
    pack("op", op, out); pack("rhs", rhs, out); pack("x0", x0, out)
    pack("als_rep2", sle.als(op, x0, rhs, repeats=2), out)
    pack("mals", sle.mals(op, x0, rhs, repeats=1, threshold=1e-12, max_rank=r), out)
    save("sle_complex", out)


def case_euler_cascade():
    out = {}
    d, r, steps = 4, 3, 3
    op = mdl.signaling_cascade(d)
    iv = tt.zeros(op.col_dims, [1] * d)
    for p in range(d):
        iv.cores[p][0, 0, 0, 0] = 1
    guess = tt.ones(op.col_dims, [1] * d, ranks=r).ortho_right()
    # the operator is sparse; store only the three distinct cores
    out["op/first"], out["op/mid"], out["op/last"] = op.cores[0], op.cores[1], op.cores[-1]
    out["d"] = np.array(d)
    pack("iv", iv, out); pack("guess", guess, out)
    sol = ode.implicit_euler(op, iv, guess, [1.0] * steps, repeats=1, tt_solver='als', progress=False)
    for k in range(1, steps + 1):
        pack(f"als/step{k}", sol[k], out)
    # MALS needs small mode sizes (two-site micro matrix); use a small diffusion generator -Laplacian
    d2, n2, r2 = 5, 4, 3
    gen = (-1.0) * laplace_like(d2, n2, c=0.0)
    iv2 = TT([np.random.default_rng(9).random((1, n2, 1, 1)) for _ in range(d2)])
    g2 = rand_tt(np.random.default_rng(10), [n2] * d2, [1] + [r2] * (d2 - 1) + [1]).ortho_right()
    pack("gen", gen, out); pack("iv2", iv2, out); pack("guess2", g2, out)
    for norm_p in (1, 2, 0):
        sol = ode.implicit_euler(gen, iv2, g2, [0.1, 0.2, 0.1], repeats=2, tt_solver='mals', max_rank=r2,
                                 normalize=norm_p, progress=False)
        for k in range(1, 4):
            pack(f"mals_norm{norm_p}/step{k}", sol[k], out)
    save("euler_cascade", out)


def case_evp():
    out = {}
    d, n, r = 5, 6, 4
    op = laplace_like(d, n, c=0.05)
    x0 = rand_tt(np.random.default_rng(6), [n] * d, [1] + [r] * (d - 1) + [1]).ortho_right()
    pack("op", op, out); pack("x0", x0, out)
    lam, x, it = evp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    out["eigh/lam"], out["eigh/it"] = np.array(lam), np.array(it)
    pack("eigh/x", x, out)
    lam, x, it = evp.als(op, x0, repeats=4, conv_eps=0, solver='eig', sigma=0.0)
    out["eig/lam"], out["eig/it"] = np.array(lam), np.array(it)
    pack("eig/x", x, out)
    lam, x, it = evp.als(op, x0, repeats=3, conv_eps=0, solver='eigh', number_ev=2)
    out["eigh2/lam"] = np.array(lam)
    for j in range(2):
        pack(f"eigh2/x{j}", x[j], out)
    # generalised problem with an SPD right-hand operator
    gev = tt.eye(op.row_dims) + 0.1 * laplace_like(d, n, c=0.0)
    pack("gevp", gev, out)
    lam, x, it = evp.als(op, x0, operator_gevp=gev, repeats=3, conv_eps=0, solver='eigh')
    out["gevp_eigh/lam"] = np.array(lam)
    pack("gevp_eigh/x", x, out)
    # deflation: second eigenpair via shift of the first (evp.py:376-381)
    lam1, x1, _ = evp.als(op, x0, repeats=4, conv_eps=0, solver='eigh')
    lam2, x2, _ = evp.als(op, x0, previous=[x1], shift=-lam1, repeats=4, conv_eps=0, solver='eigh')
    out["defl/lam1"], out["defl/lam2"] = np.array(lam1), np.array(lam2)
    pack("defl/x1", x1, out)
    save("evp_laplace", out)


def case_evp_cooxidation():
    out = {}
    d = 8
    op = mdl.co_oxidation(d, 1e4).ortho_left().ortho_right()
    pack("op_raw", op, out)
    opI = tt.eye(op.row_dims) + op
    x0 = tt.ones(op.row_dims, [1] * d, ranks=4).ortho_left().ortho_right()
    pack("x0", x0, out)
    lam, x, it = evp.als(opI, x0, repeats=5, conv_eps=0, solver='eig', sigma=1)
    out["eig/lam"], out["eig/it"] = np.array(lam), np.array(it)
    pack("eig/x", x, out)
    save("evp_cooxidation", out)


def case_ortho():
    out = {}
    rng = np.random.default_rng(8)
    t = TT([rng.standard_normal(s) for s in ((1, 3, 2, 4), (4, 4, 1, 5), (5, 5, 3, 3), (3, 2, 2, 1))])
    pack("t", t, out)
    pack("left", t.copy().ortho_left(), out)
    pack("right", t.copy().ortho_right(), out)
    pack("left_thr", t.copy().ortho_left(threshold=0.2), out)
    pack("right_mr", t.copy().ortho_right(max_rank=2), out)
    pack("ortho", t.copy().ortho(threshold=1e-12, max_rank=3), out)
    out["norm2"], out["norm1"] = np.array(t.norm(p=2)), np.array(TT([np.abs(c) for c in t.cores]).norm(p=1))
    # rank-deficient train (ones with rank 4 is numerically rank 1)
    o = tt.ones([3, 4, 5], [1, 1, 1], ranks=4)
    pack("ones", o, out)
    pack("ones_right", o.copy().ortho_right(threshold=1e-10), out)
    tc = TT([rng.standard_normal(s) + 1j * rng.standard_normal(s) for s in ((1, 3, 1, 3), (3, 4, 1, 2), (2, 3, 1, 1))])
    pack("tc", tc, out)
    pack("tc_left", tc.copy().ortho_left(), out)
    pack("tc_right", tc.copy().ortho_right(), out)
    save("ortho", out)


if __name__ == "__main__":
    np.seterr(all="ignore")
    case_kernels()
    case_sle_toeplitz()
    case_sle_laplace()
    case_sle_random_spd()
    case_sle_complex()
    case_euler_cascade()
    case_evp()
    case_evp_cooxidation()
    case_ortho()
    import numpy, scipy
    json.dump({"numpy": numpy.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
               "openblas_threads": os.environ.get("OPENBLAS_NUM_THREADS", "default"),
               "reference": "PGelss/scikit_tt @ /root/reference (read-only mount)"},
              open(os.path.join(HERE, "versions.json"), "w"), indent=1)
