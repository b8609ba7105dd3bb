#!/usr/bin/env python
"""bench.py -- ALS half-sweeps/s (fp64) on the BASELINE.json workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs same,c5,c4]

Headline (config.workload = "C3", the configuration BASELINE.json's metric is quoted on): synthetic discretised-Laplacian
TT operator d=32, n=64, operator rank 3, right-hand side rank 1 (seed 0), initial guess with interior
solution rank 64 (seed 1, right-orthonormalised).  One *step* = one `sle.als(op, x0, rhs, repeats=1)` = 2 half-sweeps over
the 32 cores: 64 interface-stack updates, 63 micro systems of 262 144 unknowns solved matrix-free (CG to a TRUE relative
residual of 1e-14, warm-started from the sweep's iterate; the reference's dense micro matrix would be 512 GiB), 62 QR/RQ
factorisations of 4096 x 64 unfoldings.

`value`   : half-sweeps/s with operator, right-hand side and guess resident in HBM (CUDA events, max over ranks).
`e2e`     : the same call through the public API with host numpy TT cores in and out (H2D + D2H inside the timed region).
`roofline`: the kernel the metric names -- the interface-stack update behind the C-ABI entry point sktt_stack_left_op --
            timed with CUDA events right after the timed region: F = 2 r^3 R (n+m) + 2 r^2 R^2 m n flops per update against
            the measured fp64 tensor-pipe peak (profiles/r01_fp64_peaks.txt).  Sub-fields: `pcg_in_situ`, the persistent CG
            kernel AS IT RUNS IN THE SWEEP (flops = matvecs counted on the device x F, time = CUDA events around every one
            of the step's solves), and `matvec_loop`, the matvec phases alone inside one cooperative launch.
`same_config`: the GPU arm on the configuration the reference arm can run (same operator family at solution rank
            --sample-rank, dense micro systems + LU on both sides) so that the driver's ratio has a same-work counterpart.
`c1`       : BASELINE configuration 0 at full size (signaling cascade d=20, implicit Euler via sle.als, rank 4) against the
            reference itself on the host cores, with the parity of the first time step.
`c5`, `c4`: the two configurations that shard (north_star): the batch of 64 co_oxidation(20) eigenproblems, block-sharded
            over the ranks with the systems of a rank batched inside its GPU, and one sle.als sweep of the d=10, n=16, R=8
            operator at solution rank 256 (--c4-rank) whose micro-matvec is rank-sharded over the ranks.  With --gpus N these legs are
            the strong-scaling curves (total work fixed); the headline stays N independent C3 replicas (the single C3
            system is sequential in the core index and does not shard: SURVEY.md 8e "replicas only").
`cpu_baseline` / `--impl reference`: the reference's own CPU implementation on the host cores -- PGelss/scikit_tt itself when
            __graft_entry__.build() could install it into oracle/_ref (kind "reference"), else the numpy/scipy restatement
            under oracle/ (kind "port") -- on a bounded sample: the same operator family at the largest solution rank whose
            dense micro matrices the reference factorises in bounded time.  Its `config` states the rank it really ran.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from workloads import laplace_cores, workload_cores  # noqa: E402,F401  (kept importable from here: tests, tools)

FP64_TENSOR_PEAK_TFLOPS = 37.1      # measured on this pool's B200: profiles/r01_fp64_peaks.txt (DMMA m8n8k4, sustained)
NCU_STACK_DRAM_BYTES = None         # per launch of stack_nat_kernel, from an ncu --set full capture (profiles/r02_stack_traffic.json)
_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_stack_traffic.json")
if os.path.exists(_TRAFFIC_FILE):
    try:
        NCU_STACK_DRAM_BYTES = json.load(open(_TRAFFIC_FILE)).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        NCU_STACK_DRAM_BYTES = None


def stack_flops(r, R, n, r2, R2):
    return 2 * r * R * r * n * r2 + 2 * r * r2 * R * R2 * n * n + 2 * r2 * R2 * r * n * r2


def stack_bytes(r, R, n, r2, R2):
    """B_stack = 8 (r^2 R + r n r' + R m n R' + r'^2 R')."""
    return 8 * (r * r * R + r * n * r2 + R * n * n * R2 + r2 * r2 * R2)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            if len(row) < 7:
                continue
            try:
                sm.append(float(row[0]))
                smax.append(float(row[1]))
            except ValueError:
                continue
            for name, val in zip(names, row[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _reference_modules():
    """PGelss/scikit_tt itself from oracle/_ref (installed there by __graft_entry__.build() in the build container, where
    /root/reference exists; git-ignored, travels to the GPU box with the snapshot), or None."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "scikit_tt")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import io
        import contextlib
        import warnings
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from scikit_tt.tensor_train import TT as RTT
            import scikit_tt.solvers.sle as rsle
            import scikit_tt.solvers.evp as revp
        return RTT, rsle, revp
    except Exception:
        return None


def cpu_sample(d, n, sample_rank, steps=1, warmup=0):
    """The reference algorithm on the bounded C3 sample, all BLAS threads.  Returns (half-sweeps/s, s/step, kind)."""
    op, rhs, x0 = workload_cores(d, n, sample_rank)
    from oracle import tt as ott
    x0 = ott.ortho_right(x0)
    mods = _reference_modules()
    if mods is not None:
        RTT, rsle, _ = mods
        A, b, g = RTT([c.copy() for c in op]), RTT([c.copy() for c in rhs]), RTT([c.copy() for c in x0])
        run = lambda: rsle.als(A, g, b, repeats=1)
        kind = "reference"
    else:
        from oracle import sle as osle
        run = lambda: osle.als(op, x0, rhs, repeats=1)
        kind = "port"
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return 2 * steps / dt, dt / steps, kind


def cpu_sample_c5(rank=8, repeats=1, solver='eigs', pressures=2):
    """One-core reference timing of config 5 items (thread thrash makes one BLAS thread the reference's best setting at
    these sizes).  Returns (half-sweeps/s per core, kind)."""
    from threadpoolctl import threadpool_limits
    from oracle import tt as ott
    d = 20
    ks = workloads.c5_pressures(64)[:: max(1, 64 // pressures)][:pressures]
    items = []
    for k in ks:
        cores = ott.ortho_right(ott.ortho_left(workloads.co_oxidation_cores(d, k)))
        items.append(workloads.add_identity(cores))
    guess = ott.ortho_right(ott.ortho_left([np.ones((1 if i == 0 else rank, 3, 1, 1 if i == d - 1 else rank)) for i in range(d)]))
    mods = _reference_modules()
    with threadpool_limits(limits=1):
        if mods is not None:
            RTT, _, revp = mods
            g = RTT([c.copy() for c in guess])
            tts = [RTT([c.copy() for c in it]) for it in items]
            run = lambda j: revp.als(tts[j], g, repeats=repeats, conv_eps=0, solver=solver)
            kind = "reference"
        else:
            from oracle import evp as oevp
            run = lambda j: oevp.als(items[j], guess, repeats=repeats, conv_eps=0, solver=solver)
            kind = "port"
        t0 = time.perf_counter()
        for j in range(len(items)):
            run(j)
        dt = time.perf_counter() - t0
    return 2 * repeats * len(items) / dt, kind


def cfg_public(cfg, rank=None):
    r = cfg["r"] if rank is None else rank
    dense = r * cfg["n"] * r <= 16384
    return {"workload": "C3: sle.als on the rank-3 Laplacian-type TT operator", "d": cfg["d"], "n": cfg["n"],
            "operator_rank": 3, "solution_rank": r, "repeats_per_step": 1, "half_sweeps_per_step": 2,
            "micro_solver": ("dense micro matrix + LU with partial pivoting, as the reference" if dense else
                             "matrix-free CG (Chronopoulos-Gear form, reductions fused into the matvec, warm start from the "
                             "sweep's current iterate), true relative residual 1e-14 (dense micro matrix impossible at this size)"),
            "l2": "256 MiB buffer written between steps (inside the timed region)",
            "parallelism": "replicas" if cfg["gpus"] > 1 else "single GPU"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    warm = min(args.warmup, 3)                                # the same warm-up as the GPU arm up to three steps (~2 s each)
    steps = max(1, min(args.steps, args.ref_max_steps))
    val, per_step, kind = cpu_sample(cfg["d"], cfg["n"], args.sample_rank, steps=steps, warmup=warm)
    sample = (f"same operator family (d={cfg['d']}, n={cfg['n']}, R=3) at solution rank {args.sample_rank}: dense micro "
              f"matrices of {args.sample_rank ** 2 * cfg['n']}^2 + LU as the reference does ({steps} timed steps of "
              f"{per_step:.2f} s, {warm} warm-up); the named rank {cfg['r']} needs a 512 GiB micro matrix and cannot run. "
              f"The GPU arm reports the same configuration under `same_config`.")
    public = cfg_public(cfg, rank=args.sample_rank)
    public["parallelism"] = "host cores (one process)"
    public["l2"] = "n/a (CPU)"
    line = {"impl": "reference", "metric": "ALS half-sweeps/s (fp64)", "value": val, "unit": "half-sweeps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": public,
            "cpu_baseline": {"value": val, "unit": "half-sweeps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "half-sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # the version banner goes to stdout, next to the JSON line
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from scikit_tt_b200._device import get_device
        self.dev = get_device()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """`warmup` untimed calls, then `steps` calls between a barrier + synchronize on both sides; CUDA events on the
        current stream and the host clock, each max over ranks.  Returns (event ms, wall ms) for the `steps` calls."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        wall = (time.perf_counter() - t_wall) * 1e3
        return self.max_over_ranks(ms), self.max_over_ranks(wall)

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t[0])


def leg_c3(env, args, cfg):
    torch = env.torch
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import sle, _local
    dev = env.dev
    d, n, r = cfg["d"], cfg["n"], cfg["r"]
    opc, rhsc, x0c = workload_cores(d, n, r)
    op, rhs = TT(opc).pin_memory(), TT(rhsc).pin_memory()     # inputs of the end-to-end leg lie in page-locked host memory
    x0 = TT(x0c).ortho_right()                                # GPU ortho path (TT.ortho_right); its cores come back page-locked
    st = sle._State(op, x0, rhs)
    x0_dev = list(st.x)

    def step_resident():
        env.flush.zero_()
        st.reset(x0_dev)
        sle._run_als(st, 1, args.solver)

    result = {}

    def step_e2e():
        env.flush.zero_()
        result["x"] = sle.als(op, x0, rhs, repeats=1, solver=args.solver)

    for _ in range(args.warmup):
        step_resident()
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    launches0 = dev.launches()
    ms, _ = env.timed(step_resident, args.steps, 0)
    launches = env.sum_over_ranks(dev.launches() - launches0)
    clocks = sampler.stop() if env.rank == 0 else None

    # end-to-end through the public API (host TT in, host TT out)
    ms_e2e_ev, ms_e2e_wall = env.timed(step_e2e, args.e2e_steps, 2)
    ms_e2e = max(ms_e2e_ev, ms_e2e_wall)
    krylov = dict(_local.stats)

    # ---- roofline 1: the persistent CG kernel in situ -- one more resident step with CUDA events around every micro solve
    solves = []
    orig_async = dev.krylov_solve_refined_async

    def spy(op_, f, u, result_, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ok = orig_async(op_, f, u, result_, **kw)
        e1.record()
        if ok:
            solves.append((e0, e1, result_, (op_.r, op_.R, op_.n, op_.r3, op_.R2)))
        return ok
    dev.krylov_solve_refined_async = spy
    try:
        st.reset(x0_dev)
        sle._run_als(st, 1, args.solver)
        torch.cuda.synchronize()
    finally:
        del dev.krylov_solve_refined_async
    pcg = None
    if solves:
        tot_ms, tot_flops, tot_mv, tot_it = 0.0, 0.0, 0, 0
        for e0, e1, res, (rr, RR, nn, r3, R2) in solves:
            o = res.cpu().numpy()
            mv = 1 + int(o[0]) + int(o[2])                    # initial true residual + CG iterations + one check per cycle
            tot_ms += e0.elapsed_time(e1)
            tot_flops += mv * stack_flops(rr, RR, nn, r3, R2)
            tot_mv += mv
            tot_it += int(o[0])
        pcg = {"kernel": "pcg_persistent_kernel, every launch of one timed-shape step (CUDA events around each launch)",
               "solves": len(solves), "cg_iterations": tot_it, "matvecs": tot_mv, "ms": tot_ms,
               "share_of_step": tot_ms / (ms / args.steps), "achieved": tot_flops / (tot_ms * 1e-3) / 1e12,
               "frac": tot_flops / (tot_ms * 1e-3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS,
               "note": "dense-formula flops of SURVEY.md 8d per matvec (edge cores with their own smaller F); the kernel "
                       "time also covers the CG vector updates, grid-wide reductions and true-residual checks"}

    # ---- roofline 2: the matvec phases alone (one cooperative launch doing `reps` matvecs)
    i = d // 2
    L, A, R = st.Lop[i], st.A[i], st.Rop[i]
    F = stack_flops(L.shape[0], A.shape[0], A.shape[2], R.shape[0], A.shape[3])
    lop = dev.local_op(L, A, R, prepare=True)
    nt = dev.tiled_len(lop)
    loop = None
    if nt > 0:
        reps = 200
        vt = torch.randn(nt, dtype=torch.float64, device="cuda")
        vt.view(-1, 68)[:, 64:] = 0.0
        yt = torch.zeros(nt, dtype=torch.float64, device="cuda")
        mv = lambda: dev.local_matvec_tiled_repeat(lop, vt, yt, reps)
        for _ in range(3):
            mv()
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(5):
            mv()
        m1.record()
        torch.cuda.synchronize()
        mv_ms = m0.elapsed_time(m1) / (5 * reps)
        Ah = opc[i]
        nnz = int(sum(bool(np.any(Ah[b, :, :, q])) for b in range(Ah.shape[0]) for q in range(Ah.shape[3])))
        rr_, R_, n_, r2_, R2_ = L.shape[0], A.shape[0], A.shape[2], R.shape[0], A.shape[3]
        F_exec = 2 * rr_ * R_ * rr_ * n_ * r2_ + 2 * rr_ * r2_ * nnz * n_ * n_ + 2 * r2_ * R2_ * rr_ * n_ * r2_
        loop = {"kernel": "pcg_persistent_kernel, matvec phases only (200 matvecs in one cooperative launch)",
                "us_per_matvec": mv_ms * 1e3, "achieved": F / (mv_ms * 1e-3) / 1e12,
                "frac": F / (mv_ms * 1e-3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS,
                "achieved_executed": F_exec / (mv_ms * 1e-3) / 1e12,
                "executed_note": f"the kernel skips the zero (b, b') blocks of the operator core ({nnz} of "
                                 f"{A.shape[0] * A.shape[3]} non-zero): achieved_executed counts the flops the tensor pipe ran"}

    # ---- roofline headline: the interface-stack update entry point (the metric's kernel)
    xs, Ls = st.x[i - 1], st.Lop[i - 1]
    for _ in range(5):
        dev.stack_left_op(Ls, xs, st.A[i - 1])
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(50):
        dev.stack_left_op(Ls, xs, st.A[i - 1])
    s1.record()
    torch.cuda.synchronize()
    stack_ms = s0.elapsed_time(s1) / 50
    Ash = st.A[i - 1].shape
    F_stack = stack_flops(Ls.shape[0], Ash[0], Ash[2], xs.shape[2], Ash[3])
    B_stack = stack_bytes(Ls.shape[0], Ash[0], Ash[2], xs.shape[2], Ash[3])
    achieved = F_stack / (stack_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP64_TENSOR_PEAK_TFLOPS, "traffic": NCU_STACK_DRAM_BYTES,
                "kernel": "sktt_stack_left_op at r=64, R=3, n=64 = ONE launch of stack_nat_kernel (csrc/stack_nat.cu: operands read "
                          "in their natural layout, no image build / tiling pass): the interface-stack update the metric names",
                "flops_per_update": F_stack, "us_per_update": stack_ms * 1e3, "algorithmic_bytes": B_stack,
                "arithmetic_intensity": F_stack / B_stack,
                "peak_source": "measured fp64 DMMA pipe peak, profiles/r01_fp64_peaks.txt (MEASURED_PEAKS.json has no fp64 entry)",
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one stack_nat_kernel launch under ncu --set full, "
                                "profiles/r02_stack_traffic.json (source capture named there); the operands are L2-resident in the "
                                "sweep, the op is compute-bound (AI ~ 194 flop/B)",
                "pcg_in_situ": pcg, "matvec_loop": loop}

    out = {"ms": ms, "ms_e2e": ms_e2e, "launches": launches, "clocks": clocks, "roofline": roofline, "krylov": krylov}
    if env.rank == 0:
        sol = result["x"]
        out["h2d"] = sum(c.nbytes for t in (op, x0, rhs) for c in t.cores)
        out["d2h"] = sum(c.nbytes for c in sol.cores)
        from scikit_tt_b200 import tensor_train as ttm          # || A x - b || / || b ||, core-wise QR evaluation
        out["residual"] = float(ttm.residual_error(op, sol, rhs) / np.prod([np.linalg.norm(c) for c in rhs.cores]))
    return out


def leg_same_config(env, args, cfg):
    """The reference arm's configuration on the GPU: same family at solution rank --sample-rank (dense micro systems, LU)."""
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import sle
    d, n, r = cfg["d"], cfg["n"], args.sample_rank
    opc, rhsc, x0c = workload_cores(d, n, r)
    op, rhs = TT(opc).pin_memory(), TT(rhsc).pin_memory()
    x0 = TT(x0c).ortho_right()
    st = sle._State(op, x0, rhs)
    x0_dev = list(st.x)

    def step_resident():
        env.flush.zero_()
        st.reset(x0_dev)
        sle._run_als(st, 1, 'solve')

    def step_e2e():
        env.flush.zero_()
        sle.als(op, x0, rhs, repeats=1, solver='solve')
    steps = max(2, min(args.steps, 5))
    ms, _ = env.timed(step_resident, steps, 2)
    ms_e, wall_e = env.timed(step_e2e, steps, 1)
    return {"config": cfg_public(cfg, rank=r), "steps": steps,
            "value": 2 * steps * env.world / (ms * 1e-3), "unit": "half-sweeps/s",
            "e2e": {"value": 2 * steps * env.world / (max(ms_e, wall_e) * 1e-3), "unit": "half-sweeps/s"},
            "note": "what --impl reference runs (dense micro systems + LU at this rank), here on the GPU"}


def leg_c1(env, args):
    """BASELINE config 0 ("the repo's own example, runs on CPU") at FULL size: signaling_cascade(d=20) CME operator (first /
    middle / last core of the live reference's operator, tests/golden/euler_cascade.npz; the cascade repeats its middle core),
    ode.implicit_euler via sle.als at solution rank 4 -- dense 1024 x 1024 micro systems, LU with partial pivoting on both
    sides.  Timed through the public API with host TT objects in and out (end to end by construction)."""
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import ode
    from oracle import tt as ott
    z = np.load(os.path.join(ROOT, "tests", "golden", "euler_cascade.npz"))
    d, n, steps = 20, 64, 2
    opc = [z["op/first"]] + [z["op/mid"].copy() for _ in range(d - 2)] + [z["op/last"]]
    iv = [np.zeros((1, n, 1, 1)) for _ in range(d)]
    for c in iv:
        c[0, 0, 0, 0] = 1.0
    ranks = [1] + [4] * (d - 1) + [1]
    guess = ott.ortho_right([np.ones((ranks[i], n, 1, ranks[i + 1])) for i in range(d)])
    out = {}

    def run():
        env.flush.zero_()
        out["sol"] = ode.implicit_euler(TT(opc), TT(iv), TT(guess), [1.0] * steps, progress=False)
    reps = 3
    ms, wall = env.timed(run, reps, 1)
    ms = max(ms, wall)
    leg = {"workload": "C1: ode.implicit_euler(signaling_cascade(20), sle.als, solution rank 4), 2 time steps = 4 half-sweeps per "
                       "call, dense 1024^2 micro systems + LU with partial pivoting (one cooperative launch per solve)",
           "value": 2 * steps * reps * env.world / (ms * 1e-3), "unit": "half-sweeps/s", "ms_per_call": ms / reps, "n_gpus": env.world,
           "scaling": "weak", "parallelism": "replicas" if env.world > 1 else "single GPU",
           "e2e_note": "timed through the public API with host numpy cores in and out every call"}
    if env.rank == 0 and env.world == 1 and not args.no_cpu:
        mods = _reference_modules()
        t0 = time.perf_counter()
        if mods is not None:
            RTT = mods[0]
            import io
            import contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                import scikit_tt.solvers.ode as rode
                ref = rode.implicit_euler(RTT([c.copy() for c in opc]), RTT([c.copy() for c in iv]), RTT([c.copy() for c in guess]),
                                          [1.0], progress=False)
            ref1, kind = ref[1].cores, "reference"
        else:
            from oracle import ode as oode
            ref1, kind = oode.implicit_euler(opc, iv, guess, [1.0])[1], "port"
        dt = time.perf_counter() - t0
        a = [np.asarray(c) for c in out["sol"][1].cores]
        b = [np.asarray(c) for c in ref1]
        leg["cpu_baseline"] = {"value": 2 / dt, "unit": "half-sweeps/s", "cores": len(os.sched_getaffinity(0)), "kind": kind,
                               "sample": "the first time step of the same call (2 half-sweeps), all BLAS threads"}
        leg["step1_rel_diff_vs_reference"] = float(ott.norm(ott.sub(a, b)) / ott.norm(b))
    return leg


def leg_c5(env, args):
    """BASELINE config 5: 64 co_oxidation(20) pressures, evp.als (solver as examples/co_oxidation.py:104), rank-8 guess.
    The 64 systems are block-sharded over the ranks (no data-path collective); inside a rank its block runs through the
    batched kernels.  Strong scaling: the total work is fixed as N grows."""
    from scikit_tt_b200 import TT
    import scikit_tt_b200.tensor_train as tt
    from scikit_tt_b200.solvers import evp, multi
    d, nsys, rank, repeats, solver = 20, args.c5_systems, args.c5_rank, 1, args.c5_solver
    lo, hi = multi.shard_bounds(nsys, env.world, env.rank)
    ks = workloads.c5_pressures(nsys)
    ops = []
    for k in ks[lo:hi]:
        t = TT(workloads.co_oxidation_cores(d, k)).ortho_left().ortho_right()      # examples/co_oxidation.py:100
        ops.append(tt.eye(t.row_dims) + t)
    guess = tt.ones([3] * d, [1] * d, ranks=rank).ortho_left().ortho_right()
    dev = env.dev
    out = {}

    def step():
        env.flush.zero_()
        out["res"] = evp.als_batch(ops, guess, repeats=repeats, conv_eps=0, solver=solver) if ops else []
    steps = max(2, min(args.steps, 5))
    l0 = dev.launches()
    ms, wall = env.timed(step, steps, 2)
    launches = env.sum_over_ranks(dev.launches() - l0) // (steps + 2)
    ms = max(ms, wall)                                       # host TT in and out every step: this leg is end to end
    hs = 2 * repeats * nsys * steps
    leg = {"workload": f"C5: {nsys} x evp.als(I + co_oxidation(20, k_ad_CO), ones guess of rank {rank}, repeats={repeats}, "
                       f"solver='{solver}'), CO pressures 10^(8+p), p = linspace(-4, 2, {nsys})",
           "value": hs / (ms * 1e-3), "unit": "half-sweeps/s", "ms_per_step": ms / steps, "steps": steps,
           "n_gpus": env.world, "scaling": "strong", "systems_per_gpu": (nsys + env.world - 1) // env.world,
           "launches_per_step": launches, "launches_per_system_half_sweep": launches / (2 * repeats * nsys),
           "parallelism": f"systems block-sharded over {env.world} GPU(s), batched inside each GPU (no data-path collective)",
           "e2e_note": "timed through the public API with host numpy cores in and out every step"}
    if env.rank == 0 and env.world == 1 and not args.no_cpu:
        v, kind = cpu_sample_c5(rank=rank, repeats=repeats, solver=solver)
        leg["cpu_baseline"] = {"value": v, "unit": "half-sweeps/s", "cores": 1, "kind": kind,
                               "sample": "2 of the 64 pressures, one BLAS thread (the reference's best setting at these sizes, "
                                         "BASELINE.md section 2); the sweep over pressures is embarrassingly parallel over host "
                                         "cores, so the per-core rate times the core count bounds the CPU aggregate",
                               "host_cores": len(os.sched_getaffinity(0))}
    return leg


def leg_c4(env, args):
    """BASELINE config 4: d=10, n=16, R=8 (SPD solve variant of SURVEY.md 8d), one sle.als sweep at solution rank
    --c4-rank.  N > 1: the micro-matvec is sharded over the output solution-rank index across the ranks (multi.py)."""
    from scikit_tt_b200 import TT
    import scikit_tt_b200.tensor_train as ttm
    from scikit_tt_b200.solvers import sle
    torch = env.torch
    dev = env.dev
    d, n, R, r = 10, 16, 8, args.c4_rank
    op = TT(workloads.c4_spd_cores(d, n, R)).pin_memory()
    rhs = TT(workloads.rank1_rhs(d, n))
    x0 = TT(workloads.random_guess(d, n, r, seed=1)).ortho_right()
    kw = {}
    if env.world > 1 and "group" in sle.als.__code__.co_varnames:
        kw["group"] = env.dist.group.WORLD
    out = {}

    def step():
        env.flush.zero_()
        out["x"] = sle.als(op, x0, rhs, repeats=1, **kw)
    steps = max(2, min(args.steps, 3))
    ms, wall = env.timed(step, steps, 1)
    ms = max(ms, wall)
    sharded = bool(kw)
    if sharded:                                              # the size policy of sle.als may keep small micro systems replicated
        from scikit_tt_b200.solvers import multi
        sharded = multi.sharded_stats.get("solves", 0) > 0
    leg = {"workload": f"C4: sle.als, random SPD TT operator d={d}, n={n}, R={R}, solution rank {r} "
                       f"({r * n * r} unknowns per micro system)",
           "value": 2 * steps * (1 if kw or env.world == 1 else env.world) / (ms * 1e-3), "unit": "half-sweeps/s",
           "ms_per_step": ms / steps, "steps": steps, "n_gpus": env.world,
           "scaling": "strong" if (kw or env.world == 1) else "weak",
           "parallelism": "single GPU" if env.world == 1 else
                          (f"micro-matvec sharded over the output solution-rank index across {env.world} GPUs" if sharded
                           else ("every rank sweeps the same system (micro systems below sle.SHARD_MIN_UNKNOWNS stay replicated)"
                                 if kw else "replicas"))}
    if env.rank == 0:
        bnorm = np.prod([np.linalg.norm(c) for c in rhs.cores])
        leg["residual"] = float(ttm.residual_error(op, out["x"], rhs) / bnorm)
    # stack update at this rank (generic DMMA contraction chain)
    L = torch.randn((r, R, r), dtype=torch.float64, device="cuda")
    x = torch.randn((r, n, r), dtype=torch.float64, device="cuda")
    A = torch.randn((R, n, n, R), dtype=torch.float64, device="cuda")
    for _ in range(3):
        dev.stack_left_op(L, x, A)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(20):
        dev.stack_left_op(L, x, A)
    s1.record()
    torch.cuda.synchronize()
    us = s0.elapsed_time(s1) / 20 * 1e3
    F = stack_flops(r, R, n, r, R)
    leg["stack_update"] = {"flops": F, "us": us, "achieved": F / (us * 1e-6) / 1e12,
                           "frac": F / (us * 1e-6) / 1e12 / FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s"}
    return leg


def run_ours(args, cfg):
    env = Env()
    legs = [s for s in args.legs.split(",") if s]
    c3 = leg_c3(env, args, cfg)
    extra = {}
    for name, fn in (("same", lambda: leg_same_config(env, args, cfg)), ("c1", lambda: leg_c1(env, args)),
                     ("c5", lambda: leg_c5(env, args)), ("c4", lambda: leg_c4(env, args))):
        if name not in legs:
            continue
        key = "same_config" if name == "same" else name
        try:
            extra[key] = fn()
        except Exception as exc:                                # a leg must never take the headline down with it
            if env.world > 1:
                raise
            extra[key] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    if env.rank == 0:
        ms, world = c3["ms"], env.world
        line = {"metric": "ALS half-sweeps/s (fp64)", "value": 2 * args.steps * world / (ms * 1e-3), "unit": "half-sweeps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg_public(cfg),
                "e2e": {"value": 2 * args.e2e_steps * world / (c3["ms_e2e"] * 1e-3), "unit": "half-sweeps/s",
                        "h2d_bytes_per_step": c3["h2d"], "d2h_bytes_per_step": c3["d2h"], "steps": args.e2e_steps},
                "gpu_launches": c3["launches"], "clocks": c3["clocks"], "roofline": c3["roofline"],
                "residual": c3["residual"], "krylov": c3["krylov"]}
        line.update(extra)
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            val, per_step, kind = cpu_sample(cfg["d"], cfg["n"], args.sample_rank)
            line["cpu_baseline"] = {
                "value": val, "unit": "half-sweeps/s", "cores": cores, "kind": kind,
                "sample": f"the reference algorithm on the same operator family at solution rank {args.sample_rank} (dense "
                          f"{args.sample_rank ** 2 * cfg['n']}^2 micro matrices, 1 step = {per_step:.1f} s; compare with "
                          f"`same_config`); rank {cfg['r']} is not runnable by the reference algorithm (512 GiB micro matrix)"}
        print(json.dumps(line), flush=True)
    if env.world > 1:
        from scikit_tt_b200.solvers import multi
        multi.close_peer_exchanges()                         # cached exchange buffers of the sharded C4 leg
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--solver", default="solve")
    ap.add_argument("--sample-rank", type=int, default=4, dest="sample_rank")
    ap.add_argument("--ref-max-steps", type=int, default=20, dest="ref_max_steps")
    ap.add_argument("--e2e-steps", type=int, default=3, dest="e2e_steps")
    ap.add_argument("--legs", default="same,c1,c5,c4", help="extra legs besides the C3 headline: same,c1,c5,c4 (comma separated)")
    ap.add_argument("--c5-systems", type=int, default=64, dest="c5_systems")
    ap.add_argument("--c5-rank", type=int, default=8, dest="c5_rank")
    ap.add_argument("--c5-solver", default="eigs", dest="c5_solver")
    ap.add_argument("--c4-rank", type=int, default=256, dest="c4_rank",
                    help="solution rank of the C4 leg (BASELINE names 128 and 256; sharding the micro-matvec over GPUs only pays "
                         "at 256: 0.15 ms of compute per matvec at 128 does not amortise the exchange, profiles/README.md)")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    args = ap.parse_args()
    cfg = {"d": args.d, "n": args.n, "r": args.rank, "gpus": args.gpus}
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
