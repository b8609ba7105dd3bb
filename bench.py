#!/usr/bin/env python
"""bench.py -- ALS half-sweeps/s (fp64) on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rank R]

Workload (config.workload = "C3"): synthetic discretised-Laplacian TT operator d=32, n=64, operator rank 3 (SURVEY.md
8d: SLIM layout, symmetric positive definite), right-hand side rank 1 (seed 0), initial guess with interior solution
rank 64 (seed 1, right-orthonormalised).  One *step* = one `sle.als(op, x0, rhs, repeats=1)` = 2 half-sweeps
(forward + backward) over the 32 cores: 64 interface-stack updates, 63 micro systems of 262 144 unknowns solved
matrix-free (CG to a TRUE relative residual of 1e-14, warm-started from the sweep's current iterate; the reference's dense
micro matrix would be 512 GiB), 62 QR/RQ factorisations of 4096 x 64 unfoldings.

`value`  : half-sweeps/s with operator, right-hand side and initial guess resident in HBM (CUDA events, max over ranks).
`e2e`    : the same call through the public API with host numpy TT cores in and out (H2D + D2H inside the timed region).
`roofline`: the dominant kernel group -- the three DMMA contractions of one matrix-free micro-matvec (= one stack
           update: F = 2 r^3 R (n+m) + 2 r^2 R^2 m n flops) -- timed with CUDA events on the same buffers right after
           the timed region, against the measured fp64 tensor-pipe peak (profiles/r01_fp64_peaks.txt).
`cpu_baseline` / `--impl reference`: the numpy/scipy restatement of the reference algorithm (oracle/, kind "port"; the
           GPU box has no /root/reference) on the host cores, on a bounded sample of the same operator family: the
           largest solution rank whose dense micro matrix the reference algorithm can factorise in bounded time.

N > 1: this single linear system does not shard ("replicas only"): every rank runs an independent replica,
no data-path collective; value = N x 2 x K / max-over-ranks time, scaling "weak".
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

FP64_TENSOR_PEAK_TFLOPS = 37.1      # measured on this pool's B200: profiles/r01_fp64_peaks.txt (DMMA m8n8k4, sustained)
NCU_MATVEC_DRAM_BYTES = 2345472 + 9412352   # profiles/r01_ncu_final_full.txt, one launch of each matvec kernel


# ------------------------------------------------------------------------------------------------ workload
from workloads import laplace_cores, workload_cores  # noqa: E402,F401  (the operator families live in workloads.py)


def stack_flops(r, R, n, r2, R2):
    return 2 * r * R * r * n * r2 + 2 * r * r2 * R * R2 * n * n + 2 * r2 * R2 * r * n * r2


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
def cpu_sample(d, n, sample_rank, steps=1, warmup=0):
    """The reference algorithm (oracle restatement, numpy/scipy, all BLAS threads) on the bounded sample."""
    from oracle import sle as osle, tt as ott
    op, rhs, x0 = workload_cores(d, n, sample_rank)
    x0 = ott.ortho_right(x0)
    for _ in range(warmup):
        osle.als(op, x0, rhs, repeats=1)
    t0 = time.perf_counter()
    for _ in range(steps):
        osle.als(op, x0, rhs, repeats=1)
    dt = time.perf_counter() - t0
    return 2 * steps / dt, dt / steps


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    val, per_step = cpu_sample(cfg["d"], cfg["n"], args.sample_rank, steps=args.steps, warmup=min(args.warmup, 1))
    sample = (f"same operator family (d={cfg['d']}, n={cfg['n']}, R=3) at solution rank {args.sample_rank}: dense micro "
              f"matrix {args.sample_rank ** 2 * cfg['n']}^2 + LU as the reference does; the named rank {cfg['r']} needs a "
              f"512 GiB micro matrix and cannot run")
    line = {"impl": "reference", "metric": "ALS half-sweeps/s (fp64)", "value": val, "unit": "half-sweeps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg_public(cfg),
            "cpu_baseline": {"value": val, "unit": "half-sweeps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "half-sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cfg_public(cfg):
    return {"workload": "C3: sle.als on the rank-3 Laplacian-type TT operator", "d": cfg["d"], "n": cfg["n"],
            "operator_rank": 3, "solution_rank": cfg["r"], "repeats_per_step": 1, "half_sweeps_per_step": 2,
            "micro_solver": "matrix-free CG (Chronopoulos-Gear form, reductions fused into the matvec, warm start from the "
                            "sweep's current iterate), true relative residual 1e-14 (dense micro matrix impossible at this size)",
            "l2": "256 MiB buffer written between steps (inside the timed region)",
            "parallelism": "replicas" if cfg["gpus"] > 1 else "single GPU"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # the version banner goes to stdout, next to the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from scikit_tt_b200 import TT
    from scikit_tt_b200.solvers import sle
    from scikit_tt_b200._device import get_device
    dev = get_device()

    d, n, r = cfg["d"], cfg["n"], cfg["r"]
    opc, rhsc, x0c = workload_cores(d, n, r)
    op, rhs = TT(opc).pin_memory(), TT(rhsc).pin_memory()     # inputs of the end-to-end leg lie in page-locked host memory
    x0 = TT(x0c).ortho_right()                                # GPU ortho path (TT.ortho_right); its cores come back page-locked
    st = sle._State(op, x0, rhs)
    x0_dev = list(st.x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()
        st.reset(x0_dev)
        sle._run_als(st, 1, args.solver)

    result = {}

    def step_e2e():
        flush.zero_()
        result["x"] = sle.als(op, x0, rhs, repeats=1, solver=args.solver)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dev.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = dev.launches() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API (host TT in, host TT out)
    for _ in range(2):                                        # untimed: page-locked result blocks enter torch's host cache
        step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e2.record()
    for _ in range(args.e2e_steps):
        step_e2e()
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t_wall) * 1e3)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])

    # roofline of the dominant kernel group: one matrix-free micro-matvec at the middle core
    i = d // 2
    L, A, R = st.Lop[i], st.A[i], st.Rop[i]
    v = st.x[i]
    F = stack_flops(L.shape[0], A.shape[0], A.shape[2], R.shape[0], A.shape[3])
    lop = dev.local_op(L, A, R, prepare=True)                # prepared exactly as the Krylov solvers prepare it
    nt = dev.tiled_len(lop)
    reps = 200
    if nt > 0:
        # the form the matvec takes inside the persistent CG kernel (the dominant kernel of the step): `reps` matvecs in
        # ONE cooperative launch of pcg_persistent_kernel, stage-1 tiles | grid barrier | stage-2+3 tiles | grid barrier
        vt = torch.randn(nt, dtype=torch.float64, device="cuda")
        vt.view(-1, 68)[:, 64:] = 0.0
        yt = torch.zeros(nt, dtype=torch.float64, device="cuda")
        mv = lambda: dev.local_matvec_tiled_repeat(lop, vt, yt, reps)
        per_call = reps
        kernel_name = ("pcg_persistent_kernel, matvec phases (stage-1 tiles | grid barrier | stage-2+3 tiles | grid barrier) = the "
                       "contraction chain of one interface-stack update (r=64, R=3, n=64); the kernel is 1 launch per micro solve")
    else:
        mv = lambda: dev.local_matvec(lop, v)
        per_call = 1
        kernel_name = "generic three-GEMM contraction chain"
    for _ in range(3):
        mv()
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls = 5 if per_call > 1 else 200
    m0.record()
    for _ in range(calls):
        mv()
    m1.record()
    torch.cuda.synchronize()
    mv_ms = m0.elapsed_time(m1) / (calls * per_call)
    achieved = F / (mv_ms * 1e-3) / 1e12
    Ah = opc[i]
    nnz_blocks = int(sum(bool(np.any(Ah[b, :, :, q])) for b in range(Ah.shape[0]) for q in range(Ah.shape[3])))
    rr_, R_, n_, r2_, R2_ = L.shape[0], A.shape[0], A.shape[2], R.shape[0], A.shape[3]
    F_exec = 2 * rr_ * R_ * rr_ * n_ * r2_ + 2 * rr_ * r2_ * nnz_blocks * n_ * n_ + 2 * r2_ * R2_ * rr_ * n_ * r2_

    # the interface-stack update itself (same flop count; generic strided-GEMM chain of stacks.cu)
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
    F_stack = stack_flops(Ls.shape[0], st.A[i - 1].shape[0], st.A[i - 1].shape[2], xs.shape[2], st.A[i - 1].shape[3])

    if rank == 0:
        half_sweeps = 2 * args.steps * world
        sol = result["x"]
        h2d = sum(c.nbytes for t in (op, x0, rhs) for c in t.cores)
        d2h = sum(c.nbytes for c in sol.cores)
        from scikit_tt_b200 import tensor_train as ttm          # || A x - b || / || b ||, core-wise QR evaluation
        res = float(ttm.residual_error(op, sol, rhs) / np.prod([np.linalg.norm(c) for c in rhs.cores]))
        line = {"metric": "ALS half-sweeps/s (fp64)", "value": half_sweeps / (ms * 1e-3), "unit": "half-sweeps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg_public(cfg),
                "e2e": {"value": 2 * args.e2e_steps * world / (ms_e2e * 1e-3), "unit": "half-sweeps/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                             "frac": achieved / FP64_TENSOR_PEAK_TFLOPS,
                             "traffic": NCU_MATVEC_DRAM_BYTES,
                             "traffic_note": "dram bytes per matvec of the two-kernel form (mv_stage1 2.35 MB + mv_stage23 9.41 MB, "
                                             "profiles/r01_ncu_final_full.txt, ncu flushes the caches per replay); inside the "
                                             "persistent kernel every operand is L2-resident: 3.1 MB of DRAM reads for 20 matvecs "
                                             "(profiles/r01_ncu_persistent.txt), 6.4 MB read + 1.2 MB written for one whole in-sweep "
                                             "solve of about 19 matvecs (profiles/r01_ncu_pcg_final.txt); algorithmic bytes 4.72 MB",
                             "kernel": kernel_name,
                             "flops_per_matvec": F, "us_per_matvec": mv_ms * 1e3,
                             "executed_flops_per_matvec": F_exec,
                             "executed_note": "F is the dense formula of SURVEY.md 8d; the kernel skips the zero (b, b') "
                                              f"blocks of the operator core ({nnz_blocks} of {A.shape[0] * A.shape[3]} non-zero "
                                              "here), so the tensor pipe executes executed_flops_per_matvec",
                             "achieved_executed": F_exec / (mv_ms * 1e-3) / 1e12,
                             "stack_update": {"kernel": "sktt_stack_left_op = image build + tiling + stack_persistent_kernel (stage-1 tiles | stage-2 tiles with the third contraction per tile | ordered reduction of the tile partials)",
                                              "flops": F_stack, "us": stack_ms * 1e3,
                                              "achieved": F_stack / (stack_ms * 1e-3) / 1e12,
                                              "frac": F_stack / (stack_ms * 1e-3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS},
                             "peak_source": "measured fp64 DMMA pipe peak, profiles/r01_fp64_peaks.txt "
                                            "(MEASURED_PEAKS.json has no fp64 entry)"},
                "residual": res}
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            val, per_step = cpu_sample(d, n, args.sample_rank)
            line["cpu_baseline"] = {
                "value": val, "unit": "half-sweeps/s", "cores": cores, "kind": "port",
                "sample": f"oracle (numpy/scipy restatement of sle.als) on the same operator family at solution rank "
                          f"{args.sample_rank} (dense {args.sample_rank ** 2 * n}^2 micro matrices, 1 step = {per_step:.1f} s); "
                          f"rank {r} is not runnable by the reference algorithm (512 GiB micro matrix)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--e2e-steps", type=int, default=3, dest="e2e_steps")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    args = ap.parse_args()
    cfg = {"d": args.d, "n": args.n, "r": args.rank, "gpus": args.gpus}
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
