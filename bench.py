#!/usr/bin/env python
"""bench.py -- the headline benchmark of the Krylov inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

Metric (BASELINE.json): CG iterations / second, fp64 CSR, 3-D 7-point Laplacian 256^3 (config C3),
with the SpMV HBM GB/s of the dominant kernel reported in `roofline`.  The matrix is the reference's
own benchmark matrix (BenchmarkStencil -dim 3 -nx 256 -ny 256 -nz 256), generated on the GPU by
the same StencilGenerator rules; b = 1, x0 = 0, exactly the BenchmarkStencil / Test06 set-up.

A STEP is one replay of a recorded trace of `iters_per_step` CG iterations (BenchmarkStencil's
`-pt`): `--steps 10` with the default 20 iterations per step is the reference protocol's 200 timed
iterations after warm-up traces.  `value` is the whole-job rate with everything resident in HBM;
`e2e` is the same rate through the host-buffer API (per step: H2D of the right-hand side from pinned
memory, a fresh solve of `iters_per_step` iterations, D2H of the solution and the residual history).

At N > 1 (launched under torchrun) the FIXED 256^3 problem is row-partitioned over the N GPUs --
strong scaling -- with the ghost-x halo exchange and the dot-product all-reduces over NVLink peer memory
(fused into the kernels that produce the data; LSK_COMM=nccl keeps them on NCCL).  `--persistent` runs the
whole CG step as one persistent kernel per trace instead of three leaf kernels per iteration.

`--impl reference` times the reference's CPU task variants (the oracle restatement; the reference
itself needs Legion and cannot be built here) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (dim_flag, (nx, ny, nz), description)
    "c3": (3, (256, 256, 256), "BenchmarkStencil -dim 3: 3-D 7-point Laplacian 256^3, fp64 CSR, CG"),
    "c2": (2, (8192, 8192, 1), "BenchmarkStencil -dim 2: 2-D 5-point Laplacian 8192^2, fp64 CSR, CG"),
    "c4": (4, (192, 192, 192), "BenchmarkStencil -dim 4: 3-D 27-point stencil 192^3, fp64 CSR, CG"),
    "c1": (2, (256, 256, 1), "2-D 5-point Laplacian 256^2 (the CPU-runnable parity case)"),
    "tiny": (3, (48, 48, 48), "3-D 7-point Laplacian 48^3 (smoke)"),
}
METRIC = "cg_iterations_per_second"
UNIT = "it/s"


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (pynvml; the recipe's nvidia-smi line, in-process)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and clock-event reasons of one GPU while it is under load, sampled by `nvidia-smi -lms` in a
    SEPARATE process (the profiling guide's recipe).  In-process NVML polling was measured to stall this process's own
    CUDA launches for milliseconds on some hosts, which desynchronises the ranks of a multi-GPU run."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device_index: int, period_ms: int = 25):
        import shutil
        import subprocess
        import tempfile

        self.proc, self.out = None, None
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        # physical index of the device this process uses (CUDA_VISIBLE_DEVICES may remap)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        ident = str(device_index)
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids):
                ident = ids[device_index]
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen([exe, "-i", ident, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def start(self):
        pass  # sampling began in the constructor, before the warm-up

    def stop(self, t0: float | None = None, t1: float | None = None):
        """t0, t1: wall-clock bounds (time.time()) of the timed region."""
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        rows = []
        for line in self.out.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, int(float(f[1])), int(float(f[2])), float(f[3]), [n for n, v in zip(self.NAMES, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        # the GPU is under the same load from the warm-up on: samples of the timed region if there are any, else
        # the ones under load around it (power well above idle)
        loaded = inside or [r for r in rows if r[3] > 250.0] or rows
        clocks = sorted(r[1] for r in loaded)
        reasons = sorted({n for r in loaded for n in r[4]})
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": rows[0][2], "reasons": reasons,
                "power_w_max": max(r[3] for r in loaded), "samples": len(loaded), "samples_in_timed_region": len(inside),
                "how": "nvidia-smi -lms in a separate process, from the warm-up to the end of the timed region"}


# ---------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: the reference's CPU task variants on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(dim_flag, shape, threads, its_warm, its_timed, steps=1):
    """CG on the oracle (restated CPU task bodies, one piece per host thread -- one Legion CPU processor
    per piece).  Returns (iterations/s, list of per-step seconds, description of the sample)."""
    from oracle import oracle as orc

    off, val = orc.benchmark_stencil(dim_flag)
    dims = shape[:3] if dim_flag >= 3 else shape[:dim_flag]
    m = orc.stencil_csr(dims, off, val)
    orc.set_threads(threads)
    pl = orc.Planner([m.n_rows], [threads])
    pl.fill(1, 1.0)
    pl.add_matrix(m)
    cg = orc.CGSolver(pl)
    for _ in range(its_warm):
        cg.step()
    per_step = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for _ in range(its_timed):
            cg.step()
        per_step.append(time.perf_counter() - t0)
    total = sum(per_step)
    sample = (f"{steps} x {its_timed} CG iterations of the FULL {'x'.join(map(str, dims))} system after {its_warm} warm-up "
              f"iterations, {threads} pieces on {threads} host threads (reference-equivalent linear-time CSR body)")
    return steps * its_timed / total, per_step, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # under torchrun only rank 0 runs the CPU arm
    dim_flag, shape, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    its = max(1, args.ref_iters_per_step)
    rate, per_step, sample = cpu_reference_run(dim_flag, shape, threads, its_warm=max(1, args.warmup),
                                               its_timed=its, steps=args.steps)
    ms = 1e3 * sum(per_step) / len(per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "iters_per_step": its, "spaces": 1, "pieces": threads,
                   "note": "reference CPU task variants restated in C (oracle/): the reference needs Legion and cannot be built here"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from legionsolvers_b200 import _abi
    from legionsolvers_b200 import solvers as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: legionsolvers_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dim_flag, shape, desc = WORKLOADS[args.workload]
    if args.shape:
        shape = tuple(int(v) for v in args.shape.split(","))
        desc += f" [grid overridden to {shape}]"
    ipt = args.iters_per_step
    # everything (our kernels, NCCL, the timing events) goes on ONE explicit non-default stream: torch's
    # default stream has handle 0, which the runtime would read as "create a private stream"
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    rt = S.Runtime(device=local_rank, rank=rank, nranks=world, stream=stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(S.Runtime.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        rt.comm_init(bytes(uid.cpu().numpy().tobytes()))

    # ---- problem set-up, all on the GPU: BenchmarkStencil's matrix, b = 1, x0 = 0 -------------------
    t_setup = time.perf_counter()
    st = S.benchmark_stencil(dim_flag, *shape)
    n = shape[0] * shape[1] * shape[2]
    pieces = world  # -vp = total GPUs (bench_all.py:206-208)
    mat = S.CSRMatrix.stencil(rt, st, pieces)
    sol = S.PartitionedVector(rt, "sol", n, pieces)
    rhs = S.PartitionedVector(rt, "rhs", n, pieces)
    sol.zero_fill()
    rhs.constant_fill(1.0)
    pl = S.SquarePlanner(rt)
    pl.add_sol_vector(sol)
    pl.add_rhs_vector(rhs)
    pl.add_row_partitioned_matrix(mat, 0, 0)
    if args.solver == "cg":
        cg = S.CGSolver(pl, fused=not args.unfused, persistent=True if args.persistent else None)
    elif args.solver == "bicgstab":
        cg = S.BiCGStabSolver(pl, fused=not args.unfused)
    else:
        cg = S.GMRESSolver(pl, 10, fused=not args.unfused)  # BenchmarkStencil hard-codes restart = 10
    rt.fence()
    setup_s = time.perf_counter() - t_setup
    nnz = mat.nnz
    own_lo, own_hi = sol.owned_range()
    n_local = own_hi - own_lo + 1
    nnz_local = mat.slab_k_hi - mat.slab_k_lo + 1

    TRACE = 51  # the reference's trace id (test/BenchmarkStencil.cpp:218)

    def step():
        rt.begin_trace(TRACE)
        for _ in range(ipt):
            cg.step()
        rt.end_trace(TRACE)

    # ---- warm-up, then EXACTLY K timed steps between barriers, CUDA events on the launching stream ---
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    rt.fence()
    # keep warming until the clocks have had ~0.4 s under load (short multi-GPU steps would otherwise be timed on
    # a GPU still ramping up from idle).  The NUMBER of extra steps is fixed by rank 0 and broadcast: the ranks
    # must launch the same steps (the collectives are part of them).
    t_w = time.perf_counter()
    step()
    rt.fence()
    one = max(time.perf_counter() - t_w, 1e-4)
    extra = torch.tensor([int(min(2000, max(1, 0.4 / one)))], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        step()
    rt.fence()
    launches0 = rt.kernel_launches
    cs0 = rt.comm_stats() if world > 1 else None
    ph0 = rt.cg_phase_stats() if getattr(cg, "persistent", False) else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    wall1 = time.time()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = rt.kernel_launches - launches0
    clocks = sampler.stop(wall0, wall1) if sampler is not None else None
    comm_us = None
    if world > 1:
        cs1 = rt.comm_stats()
        its_timed = args.steps * ipt
        mine = torch.tensor([(cs1["ar_ns"] - cs0["ar_ns"]) / 1e3 / its_timed, (cs1["halo_ns"] - cs0["halo_ns"]) / 1e3 / its_timed],
                            dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        comm_us = {"allreduce_us_per_iteration_by_rank": [round(float(v[0]), 2) for v in allv],
                   "halo_close_us_per_iteration_by_rank": [round(float(v[1]), 2) for v in allv]}
    ms_per_step = elapsed_ms / args.steps
    value = args.steps * ipt / (elapsed_ms * 1e-3)
    phase_us = None
    if ph0 is not None:
        ph1 = rt.cg_phase_stats()
        its = max(1, ph1["iterations"] - ph0["iterations"])
        phase_us = {k[:-3] + "_us_per_iteration": round((ph1[k] - ph0[k]) / 1e3 / its, 2) for k in ph1 if k.endswith("_ns")}
        if dist is not None:  # every rank's CTA-0 view, in the order of the keys above
            mine_ph = torch.tensor(list(phase_us.values()), dtype=torch.float64, device="cuda")
            all_ph = [torch.zeros_like(mine_ph) for _ in range(world)]
            dist.all_gather(all_ph, mine_ph)
            phase_us = {"keys": list(phase_us.keys()), "by_rank": [[round(float(x), 2) for x in v] for v in all_ph]}

    # ---- roofline of the dominant kernel: the fused CSR SpMV + p.Ap, timed alone on the same stream ----
    L = _abi.lib()
    e_ptr, c_ptr, rp_ptr = mat.device_fields()
    g_lo, g_hi = pl.ghost_bounds(0, pl.local_colors(0)[0])
    xg = torch.rand(g_hi - g_lo + 1, dtype=torch.float64, device="cuda")
    yv = torch.zeros(n_local, dtype=torch.float64, device="cuda")
    dslot = torch.zeros(1, dtype=torch.float64, device="cuda")
    x_shifted = xg.data_ptr() - g_lo * 8
    w_ptr = xg.data_ptr() + (own_lo - g_lo) * 8

    def spmv():
        _abi.check(L.lsk_csr_spmv_f64(rt.ctx, stream, n_local, nnz_local, e_ptr, c_ptr, rp_ptr, mat.slab_k_lo, x_shifted,
                                      yv.data_ptr(), w_ptr, dslot.data_ptr(), None, 0), "lsk_csr_spmv_f64")

    for _ in range(5):
        spmv()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    s0.record()
    for _ in range(reps):
        spmv()
    s1.record()
    torch.cuda.synchronize()
    spmv_ms = s0.elapsed_time(s1) / reps
    spmv_ms_by_rank = None
    if dist is not None:  # the same kernel on every rank's slab: tells a slow GPU from a slow collective
        mine_ms = torch.tensor([spmv_ms], dtype=torch.float64, device="cuda")
        all_ms = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(all_ms, mine_ms)
        spmv_ms_by_rank = [round(float(v[0]), 5) for v in all_ms]
    spmv_bytes = 16 * nnz_local + 32 * n_local  # 16/nnz + rowptr 16 + x 8 + y 8 per row
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tf = ROOT / "profiles" / "spmv_traffic.json"
    if tf.exists() and args.workload == "c3" and world == 1:
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
    # solver-iteration roofline: fused minimum 16 nnz + 104 N bytes per iteration
    iter_bytes = 16 * nnz_local + 104 * n_local
    iter_frac = (iter_bytes * value / 1e9) / peak

    if args.solver != "cg":
        # secondary configs: device-resident rate only (a GMRES "iteration" is one restart cycle, as in BenchmarkStencil)
        if rank == 0:
            per_it = {"cg": 16 * nnz_local + 104 * n_local, "bicgstab": 32 * nnz_local + 184 * n_local}.get(args.solver)
            print(json.dumps({
                "metric": f"{args.solver}_iterations_per_second", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "solver": args.solver, "unknowns": n, "nnz": nnz, "iters_per_step": ipt, "pieces": pieces,
                           "fused": not args.unfused,
                           "iteration_roofline_frac": (per_it * value / 1e9 / peak) if per_it else None},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None},
                "gpu_launches": int(launches), "clocks": clocks}), flush=True)
        barrier()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- end to end through the host-buffer API ------------------------------------------------------
    # Every step is a FRESH solve of `ipt` iterations: its right-hand side comes from pinned host memory, its solution
    # and residual history go back to the host.  The copies run on a second stream and are double-buffered against the
    # iterations -- the H2D of step k+1's right-hand side starts as soon as step k's reset has consumed the previous one,
    # the D2H of step k's solution (staged device-to-device) overlaps step k+1 -- the way a production caller would
    # drive independent solves.  All copies of all steps lie inside the timed region.
    b_host = torch.ones(n, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    x_stage = torch.zeros(n_local, dtype=torch.float64, device="cuda")
    e2e_steps = max(2, min(args.steps, 10))
    TRACE_E2E = 52
    copy_stream = torch.cuda.Stream()
    cs = copy_stream.cuda_stream
    b_glob, x_stage_glob = b_host.data_ptr(), x_stage.data_ptr() - 8 * own_lo
    x_host_own = x_host[own_lo:own_lo + n_local]

    def e2e_run(nsteps):
        ev_h2d, ev_reset, ev_solved, ev_d2h = (torch.cuda.Event() for _ in range(4))
        copy_stream.wait_stream(tstream)
        pl.vector_from_async(1, 0, b_glob, cs)       # H2D: right-hand side of the first step
        ev_h2d.record(copy_stream)
        ev_d2h.record(copy_stream)
        hist = None
        for k in range(nsteps):
            tstream.wait_event(ev_h2d)                # this step's right-hand side has landed
            pl.zero_fill(0)
            cg.reset()                                # P <- RHS, R <- RHS, rr0: the last readers of RHS
            ev_reset.record(tstream)
            if k + 1 < nsteps:                        # H2D of the NEXT step's right-hand side, under this step's iterations
                copy_stream.wait_event(ev_reset)
                pl.vector_from_async(1, 0, b_glob, cs)
                ev_h2d.record(copy_stream)
            rt.begin_trace(TRACE_E2E)
            for _ in range(ipt):
                cg.step()
            rt.end_trace(TRACE_E2E)
            tstream.wait_event(ev_d2h)                # the staging buffer's previous content is on the host
            pl.vector_to_async(0, 0, x_stage_glob, stream)   # solution -> staging buffer (device to device)
            ev_solved.record(tstream)
            copy_stream.wait_event(ev_solved)
            with torch.cuda.stream(copy_stream):      # D2H: this step's solution, under the next step's iterations
                x_host_own.copy_(x_stage, non_blocking=True)
            ev_d2h.record(copy_stream)
            hist = cg.residual_norm_squared           # D2H: this step's residual history (waits for its iterations)
        copy_stream.synchronize()
        return hist

    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    hist = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_steps * ipt / e2e_s
    rr_final = float(hist[-1])
    x_check = float(x_host_own.abs().max())  # the solution really arrived

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only), bounded sample ---------------------
    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, 64))
        its = 6 if n >= 1 << 24 else 20
        rate, _, sample = cpu_reference_run(dim_flag, shape, threads, its_warm=2, its_timed=its)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": desc, "unknowns": n, "nnz": nnz, "iters_per_step": ipt, "spaces": 1, "pieces": pieces,
                "solver": ("CGSolver persistent kernel (3 HBM passes / iteration, grid barriers instead of kernel boundaries, one launch per step)"
                           if getattr(cg, "persistent", False) else "CGSolver fused (3 HBM passes / iteration)" if not args.unfused
                           else "CGSolver unfused (reference call sequence)"),
                "trace": "CUDA graph replay of iters_per_step iterations", "rhs": "b = 1, x0 = 0 (BenchmarkStencil)",
                "l2": "working set per GPU exceeds the 126 MB L2 (matrix streamed once per iteration)",
                "halo_bytes_per_matvec_rank0": pl.halo_bytes_per_matvec, "setup_seconds": round(setup_s, 3),
                "collectives": "none (1 GPU)" if world == 1 else rt.collectives,
                "comm_error": rt.comm_error() if world > 1 else 0,
                "time_inside_collectives": comm_us,
                "persistent_kernel_phases": phase_us,
                "spmv_ms_per_launch_by_rank": spmv_ms_by_rank,
                "residual_norm_squared_last": rr_final,
                "iteration_roofline": {"bytes_per_iteration_per_gpu": iter_bytes, "frac_of_peak": iter_frac},
            },
            "roofline": {"bound": "hbm", "kernel": "csr_tma_kernel<1> (fused CSR SpMV + p.Ap)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms},
            "spmv_gbs": achieved,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local,
                    "d2h_bytes_per_step": 8 * n_local + 8 * (ipt + 1), "steps": e2e_steps,
                    "what": ("per step: H2D rhs (pinned) -> reset -> iters_per_step CG iterations -> D2H solution + residual history; "
                             "copies on a second stream, double-buffered against the iterations of the neighbouring steps"),
                    "solution_abs_max": x_check},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c3")
    ap.add_argument("--iters-per-step", type=int, default=20, help="CG iterations per recorded trace (BenchmarkStencil -pt)")
    ap.add_argument("--ref-iters-per-step", type=int, default=2, help="CPU reference arm: iterations per step (bounded sample)")
    ap.add_argument("--solver", choices=["cg", "bicgstab", "gmres"], default="cg",
                    help="cg is the headline; bicgstab / gmres (restart 10, as BenchmarkStencil) are reported for the other configs")
    ap.add_argument("--shape", type=str, default=None, help="developer override nx,ny,nz of the workload's grid")
    ap.add_argument("--unfused", action="store_true", help="run the reference's unfused call sequence on the GPU")
    ap.add_argument("--persistent", action="store_true",
                    help="CG as one persistent kernel per step (default: three leaf kernels per iteration, ~2 %% faster; also LSK_CG_PERSISTENT=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
