#!/usr/bin/env python
"""bench.py -- the benchmark of the Krylov inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cusparse] [--workload c3] [--solver cg]

Headline (BASELINE.json): CG iterations / second, fp64 CSR, 3-D 7-point Laplacian 256^3 (config C3), with the SpMV
HBM GB/s of the dominant kernel reported in `roofline`.  The matrix is the reference's own benchmark matrix
(BenchmarkStencil -dim 3 -nx 256 -ny 256 -nz 256), generated on the GPU by the same StencilGenerator rules; b = 1,
x0 = 0, exactly the BenchmarkStencil / Test06 set-up.

Workloads (`--workload`, BASELINE.json `configs`; each prints ONE JSON line of the same shape):
    c3  3-D 7-point 256^3, CG (default; strong-scaled over --gpus)       c2  2-D 5-point 8192^2, CG
    c4  3-D 27-point 192^3, BiCGStab                                      c5  COO power-law ~1e8 nnz, GMRES(30)
    c1  2-D 5-point 256^2 (the CPU-runnable parity case)                  tiny  48^3 (smoke)
`--solver` overrides the workload's solver; `--spaces 2` reproduces BenchmarkStencil's doubled block-diagonal system
(test/BenchmarkStencil.cpp:199-207).

A STEP is one replay of a recorded trace of `iters_per_step` solver iterations (BenchmarkStencil's `-pt`; a GMRES
"iteration" is one restart cycle, as there): `--steps 10` with the default 20 is the reference protocol's 200 timed
iterations after warm-up traces.  `value` is the whole-job rate with everything resident in HBM; `e2e` is the same
rate through the host-buffer API (per step: H2D of the right-hand side from pinned memory, a fresh solve of
`iters_per_step` iterations, D2H of the solution and the solver's history), fully pipelined -- nothing is read
back before the last step has been enqueued.  `parity` compares a fresh solve on the GPU(s) with the CPU oracle on
the SAME full-size system (history and a strided sample of the solution), at every N.

At N > 1 (launched under torchrun) the FIXED problem is row-partitioned over the N GPUs -- strong scaling -- with the
ghost-x halo exchange and the dot-product all-reduces over NVLink peer memory (fused into the kernels that produce
the data; LSK_COMM=nccl keeps them on NCCL).

`--impl reference` times the reference's CPU task variants (the oracle restatement; the reference itself needs
Legion and cannot be built here) on the host cores, on a bounded sample.  `--impl cusparse` times the reference's GPU
call sequence through cuSPARSE / cuBLAS (baseline/cusparse_ref.cu), the library path the hand-written kernels replace.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (kind, dim_flag, (nx, ny, nz), default solver, description)
    "c3": ("stencil", 3, (256, 256, 256), "cg", "BenchmarkStencil -dim 3: 3-D 7-point Laplacian 256^3, fp64 CSR"),
    "c2": ("stencil", 2, (8192, 8192, 1), "cg", "BenchmarkStencil -dim 2: 2-D 5-point Laplacian 8192^2, fp64 CSR"),
    "c4": ("stencil", 4, (192, 192, 192), "bicgstab", "BenchmarkStencil -dim 4: 3-D 27-point stencil 192^3, fp64 CSR"),
    "c5": ("coo", 0, (22, 0, 0), "gmres", "synthetic COO, power-law row lengths, N = 2^22, ~1e8 nnz, fp64"),
    "c1": ("stencil", 2, (256, 256, 1), "cg", "2-D 5-point Laplacian 256^2 (the CPU-runnable parity case), fp64 CSR"),
    "tiny": ("stencil", 3, (48, 48, 48), "cg", "3-D 7-point Laplacian 48^3 (smoke), fp64 CSR"),
}
SOLVER_NAMES = {"cg": "CG", "bicgstab": "BiCGStab", "gmres": "GMRES(restart)"}
UNIT = "it/s"
GMRES_RESTART = {"c5": 30}  # BenchmarkStencil hard-codes 10; config C5 names GMRES(30)


def metric_name(solver: str) -> str:
    return f"{solver}_iterations_per_second"


def workload_desc(args) -> str:
    kind, dim_flag, shape, default_solver, desc = WORKLOADS[args.workload]
    solver = args.solver or default_solver
    name = SOLVER_NAMES[solver].replace("restart", str(gmres_restart(args)))
    return f"{desc}, {name}"


def gmres_restart(args) -> int:
    return GMRES_RESTART.get(args.workload, 10)


def shared_config(args, n, nnz) -> dict:
    """The keys both arms print identically (the driver compares them)."""
    kind, dim_flag, shape, default_solver, desc = WORKLOADS[args.workload]
    return {"workload": workload_desc(args), "unknowns": int(n) * args.spaces, "nnz": int(nnz) * args.spaces,
            "iters_per_step": args.iters_per_step, "spaces": args.spaces, "pieces": args.gpus,
            "solver": args.solver or default_solver}


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (the recipe's nvidia-smi line, in a separate process)
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
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")  # physical index of the device this process uses
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

    def stop(self, t0: float | None = None, t1: float | None = None):
        """t0, t1: wall-clock bounds (time.time()) of the timed region."""
        import datetime

        empty = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return empty
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
            return empty
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
# the workload on the host: oracle matrix (CPU arm, parity) -- and, for C5, the arrays uploaded to the GPU
# ---------------------------------------------------------------------------------------------------
_host_matrix_cache = {}


def host_matrix(args):
    """The workload's matrix in the oracle's representation (global arrays on the host)."""
    from oracle import oracle as orc

    kind, dim_flag, shape, _, _ = WORKLOADS[args.workload]
    if args.shape:
        shape = tuple(int(v) for v in args.shape.split(","))
    key = (args.workload, shape)
    if key not in _host_matrix_cache:
        if kind == "stencil":
            off, val = orc.benchmark_stencil(dim_flag)
            dims = shape[:3] if dim_flag >= 3 else shape[:dim_flag]
            _host_matrix_cache[key] = orc.stencil_csr(dims, off, val)
        else:
            from legionsolvers_b200.workloads import power_law_coo

            n, entry, row, col = power_law_coo(shape[0])
            _host_matrix_cache[key] = orc.Matrix(n, n, entry, col, row=row)
    return _host_matrix_cache[key]


def pin_cpu_threads():
    """One oracle thread per core, pinned: must be set before libgomp starts."""
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")


def cpu_reference_run(args, threads, its_warm, its_timed, steps=1, sample_stride=None):
    """The solver on the oracle (restated CPU task bodies, one piece per host thread -- one Legion CPU processor per
    piece).  Returns a dict: rate (iterations/s), per-step seconds, description of the sample, the history after
    its_warm + steps * its_timed iterations and a strided sample of the solution of space 0."""
    pin_cpu_threads()
    from oracle import oracle as orc

    _, _, _, default_solver, _ = WORKLOADS[args.workload]
    solver = args.solver or default_solver
    m = host_matrix(args)
    orc.set_threads(threads)
    pl = orc.Planner([m.n_rows] * args.spaces, [threads] * args.spaces)
    pl.fill(1, 1.0)
    for s in range(args.spaces):
        pl.add_matrix(m, s, s)
    if solver == "cg":
        sv = orc.CGSolver(pl)
    elif solver == "bicgstab":
        sv = orc.BiCGStabSolver(pl)
    else:
        sv = orc.GMRESSolver(pl, gmres_restart(args))
    for _ in range(its_warm):
        sv.step()
    per_step = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for _ in range(its_timed):
            sv.step()
        per_step.append(time.perf_counter() - t0)
    total = sum(per_step)
    if solver == "cg":
        hist = {"residual_norm_squared": sv.residual_norm_squared.tolist()}
    elif solver == "bicgstab":
        hist = {"rho": sv.rho.tolist(), "alpha": sv.alpha.tolist(), "omega": sv.omega.tolist()}
    else:
        hist = {"hessenberg": sv.inner_products.tolist()}
    x = pl.vector(0, 0)
    stride = sample_stride or max(1, m.n_rows // 4096)
    sample = (f"{steps} x {its_timed} {SOLVER_NAMES[solver].replace('restart', str(gmres_restart(args)))} iterations of the FULL system "
              f"({m.n_rows} rows x {args.spaces} space(s), {m.nnz} nnz) after {its_warm} warm-up iterations, {threads} pieces on {threads} pinned host "
              f"threads (reference-equivalent CPU task bodies; the CSR body is the linear-time form proven bit-identical to the reference's scan)")
    return {"rate": steps * its_timed / total if total > 0 else 0.0, "per_step": per_step, "sample": sample, "history": hist,
            "iterations": its_warm + steps * its_timed, "x_sample": x[::stride].copy(), "x_stride": stride, "x_absmax": float(abs(x).max()),
            "threads": threads}


def cpu_sample_sizes(args):
    """(warm-up iterations, timed iterations) of the bounded CPU sample, sized for ~10-30 s of host work."""
    _, _, _, default_solver, _ = WORKLOADS[args.workload]
    solver = args.solver or default_solver
    if args.workload in ("c1", "tiny"):
        return (2, 20) if solver != "gmres" else (0, 2)
    if solver == "gmres":
        return 0, 1
    if solver == "bicgstab":
        return 1, 4
    return 2, 6


def reference_as_written_c1(threads=4):
    """the reference's CSR body AS WRITTEN -- for every non-zero a linear scan of the piece's
    rowptr (src/CSRMatrixTasks.cpp:73-91) -- is only runnable at C1 size: CG on the 256^2 system, 4 pieces."""
    pin_cpu_threads()
    from oracle import oracle as orc

    off, val = orc.benchmark_stencil(2)
    m = orc.stencil_csr((256, 256), off, val)
    orc.set_threads(threads)
    pl = orc.Planner([m.n_rows], [threads])
    pl.fill(1, 1.0)
    pl.add_matrix(m)
    pl.use_literal_csr(True)
    cg = orc.CGSolver(pl)
    cg.step()
    t0 = time.perf_counter()
    its = 2
    for _ in range(its):
        cg.step()
    return {"value": its / (time.perf_counter() - t0), "unit": UNIT, "what": "CG on 2-D 5-pt 256^2, 4 pieces, the quadratic rowptr scan as written",
            "cores": threads}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # under torchrun only rank 0 runs the CPU arm
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    its = max(1, args.ref_iters_per_step)
    res = cpu_reference_run(args, threads, its_warm=max(1, min(args.warmup, 2)), its_timed=its, steps=args.steps)
    m = host_matrix(args)
    # a STEP of the CPU arm is a bounded sample (`its` iterations); ms_per_step is scaled to iters_per_step iterations
    ms = 1e3 * (sum(res["per_step"]) / len(res["per_step"])) * (args.iters_per_step / its)
    _, _, _, default_solver, _ = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": metric_name(args.solver or default_solver), "value": res["rate"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": shared_config(args, m.n_rows, m.nnz),
        "cpu_baseline": {"value": res["rate"], "unit": UNIT, "cores": threads, "kind": "port", "sample": res["sample"],
                         "iterations_per_timed_step": its,
                         "note": "reference CPU task variants restated in C (oracle/): the reference needs Legion and cannot be built here"},
        "e2e": {"value": res["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# parity of a GPU history / solution against the oracle's
# ---------------------------------------------------------------------------------------------------
def parity_report(solver, gpu_hist, cpu, x_gpu_sample, restart):
    """gpu_hist / cpu['history']: dicts of lists.  Tolerances: CG residual history 1e-10 relative over the window where
    |r|^2 >= 1e-12 |b|^2 (north_star); BiCGStab rho/alpha/omega over the first 6 steps and GMRES Hessenberg over the first 10
    columns 1e-9 (Lanczos-type recurrences amplify the dot-order difference ~3x per step; tests/ derive the windows)."""
    import numpy as np

    out = {"solver": solver}
    if solver == "cg":
        g, c = np.array(gpu_hist["residual_norm_squared"]), np.array(cpu["history"]["residual_norm_squared"])
        k = min(g.size, c.size)
        g, c = g[:k], c[:k]
        err = float(np.max(np.abs(g - c) / np.maximum(np.abs(c), 1e-12 * abs(c[0]))))
        out.update({"hist_rel_err": err, "tol": 1e-10, "entries": int(k)})
    elif solver == "bicgstab":
        err = 0.0
        k = 0
        for name in ("rho", "alpha", "omega"):
            g, c = np.array(gpu_hist[name]), np.array(cpu["history"][name])
            k = min(g.size, c.size, 7)
            err = max(err, float(np.max(np.abs(g[1:k] - c[1:k]) / np.abs(c[1:k]))) if k > 1 else 0.0)
        out.update({"hist_rel_err": err, "tol": 1e-9, "entries": int(k), "what": "rho, alpha, omega of the first steps"})
    else:
        g, c = np.array(gpu_hist["hessenberg"]), np.array(cpu["history"]["hessenberg"])
        cols = min(10, restart)
        err = float(np.max(np.abs(g[:, :cols] - c[:, :cols])) / np.max(np.abs(c)))
        out.update({"hist_rel_err": err, "tol": 1e-9, "entries": int(cols), "what": "Hessenberg (inner_products), first columns of the last cycle"})
    xs_c = cpu["x_sample"]
    xerr = float(np.max(np.abs(x_gpu_sample - xs_c)) / max(cpu["x_absmax"], 1e-300)) if xs_c.size == x_gpu_sample.size else float("nan")
    out.update({"x_rel_err": xerr, "x_tol": 1e-10 if solver == "cg" else 1e-8, "x_samples": int(xs_c.size),
                "iterations": cpu["iterations"], "oracle_threads": cpu["threads"]})
    out["ok"] = bool(out["hist_rel_err"] <= out["tol"] and xerr <= out["x_tol"])
    return out


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
        import datetime

        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(minutes=20))

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

    kind, dim_flag, shape, default_solver, _ = WORKLOADS[args.workload]
    solver = args.solver or default_solver
    if args.shape:
        shape = tuple(int(v) for v in args.shape.split(","))
    ipt = args.iters_per_step
    restart = gmres_restart(args)
    spaces = args.spaces
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

    # ---- problem set-up: BenchmarkStencil's matrix generated on the GPU (or the C5 arrays uploaded), b = 1, x0 = 0 ----
    t_setup = time.perf_counter()
    pieces = world  # -vp = total GPUs (bench_all.py:206-208)
    if kind == "stencil":
        st = S.benchmark_stencil(dim_flag, *shape)
        n = shape[0] * shape[1] * shape[2]
        mat = S.CSRMatrix.stencil(rt, st, pieces)
    else:
        hm = host_matrix(args)
        n = hm.n_rows
        lo = (n * rank) // world
        hi = (n * (rank + 1)) // world - 1
        k_lo, k_hi = int(np.searchsorted(hm.row, lo, side="left")), int(np.searchsorted(hm.row, hi, side="right")) - 1
        mat = S.COOMatrix.from_host(rt, n, n, hm.entry, hm.row, hm.col, k_range=(k_lo, k_hi), nnz_global=hm.nnz)
    pl = S.SquarePlanner(rt)
    sols, rhss = [], []
    for s in range(spaces):
        v = S.PartitionedVector(rt, f"sol{s}", n, pieces)
        v.zero_fill()
        pl.add_sol_vector(v)
        sols.append(v)
    for s in range(spaces):
        v = S.PartitionedVector(rt, f"rhs{s}", n, pieces)
        v.constant_fill(1.0)
        pl.add_rhs_vector(v)
        rhss.append(v)
    for s in range(spaces):
        pl.add_row_partitioned_matrix(mat, s, s)
    if solver == "cg":
        sv = S.CGSolver(pl, fused=not args.unfused)
    elif solver == "bicgstab":
        sv = S.BiCGStabSolver(pl, fused=not args.unfused)
    else:
        sv = S.GMRESSolver(pl, restart, fused=not args.unfused)
    rt.fence()
    setup_s = time.perf_counter() - t_setup
    nnz = mat.nnz
    own_lo, own_hi = sols[0].owned_range()
    n_local = own_hi - own_lo + 1
    nnz_local = mat.slab_k_hi - mat.slab_k_lo + 1
    is_csr = bool(mat.is_csr)

    TRACE = 51  # the reference's trace id (test/BenchmarkStencil.cpp:218)

    def step():
        rt.begin_trace(TRACE)
        for _ in range(ipt):
            sv.step()
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
                   "halo_poll_us_per_iteration_by_rank": [round(float(v[1]), 2) for v in allv]}
    if os.environ.get("LSK_TAIL_STATS") == "1":  # developer switch: phase accounting of the one-launch CG tail (cumulative, warm-up included)
        import ctypes as C

        ts = (C.c_uint64 * 6)()
        _abi.check(_abi.lib().lsk_cg_tail_stats(rt.ctx, stream, ts), "lsk_cg_tail_stats")
        if ts[5]:
            print(f"tail-stats rank {rank}: " + ", ".join(f"{nm} {ts[i] / 1e3 / ts[5]:.2f} us" for i, nm in enumerate(("resolve", "phase1", "rr wait", "phase2", "unpack")))
                  + f" per launch ({ts[5]} launches)", file=sys.stderr)
    ms_per_step = elapsed_ms / args.steps
    value = args.steps * ipt / (elapsed_ms * 1e-3)
    # ---- roofline of the dominant kernel: the (fused) mat-vec, timed alone on the same stream -----------------
    L = _abi.lib()
    e_ptr, c_ptr, third_ptr = mat.device_fields()
    g_lo, g_hi = pl.ghost_bounds(0, pl.local_colors(0)[0])
    xg = torch.rand(g_hi - g_lo + 1, dtype=torch.float64, device="cuda")
    yv = torch.zeros(n_local, dtype=torch.float64, device="cuda")
    dslot = torch.zeros(2, dtype=torch.float64, device="cuda")
    x_shifted = xg.data_ptr() - g_lo * 8
    w_ptr = xg.data_ptr() + (own_lo - g_lo) * 8
    if is_csr:
        kernel_name = ("csr_ws_kernel<NDOT=1> (warp-specialised TMA-pipelined CSR SpMV fused with y.w)")
        spmv_bytes = 16 * nnz_local + 32 * n_local  # 16/nnz + rowptr 16 + x 8 + y 8 per row

        def spmv():
            _abi.check(L.lsk_csr_spmv_f64(rt.ctx, stream, n_local, nnz_local, e_ptr, c_ptr, third_ptr, mat.slab_k_lo, x_shifted,
                                          yv.data_ptr(), w_ptr, dslot.data_ptr(), None, 0), "lsk_csr_spmv_f64")
    else:
        kernel_name = "coo_segreduce_kernel (segmented warp-shuffle COO SpMV, beta = 1)"
        spmv_bytes = 24 * nnz_local + 8 * (g_hi - g_lo + 1) + 16 * n_local  # 24/nnz + x 8 per column touched + y read-modify-write 16 per row
        y_shifted = yv.data_ptr() - own_lo * 8

        def spmv():
            _abi.check(L.lsk_coo_spmv_f64(rt.ctx, stream, nnz_local, e_ptr, third_ptr, c_ptr, x_shifted, y_shifted, own_lo, own_hi, g_lo, g_hi),
                       "lsk_coo_spmv_f64")

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
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # DRAM traffic of that kernel: from the committed ncu capture of the SAME kernel source, else unknown
    traffic, traffic_src = None, "none: no ncu capture of this workload / kernel source committed"
    tf = ROOT / "profiles" / "spmv_traffic.json"
    if tf.exists() and world == 1:
        rec = json.loads(tf.read_text()).get(args.workload)
        if rec:
            sha = hashlib.sha256(b"".join((ROOT / "legionsolvers_b200" / "csrc" / f).read_bytes() for f in rec.get("sources", []))).hexdigest()
            if sha == rec.get("sources_sha256"):
                traffic, traffic_src = rec["dram_bytes_per_launch"], f"{rec.get('from', 'profiles/')} (ncu --set full of this kernel source, not this run)"
            else:
                traffic_src = "stale: the kernel source changed since the committed ncu capture"
    # solver-iteration roofline: fewest-pass bytes per iteration
    mv = spmv_bytes + (8 * n_local if not is_csr else 0)  # the planner zero-fills before an accumulating (COO) mat-vec
    if solver == "cg":
        iter_bytes = mv + 72 * n_local
    elif solver == "bicgstab":
        iter_bytes = 2 * mv + 120 * n_local
    else:  # one restart cycle: residual (mat-vec, xpay, dot, scal), m Arnoldi steps (mat-vec + (j + 1) fused axpy.dot + scal), m update axpys
        m_ = restart
        iter_bytes = (mv + 24 * n_local + 16 * n_local + 16 * n_local) + sum(mv + 32 * n_local * (j + 1) + 16 * n_local for j in range(m_)) + 24 * n_local * m_
    iter_bytes *= spaces
    iter_frac = (iter_bytes * value / 1e9) / peak

    # ---- end to end through the host-buffer API, fully pipelined ---------------------------------------------------
    # Every step is a FRESH solve of `ipt` iterations: its right-hand side comes from pinned host memory, its solution and
    # history go back to the host.  Copies run on a second stream, double-buffered against the iterations (the H2D of step
    # k + 1's right-hand side starts as soon as step k's reset has consumed the previous one; the D2H of step k's solution,
    # staged device-to-device, overlaps step k + 1).  reset() and the iterations are recorded traces (CUDA graphs); the host
    # enqueues all steps without reading anything back and synchronises once at the end.
    e2e = None
    if not args.no_e2e:
        b_host = [torch.ones(n, dtype=torch.float64).pin_memory() for _ in range(spaces)]
        x_host = [torch.zeros(n, dtype=torch.float64).pin_memory() for _ in range(spaces)]
        x_stage = [torch.zeros(n_local, dtype=torch.float64, device="cuda") for _ in range(spaces)]
        e2e_steps = max(2, min(3 * args.steps, 30))  # enough steps that the pipeline's fill and drain (first H2D, last D2H) are amortised
        nh = {"cg": 1, "bicgstab": 3, "gmres": 0}[solver]
        hist_len = ipt + 1
        hist_stage = torch.zeros(max(1, nh) * hist_len, dtype=torch.float64, device="cuda")
        hist_host = torch.zeros((e2e_steps + 2, max(1, nh) * hist_len), dtype=torch.float64).pin_memory()
        TRACE_RESET, TRACE_ITERS = 52, 53
        # two copy streams: H2D and D2H run concurrently (PCIe is full duplex); on ONE stream the H2D of step k + 2 queues
        # behind the D2H of step k and the step period becomes the SUM of the two copies (measured at 8 GPUs, where eight
        # ranks share the host's PCIe fabric: e2e 0.48 of the resident rate)
        copy_stream = torch.cuda.Stream()
        d2h_stream = torch.cuda.Stream()
        cs = copy_stream.cuda_stream

        e2e_trace = os.environ.get("LSK_E2E_TRACE") == "1"  # developer switch: device-side timeline of the e2e loop
        e2e_skip = os.environ.get("LSK_E2E_SKIP", "")       # developer switch: "h2d" / "d2h" leaves that copy out (diagnosis only)
        marks = []

        def mark(stream_, what, k):
            if e2e_trace:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream_)
                marks.append((what, k, ev, time.perf_counter()))

        def e2e_run(nsteps):
            ev_h2d, ev_reset, ev_solved, ev_d2h = (torch.cuda.Event() for _ in range(4))
            copy_stream.wait_stream(tstream)
            d2h_stream.wait_stream(tstream)
            for s in range(spaces):
                pl.vector_from_async(1, s, b_host[s].data_ptr(), cs)       # H2D: right-hand side of the first step
            ev_h2d.record(copy_stream)
            ev_d2h.record(d2h_stream)
            for k in range(nsteps):
                tstream.wait_event(ev_h2d)                # this step's right-hand side has landed
                mark(tstream, "reset_begin", k)
                rt.begin_trace(TRACE_RESET)
                pl.zero_fill(0)
                sv.reset()                                # the last readers of RHS
                rt.end_trace(TRACE_RESET)
                ev_reset.record(tstream)
                mark(tstream, "iters_begin", k)
                if k + 1 < nsteps and "h2d" not in e2e_skip:  # H2D of the NEXT step's right-hand side, under this step's iterations
                    copy_stream.wait_event(ev_reset)
                    mark(copy_stream, "h2d_begin", k + 1)
                    for s in range(spaces):
                        pl.vector_from_async(1, s, b_host[s].data_ptr(), cs)
                    mark(copy_stream, "h2d_end", k + 1)
                    ev_h2d.record(copy_stream)
                rt.begin_trace(TRACE_ITERS)
                for _ in range(ipt):
                    sv.step()
                rt.end_trace(TRACE_ITERS)
                mark(tstream, "iters_end", k)
                tstream.wait_event(ev_d2h)                # the staging buffers' previous content is on the host
                for s in range(spaces):                   # solution + history -> staging buffers (device to device)
                    pl.vector_to_async(0, s, x_stage[s].data_ptr() - 8 * own_lo, stream)
                for h in range(nh):
                    sv.history_copy_async(h, hist_stage.data_ptr() + 8 * h * hist_len, hist_len, stream)
                ev_solved.record(tstream)
                d2h_stream.wait_event(ev_solved)
                mark(d2h_stream, "d2h_begin", k)
                with torch.cuda.stream(d2h_stream):       # D2H under the next step's iterations
                    for s in range(spaces if "d2h" not in e2e_skip else 0):
                        x_host[s][own_lo:own_lo + n_local].copy_(x_stage[s], non_blocking=True)
                    hist_host[k].copy_(hist_stage, non_blocking=True)
                mark(d2h_stream, "d2h_end", k)
                ev_d2h.record(d2h_stream)
            copy_stream.synchronize()
            d2h_stream.synchronize()
            torch.cuda.synchronize()

        e2e_run(2)
        barrier()
        marks.clear()
        cs_e0 = rt.comm_stats() if world > 1 else None
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        if cs_e0 is not None and e2e_trace:
            cs_e1 = rt.comm_stats()
            print(f"e2e-trace rank {rank}: inside all-reduces {(cs_e1['ar_ns'] - cs_e0['ar_ns']) / 1e3 / (e2e_steps * ipt):.2f} us / iteration, "
                  f"halo {(cs_e1['halo_ns'] - cs_e0['halo_ns']) / 1e3 / (e2e_steps * ipt):.2f} us / iteration (reset included)", file=sys.stderr)
        if e2e_trace and rank == 0 and marks:
            base_ev, base_t = marks[0][2], marks[0][3]
            for what, k, ev, th in marks:
                print(f"e2e-trace step {k:2d} {what:12s} device {base_ev.elapsed_time(ev):9.3f} ms   host-enqueue {1e3 * (th - base_t):9.3f} ms", file=sys.stderr)
        e2e_value = e2e_steps * ipt / e2e_s
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local * spaces,
               "d2h_bytes_per_step": 8 * n_local * spaces + 8 * nh * hist_len, "steps": e2e_steps,
               "what": ("per step: H2D rhs (pinned) -> zero_fill + reset (trace) -> iters_per_step iterations (trace) -> D2H solution + history; copies on a "
                        "H2D and a D2H stream, double-buffered; nothing is read back until every step has been enqueued"),
               "ratio_to_resident": e2e_value / value,
               "solution_abs_max": float(x_host[0][own_lo:own_lo + n_local].abs().max()),
               "history_last": float(hist_host[e2e_steps - 1][hist_len - 1]) if nh else None}

    # ---- parity at FULL size against the oracle, at every N; CPU baseline from the same oracle run at N = 1 ---------
    # GPU side: a fresh solve of K iterations (K = the oracle's warm-up + timed iterations), history + strided solution sample
    cpu_baseline, parity, as_written = None, None, None
    if not args.no_parity:
        its_warm, its_timed = cpu_sample_sizes(args)
        K = its_warm + its_timed
        for s in range(spaces):
            rhss[s].constant_fill(1.0)
        pl.zero_fill(0)
        sv.reset()
        for _ in range(K):
            sv.step()
        rt.fence()
        if solver == "cg":
            gpu_hist = {"residual_norm_squared": sv.residual_norm_squared.tolist()}
        elif solver == "bicgstab":
            gpu_hist = {"rho": sv.rho.tolist(), "alpha": sv.alpha.tolist(), "omega": sv.omega.tolist()}
        else:
            gpu_hist = {"hessenberg": sv.inner_products.tolist()}
        stride = max(1, n // 4096)
        xs_local = torch.zeros(n, dtype=torch.float64).pin_memory()
        pl.vector_to_async(0, 0, xs_local.data_ptr(), stream)
        rt.fence()
        first = ((own_lo + stride - 1) // stride) * stride
        mine_idx = np.arange(first, own_hi + 1, stride)
        mine_val = xs_local.numpy()[mine_idx]
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, (mine_idx, mine_val))
            x_gpu = np.concatenate([v for _, v in gathered])
        else:
            x_gpu = mine_val
        if rank == 0:
            cores = os.cpu_count() or 1
            threads = max(1, min(cores, 64))
            cpu = cpu_reference_run(args, threads, its_warm=its_warm, its_timed=its_timed, sample_stride=stride)
            parity = parity_report(solver, gpu_hist, cpu, x_gpu, restart)
            if world == 1 and not args.no_cpu_baseline:
                cpu_baseline = {"value": cpu["rate"], "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu["sample"]}
                if args.workload == "c3":
                    cpu_baseline["reference_as_written_c1"] = reference_as_written_c1()

    # ---- the library the kernels replace, on the same GPU (comparator only) ----------------------------------------------
    vs_cusparse = None
    if world == 1 and is_csr and not args.no_cusparse:
        try:
            from baseline import cusparse_ref

            vs_cusparse = cusparse_ref.compare(rt, mat, n, nnz, stream, ours_spmv_ms=spmv_ms, ours_iteration_ms=ms_per_step / ipt, solver=solver)
        except Exception as exc:  # the comparator is optional equipment: say why it is missing
            vs_cusparse = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    if rank == 0:
        cfg = shared_config(args, n, nnz)
        cfg.update({
            "solver_form": "fused (fewest HBM passes: CG 3, BiCGStab 5)" if not args.unfused else "unfused (reference call sequence)",
            "gmres_restart": restart if solver == "gmres" else None,
            "trace": "CUDA graph replay of iters_per_step iterations", "rhs": "b = 1, x0 = 0 (BenchmarkStencil)",
            "l2": "working set per GPU exceeds the 126 MB L2 (matrix streamed once per iteration)" if (16 if is_csr else 24) * nnz_local > 2 * 126e6
                  else "matrix slab of this rank fits the L2 only partly; vectors are L2-resident",
            "halo_bytes_per_matvec_rank0": pl.halo_bytes_per_matvec, "setup_seconds": round(setup_s, 3),
            "collectives": "none (1 GPU)" if world == 1 else rt.collectives,
            "comm_error": rt.comm_error() if world > 1 else 0,
            "time_inside_collectives": comm_us,
            "spmv_ms_per_launch_by_rank": spmv_ms_by_rank,
            "iteration_roofline": {"bytes_per_iteration_per_gpu": iter_bytes, "frac_of_peak": iter_frac,
                                   "frac_of_nominal_8TBs": iter_bytes * value / 1e9 / 8000.0},
        })
        line = {
            "metric": metric_name(solver), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms},
            "spmv_gbs": achieved,
            "parity": parity,
            "cpu_baseline": cpu_baseline,
            "vs_cusparse": vs_cusparse,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_cusparse_arm(args):
    """The reference's GPU call sequence through cuSPARSE / cuBLAS 12.9 (comparator; never reachable from the package)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from baseline import cusparse_ref

    line = cusparse_ref.bench_line(args, WORKLOADS, shared_config, metric_name)
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference", "cusparse"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c3")
    ap.add_argument("--iters-per-step", type=int, default=20, help="solver iterations per recorded trace (BenchmarkStencil -pt)")
    ap.add_argument("--ref-iters-per-step", type=int, default=2, help="CPU reference arm: iterations per timed step (bounded sample)")
    ap.add_argument("--solver", choices=["cg", "bicgstab", "gmres"], default=None, help="default: the workload's solver (c3/c2 cg, c4 bicgstab, c5 gmres)")
    ap.add_argument("--spaces", type=int, default=1, choices=[1, 2], help="2 = BenchmarkStencil's doubled block-diagonal system")
    ap.add_argument("--shape", type=str, default=None, help="developer override nx,ny,nz of the workload's grid (c5: log2 N)")
    ap.add_argument("--unfused", action="store_true", help="run the reference's unfused call sequence on the GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size comparison with the CPU oracle")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cusparse", action="store_true")
    args = ap.parse_args()
    if args.workload == "c5" and args.iters_per_step == 20:
        args.iters_per_step = 2  # a GMRES(30) restart cycle is ~30 iterations' worth of work
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "cusparse":
        return run_cusparse_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
