"""Multi-GPU parity worker, launched by tests/test_multi_gpu.py under torchrun (one rank per GPU):
row-partitioned CG / BiCGStab / GMRES with NCCL halo exchange + all-reduced dots, checked on every
rank against the single-process CPU oracle."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legionsolvers_b200 import solvers as S  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    rt = S.Runtime(device=local, rank=rank, nranks=world, stream=ts.cuda_stream)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(S.Runtime.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    rt.comm_init(uid.cpu().numpy().tobytes())

    results = {}
    trace = iter(range(100, 10000))
    for name, dim_flag, shape, ppr, solver, its in [
        ("cg_7pt", 3, (24, 20, 16), 1, "cg", 40),
        ("cg_7pt_unfused", 3, (24, 20, 16), 1, "cg_unfused", 40),
        ("cg_7pt_2pieces_per_rank", 3, (24, 20, 16), 2, "cg", 40),
        ("cg_27pt", 4, (16, 16, 16), 1, "cg", 30),
        ("cg_2d", 2, (64, 48, 1), 1, "cg", 50),
        ("bicgstab_27pt", 4, (16, 16, 16), 1, "bicgstab", 12),
        ("gmres_7pt", 3, (16, 12, 12), 1, "gmres", 1),
    ]:
        pieces = world * ppr
        n = shape[0] * shape[1] * shape[2]
        st = S.benchmark_stencil(dim_flag, *shape)
        mat = S.CSRMatrix.stencil(rt, st, pieces)
        sol, rhs = S.PartitionedVector(rt, "sol", n, pieces), S.PartitionedVector(rt, "rhs", n, pieces)
        sol.zero_fill()
        rhs.constant_fill(1.0)
        pl = S.SquarePlanner(rt)
        pl.add_sol_vector(sol)
        pl.add_rhs_vector(rhs)
        pl.add_row_partitioned_matrix(mat, 0, 0)
        # oracle, whole problem in this process
        off, val = orc.benchmark_stencil(dim_flag)
        dims = shape[:3] if dim_flag >= 3 else shape[:dim_flag]
        m = orc.stencil_csr(dims, off, val)
        opl = orc.Planner([n], [pieces])
        opl.fill(1, 1.0)
        opl.add_matrix(m)
        # this rank's slab of the generated matrix is bit-identical to the oracle's
        e, c, rp = mat.slab_to_numpy()
        ok_gen = (np.array_equal(e, m.entry[mat.slab_k_lo:mat.slab_k_hi + 1]) and np.array_equal(c, m.col[mat.slab_k_lo:mat.slab_k_hi + 1])
                  and np.array_equal(rp, m.rowptr[mat.slab_r_lo:mat.slab_r_hi + 1]))
        # partitions of the local colours are the oracle's
        first, end = pl.local_colors(0)
        ok_part = all(pl.range_bounds(0, col) == opl.piece_bounds(0, col) and pl.kernel_bounds(0, col) == opl.kernel_bounds(0, col)
                      and pl.ghost_bounds(0, col) == opl.ghost_bounds(0, col) for col in range(first, end))
        tid = next(trace)
        if solver == "cg":
            s, o = S.CGSolver(pl), orc.CGSolver(opl)
        elif solver == "cg_unfused":  # the reference's call sequence: stand-alone halo exchange before every mat-vec
            s, o = S.CGSolver(pl, fused=False), orc.CGSolver(opl)
            solver = "cg"
        elif solver == "bicgstab":
            s, o = S.BiCGStabSolver(pl), orc.BiCGStabSolver(opl)
        else:
            s, o = S.GMRESSolver(pl, 8), orc.GMRESSolver(opl, 8)
        for _ in range(its):
            rt.begin_trace(tid)
            s.step()
            rt.end_trace(tid)
            o.step()
        if solver == "cg":
            got, want = s.residual_norm_squared, o.residual_norm_squared
            err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12 * want[0])))
            tol = 1e-10
        elif solver == "bicgstab":
            got, want = s.rho, o.rho
            err = float(np.max(np.abs(got[1:9] - want[1:9]) / np.abs(want[1:9])))
            tol = 1e-10
        else:
            got, want = s.inner_products, o.inner_products
            err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
            tol = 1e-10
        lo, hi = sol.owned_range()
        x, xo = sol.to_numpy()[lo:hi + 1], opl.vector(0)[lo:hi + 1]
        xerr = float(np.max(np.abs(x - xo)) / np.max(np.abs(opl.vector(0))))
        if name in ("cg_7pt", "cg_7pt_2pieces_per_rank", "cg_27pt"):
            # transposed mat-vec across ranks: y = A^T x; a rank's contributions to columns it does not own accumulate in its
            # ghost elements and travel back to their owners in the reverse halo exchange (lsk_halo_reduce_f64)
            pl.rmatvec(3, 0)
            yt = pl.vector_to_numpy(3, 0, n)[lo:hi + 1]
            want_t = np.zeros(n)
            orc.rmatvec(m, opl.vector(0), want_t)
            xerr = max(xerr, float(np.max(np.abs(yt - want_t[lo:hi + 1])) / np.max(np.abs(want_t))))
        results[name] = {"hist_err": err, "tol": tol, "x_err": xerr, "gen": bool(ok_gen), "part": bool(ok_part),
                         "halo_bytes": pl.halo_bytes_per_matvec, "n_hist": int(np.size(got))}
    comm_err = rt.comm_error()
    results["_comm"] = {"peer_memory": rt.uses_peer_memory, "mode": rt.collectives, "error": comm_err, "hist_err": 0.0, "tol": 1.0, "x_err": 0.0,
                        "gen": comm_err == 0, "part": True}
    ok = all(r["hist_err"] <= r["tol"] and r["x_err"] <= 1e-8 and r["gen"] and r["part"] for r in results.values())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out_dir = os.environ.get("LSK_MP_OUT")
    payload = json.dumps({"rank": rank, "ok": ok, "results": results})
    if out_dir:  # one file per rank: stdout of concurrent ranks can interleave
        Path(out_dir, f"rank{rank}.json").write_text(payload)
    else:
        print(payload, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
