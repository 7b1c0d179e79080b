"""N > 1: (gpu) torchrun with 2 ranks on 2 GPUs when the box has them; (cpu) world_size-2 / -4 gloo tests of the
host-side logic of the same path, THROUGH THE PRODUCT'S OWN entry points (lsk_equal_partition, lsk_shard, lsk_halo_plan --
the function SquarePlanner::add_row_partitioned_matrix calls): partition -> ghost intervals -> halo plan -> SpMV / CG."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["peer-memory", "nccl"])
def test_multi_gpu_cg_bicgstab_gmres_parity(comm):
    """Every rank checks generator slab, partitions, solver histories and its owned solution rows against
    the single-process oracle; once with the peer-memory collectives (CUDA IPC over NVLink), once on NCCL."""
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if ngpu >= 4 else 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if comm == "nccl":
        env["LSK_COMM"] = "nccl"
    else:
        env.pop("LSK_COMM", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29511" if comm == "nccl" else "29512", str(ROOT / "tests" / "mp_worker.py")]
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        env["LSK_MP_OUT"] = tmp
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
        lines = [json.loads(f.read_text()) for f in sorted(Path(tmp).glob("rank*.json"))]
    assert proc.returncode == 0, (lines, proc.stdout[-3000:] + proc.stderr[-3000:])
    assert len(lines) == world and all(l["ok"] for l in lines), lines
    assert all(l["results"]["_comm"]["error"] == 0 for l in lines)
    if comm == "nccl":
        assert not any(l["results"]["_comm"]["peer_memory"] for l in lines)
    by_rank = {l["rank"]: l for l in lines}
    # 3-D 7-point, row-major: rank 0 receives exactly one ny*nz plane from its single neighbour
    assert by_rank[0]["results"]["cg_7pt"]["halo_bytes"] == 20 * 16 * 8


# ---- CPU: the host logic of the N > 1 path over gloo, world_size = 2 ---------------------------------------
def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C

        from legionsolvers_b200 import _abi

        L = _abi.lib()  # the PRODUCT's host-side entry points (no GPU needed for these)
        off, val = orc.benchmark_stencil(3)
        shape = (8, 6, 5)
        m = orc.stencil_csr(shape, off, val)
        n, pieces = m.n_rows, world
        # create_equal_partition and the blocked sharding rule: the product's functions, checked against the oracle's
        lo, hi = np.zeros(pieces, dtype=np.int64), np.zeros(pieces, dtype=np.int64)
        assert L.lsk_equal_partition(n, pieces, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)) == 0
        olo, ohi = orc.equal_partition(n, pieces)
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
        mine = [c for c in range(pieces) if L.lsk_shard(c, pieces, world) == rank]
        assert mine == [rank] == [c for c in range(pieces) if orc.shard(c, pieces, world) == rank]
        pl = orc.Planner([n], [pieces])
        b = pl.add_matrix(m)
        own_lo, own_hi = int(lo[rank]), int(hi[rank])
        g_lo, g_hi = pl.ghost_bounds(b, rank)
        # every rank learns every rank's owned rows and ghost interval (the planner's all-gather)
        mine_t = torch.tensor([own_lo, own_hi, g_lo, g_hi], dtype=torch.int64)
        allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine_t)
        allr = [tuple(int(v) for v in t) for t in allr]
        # the halo plan: SquarePlanner::add_row_partitioned_matrix's own arithmetic, through its C entry point
        ranges = np.ascontiguousarray(np.array(allr, dtype=np.int64).reshape(world, 4))
        moves5 = np.zeros((world, 5), dtype=np.int64)
        nmoves = C.c_int(0)
        assert L.lsk_halo_plan(rank, world, ranges.ctypes.data_as(C.c_void_p), moves5.ctypes.data_as(C.c_void_p), C.byref(nmoves)) == 0
        planned = {int(r[0]): tuple(int(v) for v in r[1:]) for r in moves5[:nmoves.value]}
        # (peers with nothing to trade are not in the plan; the checks below want one entry per peer)
        moves = [(q_,) + planned.get(q_, (0, 0, 0, 0)) for q_ in range(world) if q_ != rank]
        # run 12 CG iterations with a DISTRIBUTED x: each rank holds only [g_lo, g_hi] of p
        x_full = np.zeros(n); r = np.ones(n); p = np.ones(n); qv = np.zeros(n)
        own = slice(own_lo, own_hi + 1)
        rr = torch.tensor([float(r[own] @ r[own])], dtype=torch.float64)
        dist.all_reduce(rr)
        hist = [float(rr)]
        for _ in range(12):
            # halo exchange of p
            reqs = []
            for (q_, s_lo, s_n, r_lo, r_n) in moves:
                if s_n:
                    reqs.append(dist.isend(torch.from_numpy(p[s_lo:s_lo + s_n].copy()), q_))
            for (q_, s_lo, s_n, r_lo, r_n) in moves:
                if r_n:
                    buf = torch.zeros(r_n, dtype=torch.float64)
                    dist.recv(buf, q_)
                    p[r_lo:r_lo + r_n] = buf.numpy()
            for rq in reqs:
                rq.wait()
            pg = np.full(n, np.nan); pg[g_lo:g_hi + 1] = p[g_lo:g_hi + 1]  # poison everything outside the ghost piece
            qv[own] = 0.0
            k_lo, k_hi = pl.kernel_bounds(b, rank)
            orc.csr_matvec(m, pg, qv, k=(k_lo, k_hi), r=(own_lo, own_hi), cols=(g_lo, g_hi))
            pq = torch.tensor([float(p[own] @ qv[own])], dtype=torch.float64); dist.all_reduce(pq)
            alpha = hist[-1] / float(pq)
            x_full[own] += alpha * p[own]; r[own] -= alpha * qv[own]
            rn = torch.tensor([float(r[own] @ r[own])], dtype=torch.float64); dist.all_reduce(rn)
            p[own] = r[own] + (float(rn) / hist[-1]) * p[own]
            hist.append(float(rn))
        # single-process oracle on the same system
        opl = orc.Planner([n], [pieces]); opl.fill(1, 1.0); opl.add_matrix(m)
        ocg = orc.CGSolver(opl)
        for _ in range(12):
            ocg.step()
        want = ocg.residual_norm_squared
        err = float(np.max(np.abs(np.array(hist) - want) / want))
        xerr = float(np.max(np.abs(x_full[own] - opl.vector(0)[own])))
        q.put((rank, err, xerr, moves, (g_lo, g_hi)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_plan_and_cg():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29533
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, e0, x0, m0, g0), (r1, e1, x1, m1, g1) = out
    assert e0 <= 1e-12 and e1 <= 1e-12 and x0 <= 1e-12 and x1 <= 1e-12
    # plans are consistent: what rank 0 sends to 1 is what rank 1 receives from 0, and vice versa
    (_, s0_lo, s0_n, r0_lo, r0_n), = m0
    (_, s1_lo, s1_n, r1_lo, r1_n), = m1
    assert (s0_lo, s0_n) == (r1_lo, r1_n) and (s1_lo, s1_n) == (r0_lo, r0_n)
    # 3-D 7-point, row-major: the halo is exactly one ny*nz plane each way
    assert s0_n == s1_n == 6 * 5
    assert g0 == (0, 4 * 30 + 30 - 1) and g1 == (4 * 30 - 30, 8 * 30 - 1)


def test_gloo_world4_halo_plan_and_cg():
    """Four ranks over gloo: the two middle ranks trade a plane with BOTH neighbours and nothing with the others;
    distributed CG on the planned halo equals the single-process oracle on every rank."""
    import torch.multiprocessing as mp

    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, 29541, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plane = 6 * 5
    plans = {}
    for rank, err, xerr, moves, ghost in out:
        assert err <= 1e-12 and xerr <= 1e-12, (rank, err, xerr)
        plans[rank] = {peer: (s_lo, s_n, r_lo, r_n) for peer, s_lo, s_n, r_lo, r_n in moves}
        # ghost interval = owned rows (2 planes) plus one plane on each interior side
        lo, hi = rank * 2 * plane, (rank + 1) * 2 * plane - 1
        assert ghost == (max(0, lo - plane), min(world * 2 * plane - 1, hi + plane))
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            s_lo, s_n, _, _ = plans[a][b]
            _, _, r_lo, r_n = plans[b][a]
            assert (s_lo, s_n) == (r_lo, r_n) or (s_n == 0 and r_n == 0)   # what a sends to b is what b receives from a
            assert s_n == (plane if abs(a - b) == 1 else 0)                  # one plane to each neighbour, nothing further
