"""The peer-memory collectives on ONE GPU: two "ranks" in one process -- two contexts, two comm windows, two streams -- so
that the kernels of the multi-GPU path (LL all-reduce, LL halo exchange stand-alone and fused into xpay / the CG direction
update, the deferred all-reduce) are exercised by the single-GPU test tier as well.  Rank r's kernels run on stream r;
a kernel of one rank polls for packets the other rank's kernel sends, so the two streams must make progress
concurrently (they do: every collective kernel is sized to be resident at once).  A protocol error shows up as the
4 s spin-wait timeout (comm error flag + NaN), never as a hang.

Through the C ABI (include/lsk.h), like every other GPU test.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Rank:
    def __init__(self, rank, nranks, windows):
        from legionsolvers_b200 import _abi
        from legionsolvers_b200.kernels import Context

        self.rank = rank
        self.ctx = Context()
        self.stream = torch.cuda.Stream()
        self.peers = _abi.Peers()
        self.peers.rank, self.peers.nranks = rank, nranks
        for r in range(nranks):
            self.peers.window[r] = windows[r].data_ptr()

    def set_peers(self, on=True):
        from legionsolvers_b200 import _abi

        _abi.check(_abi.lib().lsk_ctx_set_peers(self.ctx.h, C.byref(self.peers) if on else None), "lsk_ctx_set_peers")

    def comm_error(self):
        from legionsolvers_b200 import _abi

        e = C.c_int(0)
        _abi.check(_abi.lib().lsk_comm_error(self.ctx.h, self.stream.cuda_stream, C.byref(self.peers), C.byref(e)), "lsk_comm_error")
        return e.value


@pytest.fixture()
def ranks():
    from legionsolvers_b200 import _abi

    nbytes = _abi.lib().lsk_comm_window_bytes()
    windows = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    rs = [Rank(r, 2, windows) for r in range(2)]
    torch.cuda.synchronize()
    yield rs
    torch.cuda.synchronize()
    for r in rs:
        r.set_peers(False)
        r.ctx.close()


def landing(count):
    from legionsolvers_b200 import _abi

    return torch.zeros(_abi.lib().lsk_halo_landing_bytes(count), dtype=torch.uint8, device="cuda")


def make_moves(send, recv, land):
    """send[r] / recv[r]: tensors rank r sends to / receives from the other rank; land[r]: rank r's landing buffer."""
    from legionsolvers_b200 import _abi

    out = []
    for r in range(2):
        m = (_abi.HaloMove * 1)()
        m[0].peer = 1 - r
        m[0].n, m[0].src = send[r].numel(), send[r].data_ptr() if send[r].numel() else None
        m[0].recv_n, m[0].recv_dst = recv[r].numel(), recv[r].data_ptr() if recv[r].numel() else None
        m[0].ll_send, m[0].ll_recv = land[1 - r].data_ptr(), land[r].data_ptr()
        out.append(m)
    return out


def test_allreduce_two_ranks_one_gpu(ranks):
    from legionsolvers_b200 import _abi

    L = _abi.lib()
    slots = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(2)]
    for it in range(7):  # both parities of the packet slots, several times over
        vals = [np.array([1.5 + it, -2.25 * it]), np.array([1e-3 * it, 7.0])]
        for r in range(2):
            slots[r].copy_(torch.from_numpy(vals[r]))
        torch.cuda.synchronize()
        for r in ranks:
            _abi.check(L.lsk_allreduce_sum_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), slots[r.rank].data_ptr(), 2), "allreduce")
        torch.cuda.synchronize()
        want = vals[0] + vals[1]  # rank-order sum of two terms: exact either way
        for r in range(2):
            np.testing.assert_array_equal(slots[r].cpu().numpy(), want)
    assert all(r.comm_error() == 0 for r in ranks)


@pytest.mark.parametrize("n01,n10", [(1000, 37), (65536, 65536), (513, 0), (0, 129), (1, 1)])
def test_halo_exchange_two_ranks_one_gpu(ranks, n01, n10):
    """Stand-alone exchange, asymmetric and one-way moves (the token still travels both ways), six exchanges in a row with
    NO other synchronisation between them: both halves of the landing buffers are reused three times."""
    from legionsolvers_b200 import _abi

    L = _abi.lib()
    gen = torch.Generator(device="cuda").manual_seed(5)
    send = [torch.zeros(n01, dtype=torch.float64, device="cuda"), torch.zeros(n10, dtype=torch.float64, device="cuda")]
    recv = [torch.zeros(n10, dtype=torch.float64, device="cuda"), torch.zeros(n01, dtype=torch.float64, device="cuda")]
    land = [landing(n10), landing(n01)]
    moves = make_moves(send, recv, land)
    history = []
    for it in range(6):
        for r in range(2):
            if send[r].numel():
                send[r].copy_(torch.rand(send[r].numel(), dtype=torch.float64, device="cuda", generator=gen) - 0.5)
        torch.cuda.synchronize()
        history.append([s.clone() for s in send])
        for r in ranks:
            _abi.check(L.lsk_halo_exchange_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), moves[r.rank], 1), "halo exchange")
        torch.cuda.synchronize()
        assert torch.equal(recv[0], send[1]) and torch.equal(recv[1], send[0])  # bit-exact, including signs / tiny values
    assert all(r.comm_error() == 0 for r in ranks)


def test_halo_reduce_two_ranks_one_gpu(ranks):
    """lsk_halo_reduce_f64, the reverse exchange of a transposed mat-vec: what arrives is ADDED to the destination."""
    from legionsolvers_b200 import _abi

    L = _abi.lib()
    gen = torch.Generator(device="cuda").manual_seed(9)
    n01, n10 = 777, 4096
    rnd = lambda k: torch.rand(k, dtype=torch.float64, device="cuda", generator=gen) - 0.5  # noqa: E731
    send = [rnd(n01), rnd(n10)]
    recv = [rnd(n10), rnd(n01)]
    land = [landing(n10), landing(n01)]
    moves = make_moves(send, recv, land)
    want = [recv[0].clone(), recv[1].clone()]
    for it in range(4):
        torch.cuda.synchronize()
        for r in ranks:
            _abi.check(L.lsk_halo_reduce_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), moves[r.rank], 1), "halo reduce")
        torch.cuda.synchronize()
        want[0] += send[1]
        want[1] += send[0]
        assert torch.equal(recv[0], want[0]) and torch.equal(recv[1], want[1])  # one rounded addition per element and exchange
    assert all(r.comm_error() == 0 for r in ranks)


def test_halo_exchange_rejects_bad_moves(ranks):
    from legionsolvers_b200 import _abi

    L = _abi.lib()
    r = ranks[0]
    a = torch.zeros(8, dtype=torch.float64, device="cuda")
    land = [landing(8), landing(8)]
    m = make_moves([a, a], [a.clone(), a.clone()], land)[0]
    m[0].peer = 0  # a move to oneself
    assert L.lsk_halo_exchange_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), m, 1) == -1
    m[0].peer = 1
    m[0].ll_recv = land[0].data_ptr() + 8  # landing buffers are read with 16-byte loads
    assert L.lsk_halo_exchange_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), m, 1) == -1
    m[0].ll_recv = None
    assert L.lsk_halo_exchange_f64(r.ctx.h, r.stream.cuda_stream, C.byref(r.peers), m, 1) == -1


@pytest.mark.parametrize("n,lo_send,hi_send", [(5000, 64, 64), (200_001, 4096, 4099), (70, 3, 5)])
@pytest.mark.parametrize("off", [0, 1])
def test_xpay_halo_two_ranks_one_gpu(ranks, oracle, n, lo_send, hi_send, off):
    """lsk_xpay_halo_f64 on both ranks: y = fma(alpha, y, x) bit-exact as the plain xpay, the first `lo_send` elements of
    rank 1's y land in rank 0's upper ghosts and the last `hi_send` of rank 0's y in rank 1's lower ghosts."""
    from legionsolvers_b200 import _abi

    L = _abi.lib()
    rng = np.random.default_rng(n + off)
    for r in ranks:
        r.set_peers()
    # rank r's buffer: [lower ghosts | owned n | upper ghosts], owned part starting `off` doubles past a 32-byte boundary
    glo, ghi = [0, hi_send], [lo_send, 0]
    bufs, xs, ys_ref = [], [], []
    for r in range(2):
        b = torch.zeros(glo[r] + n + ghi[r] + 8, dtype=torch.float64, device="cuda")
        bufs.append(b)
        x0, y0 = rng.standard_normal(n), rng.standard_normal(n)
        xs.append(torch.from_numpy(x0).cuda())
        b[off + glo[r]: off + glo[r] + n].copy_(torch.from_numpy(y0))
        ys_ref.append((x0, y0))
    own = [bufs[r][off + glo[r]: off + glo[r] + n] for r in range(2)]
    send = [own[0][n - hi_send:], own[1][:lo_send]]
    recv = [bufs[0][off + n: off + n + lo_send], bufs[1][off: off + hi_send]]
    land = [landing(lo_send), landing(hi_send)]
    moves = make_moves(send, recv, land)
    num, den = torch.tensor([3.0], dtype=torch.float64, device="cuda"), torch.tensor([7.0], dtype=torch.float64, device="cuda")
    alpha = oracle.get_alpha([3.0, 7.0])
    for it in range(3):
        torch.cuda.synchronize()
        for r in ranks:
            _abi.check(L.lsk_xpay_halo_f64(r.ctx.h, r.stream.cuda_stream, n, 2, num.data_ptr(), den.data_ptr(), None, None,
                                           xs[r.rank].data_ptr(), own[r.rank].data_ptr(), moves[r.rank], 1), "xpay_halo")
        torch.cuda.synchronize()
        for r in range(2):
            x0, y0 = ys_ref[r]
            oracle.xpay(alpha, x0, y0)  # y0 <- fma(alpha, y0, x0)
            np.testing.assert_array_equal(own[r].cpu().numpy(), y0)
        np.testing.assert_array_equal(recv[0].cpu().numpy(), ys_ref[1][1][:lo_send])
        np.testing.assert_array_equal(recv[1].cpu().numpy(), ys_ref[0][1][n - hi_send:])
    assert all(r.comm_error() == 0 for r in ranks)


@pytest.mark.parametrize("shape", [(24, 20, 16), (64, 32, 32)], ids=["24x20x16", "64x32x32 (several CTAs per kernel)"])
@pytest.mark.parametrize("tail", [False, True], ids=["update + direction", "one-launch tail"])
def test_fused_cg_step_two_ranks_one_gpu(ranks, oracle, tail, shape):
    """The whole fused CG step of a row-partitioned system on two ranks that share one GPU: mat-vec with fused p.q (deferred
    all-reduce: sent by the mat-vec, resolved by the update kernel), x / r update with fused r.r (sent by the update kernel,
    resolved by the direction kernel), direction update with the halo exchange of p inside.  Residual history within 1e-10
    of the oracle's single-process CG, solution within 1e-10.  tail: lsk_cg_tail_f64 instead of the last two launches (its
    r.r all-reduce happens inside the kernel, between its two phases)."""
    from legionsolvers_b200 import _abi
    from legionsolvers_b200 import kernels as K

    L = _abi.lib()
    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr(shape, off, val)
    n, its = m.n_rows, 30
    plane = shape[1] * shape[2]
    half = (shape[0] // 2) * plane
    own_lo, own_n = [0, half], [half, n - half]
    g_lo, g_hi = [0, half - plane], [half + plane - 1, n - 1]  # rows a rank's columns reach
    opl = oracle.Planner([n], [2])
    opl.fill(1, 1.0)
    opl.add_matrix(m)
    ocg = oracle.CGSolver(opl)
    for _ in range(its):
        ocg.step()
    want = ocg.residual_norm_squared

    st = []
    rp = np.ascontiguousarray(m.rowptr).view(np.int64).reshape(-1, 2)
    for r in range(2):
        k_lo, k_hi = int(rp[own_lo[r], 0]), int(rp[own_lo[r] + own_n[r] - 1, 1])
        z = lambda cnt: torch.zeros(cnt, dtype=torch.float64, device="cuda")  # noqa: E731
        p_full = z(g_hi[r] - g_lo[r] + 1)
        d = dict(entry=torch.from_numpy(m.entry[k_lo:k_hi + 1].copy()).cuda(), col=torch.from_numpy(m.col[k_lo:k_hi + 1].copy()).cuda(),
                 rowptr=K.rect_tensor(m.rowptr[own_lo[r]:own_lo[r] + own_n[r]]), k_lo=k_lo, nnz=k_hi - k_lo + 1, p_full=p_full,
                 p=p_full[own_lo[r] - g_lo[r]: own_lo[r] - g_lo[r] + own_n[r]], x=z(own_n[r]), r=z(own_n[r]), q=z(own_n[r]),
                 rr_cur=z(1), rr_new=z(1), pq=z(1), hist=z(its + 1), count=torch.zeros(1, dtype=torch.int64, device="cuda"))
        d["p_full"].fill_(1.0)  # P = RHS = 1 everywhere, ghosts included: the state after reset() + refresh_halo
        d["r"].fill_(1.0)
        d["rr_cur"].fill_(float(n))
        st.append(d)
    # halo plan of the two slabs: one plane each way
    send = [st[0]["p"][own_n[0] - plane:], st[1]["p"][:plane]]
    recv = [st[0]["p_full"][own_n[0]:own_n[0] + plane], st[1]["p_full"][:plane]]
    land = [landing(plane), landing(plane)]
    moves = make_moves(send, recv, land)
    # One process, one CUDA context: the first launch of a kernel loads it lazily (and the first use per lsk_ctx sets its
    # shared-memory attribute), which may wait for the device to drain -- fatal while the other "rank" spins for this one.
    # So every kernel of the step runs once per context without peers first.  (Ranks in separate processes do not share a
    # context and cannot block each other this way.)
    for rk in ranks:
        d = st[rk.rank]
        with torch.cuda.stream(rk.stream):
            tmp = {k: d[k].clone() for k in ("p_full", "x", "r", "q", "rr_cur", "rr_new", "pq", "hist", "count")}
            tp = tmp["p_full"][own_lo[rk.rank] - g_lo[rk.rank]: own_lo[rk.rank] - g_lo[rk.rank] + own_n[rk.rank]]
            rk.ctx.csr_spmv(own_n[rk.rank], d["nnz"], d["entry"], d["col"], d["rowptr"], d["k_lo"], tmp["p_full"], g_lo[rk.rank], tmp["q"],
                            dot_w=tp, dot_out=tmp["pq"])
            rk.ctx.cg_update(tmp["rr_cur"], tmp["pq"], tp, tmp["q"], tmp["x"], tmp["r"], tmp["rr_new"])
            rk.ctx.cg_direction(tmp["rr_cur"], tmp["rr_new"], tmp["r"], tp, tmp["hist"], tmp["count"])
            rk.ctx.cg_tail(tmp["rr_cur"], tmp["pq"], tmp["rr_new"], tp, tmp["q"], tmp["x"], tmp["r"], tmp["hist"], tmp["count"])
    torch.cuda.synchronize()
    for r in ranks:
        r.set_peers()
    for it in range(its):
        for rk in ranks:
            d, h, s = st[rk.rank], rk.ctx.h, rk.stream.cuda_stream
            with torch.cuda.stream(rk.stream):
                _abi.check(L.lsk_ctx_defer_next_allreduce(h), "defer")
                rk.ctx.csr_spmv(own_n[rk.rank], d["nnz"], d["entry"], d["col"], d["rowptr"], d["k_lo"], d["p_full"], g_lo[rk.rank], d["q"],
                                dot_w=d["p"], dot_out=d["pq"])
                if tail:
                    _abi.check(L.lsk_cg_tail_f64(h, s, own_n[rk.rank], d["rr_cur"].data_ptr(), d["pq"].data_ptr(), d["rr_new"].data_ptr(),
                                                 d["p"].data_ptr(), d["q"].data_ptr(), d["x"].data_ptr(), d["r"].data_ptr(), moves[rk.rank], 1,
                                                 d["hist"].data_ptr(), its + 1, d["count"].data_ptr()), "cg_tail")
                    continue
                _abi.check(L.lsk_ctx_defer_next_allreduce(h), "defer")
                rk.ctx.cg_update(d["rr_cur"], d["pq"], d["p"], d["q"], d["x"], d["r"], d["rr_new"])
                assert L.lsk_cg_direction_supported(own_n[rk.rank], d["r"].data_ptr(), d["p"].data_ptr())
                _abi.check(L.lsk_cg_direction_f64(h, s, own_n[rk.rank], d["rr_cur"].data_ptr(), d["rr_new"].data_ptr(), d["r"].data_ptr(),
                                                  d["p"].data_ptr(), moves[rk.rank], 1, d["hist"].data_ptr(), its + 1, d["count"].data_ptr()),
                           "cg_direction")
    torch.cuda.synchronize()
    assert all(r.comm_error() == 0 for r in ranks)
    for r in range(2):
        got = st[r]["hist"][:its].cpu().numpy()
        assert np.max(np.abs(got - want[1:its + 1]) / want[1:its + 1]) <= 1e-10
    x = np.concatenate([st[0]["x"].cpu().numpy(), st[1]["x"].cpu().numpy()])
    xo = opl.vector(0)
    assert np.max(np.abs(x - xo)) <= 1e-10 * np.max(np.abs(xo))
    # identical bits on both ranks: the cross-rank sums are added in rank order everywhere
    assert torch.equal(st[0]["hist"], st[1]["hist"])
