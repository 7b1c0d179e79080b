"""CPU model check of the peer-memory protocols of the multi-GPU solver step (legionsolvers_b200/csrc/lsk_common.cuh
`allreduce_warp` / `allreduce_send` + `allreduce_resolve`, `halo_begin` / `ll_store` / `halo_unpack` / `halo_finish`).

The kernels cannot run without GPUs, but what makes them correct is a small protocol that can be executed exhaustively
enough on the CPU: every rank is a generator that yields before each remote store / poll, a seeded random scheduler
interleaves them, and assertions check the properties the kernels rely on.  Everything that crosses NVLink is an LL
packet (data + the number of the exchange it belongs to in one atomic word), there is no fence and no flag:

  1. all-reduce with TWO packet slots per source (by epoch parity): a poll never accepts a stale packet, and a packet
     is never overwritten before its reader has consumed it (a peer can run at most one reduction ahead);
  2. halo exchange into a LANDING BUFFER owned by the receiver, two halves by the parity of the pair's exchange number:
     the receiver itself copies the packets into its ghost region, so ghost values are never written while the
     receiver's mat-vec reads them; a half is never overwritten before the receiver has unpacked it -- because every
     exchange of a pair carries a token packet in BOTH directions (data or not), a rank can start exchange e + 1 only
     after its peer has started exchange e, i.e. finished unpacking e - 1, which lives in the half e + 1 writes;
  3. exchanges are numbered per PAIR of ranks, so ranks that skip an exchange (different halos per block) do not
     desynchronise the others.

The model mirrors the kernels' state (epochs, parity slots, pair counters) one to one; it is not product code.
"""
import random

import pytest

NPK = 3  # data packets per halo move in the model (the token is packet NPK)


class Window:  # CommWindow + landing buffers of one rank
    def __init__(self, nranks):
        self.ar_pkt = [[(0, None)] * nranks for _ in range(2)]  # [parity][source] = (epoch, value)
        self.ar_consumed = [[0] * nranks for _ in range(2)]     # bookkeeping of the model: last epoch read per slot
        # landing[parity][source][i] = (tag, value); unpacked[parity][source][i] = last tag the receiver consumed
        self.landing = [[[(0, None)] * (NPK + 1) for _ in range(nranks)] for _ in range(2)]
        self.unpacked = [[[0] * (NPK + 1) for _ in range(nranks)] for _ in range(2)]
        self.halo_sent = [0] * nranks                           # pair counters
        self.ghost_version = {}                                 # neighbour -> iteration of the values in my ghost region
        self.reading_ghosts = False                             # my mat-vec is in flight


def ll_send(me, wins, peer, e, values, token=True):
    """Sender side of one move of exchange e: packets (and the token) into the peer's landing half e & 1."""
    half = wins[peer].landing[e & 1][me]
    seen = wins[peer].unpacked[e & 1][me]
    idx = list(range(len(values))) + ([NPK] if token else [])
    for i in idx:
        old_tag, _ = half[i]
        # property 2: the packet being overwritten (exchange e - 2) has been unpacked by its receiver
        assert seen[i] >= old_tag, f"rank {me} overwrites a packet of exchange {old_tag} that rank {peer} has not unpacked"
        half[i] = (e, values[i] if i < len(values) else 0.0)
        yield


def ll_unpack(me, wins, peer, e, count, token=True):
    """Receiver side: poll every packet of the peer's move in MY landing half until it carries tag e."""
    half = wins[me].landing[e & 1][peer]
    seen = wins[me].unpacked[e & 1][peer]
    out = []
    idx = list(range(count)) + ([NPK] if token else [])
    for i in idx:
        while True:
            tag, v = half[i]
            assert tag <= e, "a packet from the future in this half"
            if tag == e:
                break
            yield
        seen[i] = e
        if i < count:
            out.append(v)
    return out


def rank_program(me, nranks, wins, iters, log):
    """One rank's fused CG steps as a generator; yields where the GPU could be descheduled relative to its peers."""
    w = wins[me]
    nbrs = [r for r in (me - 1, me + 1) if 0 <= r < nranks]
    for nb in nbrs:
        w.ghost_version[nb] = 0
    ar_epoch = 0

    def allreduce(value):
        nonlocal ar_epoch
        e = ar_epoch + 1
        par = e & 1
        for r in range(nranks):  # lane r stores this rank's packet into rank r's window
            old_epoch, _ = wins[r].ar_pkt[par][me]
            # property 1b: the packet being overwritten (epoch e - 2) has been consumed by its reader
            assert wins[r].ar_consumed[par][me] >= old_epoch, f"rank {me} overwrites an unread packet of rank {r}"
            wins[r].ar_pkt[par][me] = (e, value)
            yield
        total = 0
        for r in range(nranks):  # lane r polls rank r's packet in MY window
            while True:
                got_epoch, got = w.ar_pkt[par][r]
                assert got_epoch <= e, "a packet from the future in this parity slot"  # property 1a
                if got_epoch == e:
                    break
                yield
            w.ar_consumed[par][r] = e
            total += got
        ar_epoch = e
        return total

    for k in range(iters):
        # ---- mat-vec of iteration k: reads the ghosts of P_k, which this rank's previous kernel put in place
        w.reading_ghosts = True
        for nb in nbrs:
            assert w.ghost_version[nb] == k, f"rank {me} iteration {k}: ghost of rank {nb} holds iteration {w.ghost_version[nb]}"
            yield
        w.reading_ghosts = False
        pq = yield from allreduce(float(me + 1))
        assert pq == nranks * (nranks + 1) / 2
        # ---- x / r update
        rr = yield from allreduce(float(2 * me + k))
        assert rr == sum(2 * r + k for r in range(nranks))
        # ---- direction update: P_{k+1}; boundary chunks first (packets out), unpacking last
        for nb in nbrs:
            e = w.halo_sent[nb] + 1
            yield from ll_send(me, wins, nb, e, [(me, k + 1, i) for i in range(NPK)])
        for nb in nbrs:
            e = w.halo_sent[nb] + 1
            got = yield from ll_unpack(me, wins, nb, e, NPK)
            assert got == [(nb, k + 1, i) for i in range(NPK)], "halo of an iteration skipped, repeated or torn"
            assert not w.reading_ghosts  # ghost values are written by their reader, in its own stream order
            w.ghost_version[nb] = k + 1
        for nb in nbrs:  # last CTA: the pair counters advance
            w.halo_sent[nb] += 1
        log.append((me, k))


def _run(progs, rng, limit=2_000_000):
    steps = 0
    while progs:
        # biased scheduler: sometimes let one rank run far ahead, which is what breaks naive protocols
        r = rng.choice(list(progs))
        for _ in range(rng.choice((1, 1, 2, 5, 40))):
            try:
                next(progs[r])
            except StopIteration:
                del progs[r]
                break
        steps += 1
        assert steps < limit, "no progress: the protocol dead-locked"


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_cg_step_protocol_under_random_interleavings(nranks):
    for seed in range(40):
        rng = random.Random(1000 * nranks + seed)
        wins = [Window(nranks) for _ in range(nranks)]
        log = []
        progs = {r: rank_program(r, nranks, wins, iters=6, log=log) for r in range(nranks)}
        _run(progs, rng)
        assert sorted(log) == [(r, k) for r in range(nranks) for k in range(6)]
        # nobody ever ran more than one iteration ahead of a neighbour
        done = {}
        for r, k in log:
            done[r] = k
            for nb in (r - 1, r + 1):
                if 0 <= nb < nranks:
                    assert done.get(nb, -1) >= k - 1


def test_model_detects_a_broken_allreduce():
    """Sanity of the model itself: with ONE packet slot instead of two, a fast rank overwrites an unread packet."""
    nranks = 3

    class OneSlot(Window):
        def __init__(self, n):
            super().__init__(n)
            self.ar_pkt[1] = self.ar_pkt[0]          # both parities share a slot
            self.ar_consumed[1] = self.ar_consumed[0]

    caught = 0
    for seed in range(200):
        rng = random.Random(seed)
        wins = [OneSlot(nranks) for _ in range(nranks)]
        progs = {r: rank_program(r, nranks, wins, iters=4, log=[]) for r in range(nranks)}
        try:
            _run(progs, rng, limit=100_000)
        except AssertionError:
            caught += 1
    assert caught > 0


# ---- back-to-back exchanges with NO all-reduce in between (the stand-alone lsk_halo_exchange_f64) -------------------------
def _exchange_only_program(me, wins, iters, one_way, token, log):
    """Two ranks, `iters` exchanges in a row.  one_way: only rank 0 has data for rank 1 (n == 0 the other way round)."""
    peer = 1 - me
    w = wins[me]
    for k in range(iters):
        e = w.halo_sent[peer] + 1
        n_send = NPK if (me == 0 or not one_way) else 0
        n_recv = NPK if (me == 1 or not one_way) else 0
        yield from ll_send(me, wins, peer, e, [(me, k, i) for i in range(n_send)], token=token)
        got = yield from ll_unpack(me, wins, peer, e, n_recv, token=token)
        assert got == [(peer, k, i) for i in range(n_recv)], "torn or stale halo"
        w.halo_sent[peer] += 1
        log.append((me, k))


@pytest.mark.parametrize("one_way", [False, True], ids=["data both ways", "data one way, token back"])
def test_back_to_back_exchanges_never_overwrite_a_half_in_use(one_way):
    for seed in range(200):
        rng = random.Random(seed)
        wins = [Window(2) for _ in range(2)]
        log = []
        progs = {r: _exchange_only_program(r, wins, 8, one_way, True, log) for r in range(2)}
        _run(progs, rng)
        assert sorted(log) == [(r, k) for r in range(2) for k in range(8)]
        # the token bounds the run-ahead: a rank completes exchange k only after its peer has STARTED exchange k
        done = {0: -1, 1: -1}
        for r, k in log:
            done[r] = k
            assert done[1 - r] >= k - 1


def test_model_detects_missing_token():
    """Without the token a rank that only SENDS never waits for its peer: it laps the receiver and overwrites a half of the
    landing buffer that has not been unpacked."""
    caught = 0
    for seed in range(200):
        rng = random.Random(seed)
        wins = [Window(2) for _ in range(2)]
        progs = {r: _exchange_only_program(r, wins, 8, True, False, []) for r in range(2)}
        try:
            _run(progs, rng, limit=100_000)
        except AssertionError:
            caught += 1
    assert caught > 0


# ---- exchanges are numbered PER PAIR of ranks (CommWindow::halo_sent) ------------------------------------------------------
def _pair_counter_program(me, nranks, wins, plans, use_pair_counters, log):
    """One rank issuing a sequence of halo exchanges; plans[x][me] = peers `me` trades with in exchange x (possibly none:
    the stand-alone exchange kernel is not even launched then).  Yields at every store / poll."""
    w = wins[me]
    single = 0           # the round-1 scheme: ONE exchange number per rank
    for x, plan in enumerate(plans):
        peers = plan[me]
        if not peers:
            continue     # lsk_halo_exchange_f64 returns without a launch: no counter moves
        single += 1
        for p in peers:
            e = w.halo_sent[p] + 1 if use_pair_counters else single
            half = wins[p].landing[e & 1][me]
            half[NPK] = (e, 0.0)  # the token is enough for this property
            yield
        for p in peers:
            e = w.halo_sent[p] + 1 if use_pair_counters else single
            spins = 0
            while wins[me].landing[e & 1][p][NPK][0] != e:
                spins += 1
                if spins > 2000:
                    log.append((me, x, p, "timeout"))
                    return
                yield
            w.halo_sent[p] += 1
    log.append((me, "done"))


@pytest.mark.parametrize("seed", range(5))
def test_per_pair_exchange_counters_survive_asymmetric_plans(seed):
    """Three ranks, blocks whose halos differ: exchange 0 involves only the pair (0, 1), exchange 1 only (1, 2), exchange 2
    everybody.  With ONE number per rank (round 1) rank 1 has counted two exchanges when rank 0 and rank 2 have counted one:
    in exchange 2 it waits for packets tagged 3 from peers that tag theirs 2 -- a dead wait (the 4 s timeout, then NaN
    ghosts).  With a counter per PAIR both ends of every pair have always counted the same number of exchanges."""
    plans = [
        {0: [1], 1: [0], 2: []},
        {0: [], 1: [2], 2: [1]},
        {0: [1], 1: [0, 2], 2: [1]},
        {0: [1], 1: [0], 2: []},
        {0: [1], 1: [0, 2], 2: [1]},
    ]
    for use_pair, expect_ok in ((True, True), (False, False)):
        rng = random.Random(seed)
        wins = [Window(3) for _ in range(3)]
        log = []
        progs = [_pair_counter_program(r, 3, wins, plans, use_pair, log) for r in range(3)]
        live = list(range(3))
        while live:
            r = rng.choice(live)
            try:
                next(progs[r])
            except StopIteration:
                live.remove(r)
        timeouts = [e for e in log if e[-1] == "timeout"]
        if expect_ok:
            assert not timeouts and sorted(e[0] for e in log) == [0, 1, 2]
        else:
            assert timeouts
