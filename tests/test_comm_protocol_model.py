"""CPU model check of the peer-memory protocols of the multi-GPU CG step (legionsolvers_b200/csrc/lsk_common.cuh
`allreduce_warp`, lsk_blas1.cu `cg_direction_tma_kernel`, lsk_cg.cu sync 3 + lsk_spmv_tma.cuh `GhostGate`).

The kernels cannot run without GPUs, but what makes them correct is a small protocol that can be executed exhaustively
enough on the CPU: every rank is a generator that yields before each remote store / poll, a seeded random scheduler
interleaves them, and assertions check the three properties the kernels rely on:

  1. all-reduce with TWO packet slots per source (by epoch parity): a poll never accepts a stale packet, and a packet
     is never overwritten before its reader has consumed it (a peer can run at most one reduction ahead);
  2. halo without a ready-handshake: a rank's direction update stores into its neighbours' ghost regions while those
     neighbours may be anywhere in their own step -- the all-reduce of p.Ap is what separates the store from the
     neighbour's reads of the previous ghost values;
  3. the deferred form (persistent kernel): nobody waits when the halo is published; the consumer waits, per ghost
     access, for the neighbour's epoch -- and still never reads a ghost of the wrong iteration.

The model mirrors the kernels' state (epochs, parity slots, `halo_done` flags) one to one; it is not product code.
"""
import random

import pytest


class Window:  # CommWindow of one rank
    def __init__(self, nranks):
        self.ar_pkt = [[(0, None)] * nranks for _ in range(2)]  # [parity][source] = (epoch, value)
        self.ar_consumed = [[0] * nranks for _ in range(2)]     # bookkeeping of the model: last epoch read per slot
        self.halo_done = [0] * nranks
        self.ghost_version = {}                                 # neighbour -> iteration of the values in my ghost region
        self.reading_ghosts = False                             # my mat-vec is in flight


def rank_program(me, nranks, wins, iters, deferred, log):
    """One rank's CG steps as a generator; yields where the GPU could be descheduled relative to its peers."""
    w = wins[me]
    nbrs = [r for r in (me - 1, me + 1) if 0 <= r < nranks]
    for nb in nbrs:
        w.ghost_version[nb] = 0
    ar_epoch = 0
    halo_epoch = 0

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
        # ---- mat-vec of iteration k: reads the ghosts of P_k
        if deferred and k > 0:
            for nb in nbrs:  # GhostGate: wait for the neighbour's epoch before the first ghost access
                while w.halo_done[nb] < halo_epoch:
                    yield
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
        # ---- direction update: P_{k+1}, boundary stored into the neighbours' ghost regions
        e = halo_epoch + 1
        for nb in nbrs:
            # property 2: no ready-handshake, yet the neighbour is never still reading the previous values
            assert not wins[nb].reading_ghosts or wins[nb].ghost_version[me] == k + 1, "store into a ghost region that is being read"
            assert wins[nb].ghost_version[me] == k, "halo of an iteration skipped or repeated"
            wins[nb].ghost_version[me] = k + 1
            yield
        for nb in nbrs:  # last CTA: publish the epoch
            wins[nb].halo_done[me] = e
            yield
        if not deferred:
            for nb in nbrs:  # ... and wait for the neighbours' (leaf kernels close the exchange here)
                while w.halo_done[nb] < e:
                    yield
        halo_epoch = e
        log.append((me, k))
    if deferred:  # the persistent kernel closes the last epoch before it exits
        for nb in nbrs:
            while w.halo_done[nb] < halo_epoch:
                yield


@pytest.mark.parametrize("deferred", [False, True], ids=["leaf kernels (wait at the halo close)", "persistent kernel (wait at the ghost access)"])
@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_cg_step_protocol_under_random_interleavings(nranks, deferred):
    for seed in range(40):
        rng = random.Random(1000 * nranks + seed)
        wins = [Window(nranks) for _ in range(nranks)]
        log = []
        progs = {r: rank_program(r, nranks, wins, iters=6, deferred=deferred, log=log) for r in range(nranks)}
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
            assert steps < 2_000_000, "no progress: the protocol dead-locked"
        assert sorted(log) == [(r, k) for r in range(nranks) for k in range(6)]
        # nobody ever ran more than one iteration ahead of a neighbour (the all-reduces are barriers)
        done = {}
        for r, k in log:
            done[r] = k
            for nb in (r - 1, r + 1):
                if 0 <= nb < nranks:
                    assert done.get(nb, -1) >= k - 1


def test_model_detects_a_broken_protocol():
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
        progs = {r: rank_program(r, nranks, wins, iters=4, deferred=False, log=[]) for r in range(nranks)}
        try:
            steps = 0
            while progs and steps < 100_000:
                r = rng.choice(list(progs))
                for _ in range(rng.choice((1, 2, 5, 40))):
                    try:
                        next(progs[r])
                    except StopIteration:
                        del progs[r]
                        break
                steps += 1
        except AssertionError:
            caught += 1
    assert caught > 0


# ---- exchanges are numbered PER PAIR of ranks (CommWindow::halo_sent) ------------------------------------------------------
def _pair_counter_program(me, nranks, wins_done, plans, use_pair_counters, log):
    """One rank issuing a sequence of halo exchanges; plans[x][me] = peers `me` trades with in exchange x (possibly none:
    the stand-alone exchange kernel is not even launched then).  Yields at every store / poll."""
    sent = [0] * nranks  # halo_sent[peer]
    single = 0           # the round-1 scheme: ONE epoch per rank, compared with per-peer flags
    for x, plan in enumerate(plans):
        peers = plan[me]
        if not peers:
            continue     # lsk_halo_exchange_f64 returns without a launch: no counter moves
        single += 1
        for p in peers:  # publish "exchange #e of our pair has landed"
            e = sent[p] + 1 if use_pair_counters else single
            wins_done[p][me] = e
            yield
        for p in peers:  # wait for the peer's
            e = sent[p] + 1 if use_pair_counters else single
            spins = 0
            while wins_done[me][p] < e:
                spins += 1
                if spins > 2000:
                    log.append((me, x, p, "timeout"))
                    return
                yield
            sent[p] = e
    log.append((me, "done"))


@pytest.mark.parametrize("seed", range(5))
def test_per_pair_exchange_counters_survive_asymmetric_plans(seed):
    """Three ranks, blocks whose halos differ: exchange 0 involves only the pair (0, 1), exchange 1 only (1, 2), exchange 2
    everybody.  With ONE epoch per rank (round 1) rank 1 has counted two exchanges when rank 0 and rank 2 have counted one:
    in exchange 2 it waits for epoch 3 from peers that publish 2 -- a dead wait (the 4 s timeout, then stale ghosts).  With
    a counter per PAIR both ends of every pair have always counted the same number of exchanges."""
    plans = [
        {0: [1], 1: [0], 2: []},
        {0: [], 1: [2], 2: [1]},
        {0: [1], 1: [0, 2], 2: [1]},
        {0: [1], 1: [0], 2: []},
        {0: [1], 1: [0, 2], 2: [1]},
    ]
    for use_pair, expect_ok in ((True, True), (False, False)):
        rng = random.Random(seed)
        wins_done = [[0] * 3 for _ in range(3)]
        log = []
        progs = [_pair_counter_program(r, 3, wins_done, plans, use_pair, log) for r in range(3)]
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
            assert timeouts, "the single-epoch scheme should have diverged on this plan"
