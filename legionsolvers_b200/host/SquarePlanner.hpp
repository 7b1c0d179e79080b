// SquarePlanner.hpp -- the reference's SquarePlanner<ENTRY_T> (src/SquarePlanner.hpp): a square
// block system over one or more index spaces, vectors addressed by integer id (0 = SOL, 1 = RHS,
// 2.. = workspace), row-partitioned matrices registered per (domain, range) block with their kernel
// and ghost partitions derived ONCE (add_row_partitioned_matrix, :209-235).
//
// Same method names and meaning (copy / zero_fill / scal / axpy(1-3) / xpay(1-2) / dot / matvec).
// What is new is below the API: scalars stay on the device, the ghost partition becomes a halo plan
// (grouped ncclSend/ncclRecv with the ranks that own the ghost rows) and per-piece dot futures
// become device partials folded in colour order and all-reduced.  The `*_fused` methods are the
// B200 fast path the solvers use: each equals a fixed sequence of the reference's calls, with the
// same element-wise arithmetic, in fewer HBM passes.
#pragma once

#include <set>
#include <tuple>

#include "Matrices.hpp"

namespace LegionSolvers {

template <typename T>
class SquarePlanner {
    struct HaloMove {
        int peer;
        int64_t send_lo, send_n, recv_lo, recv_n;
    };
    struct Block {
        const AbstractMatrix<T> *matrix;
        size_t domain_index, range_index;
        IntervalPartition kernel_partition, ghost_partition;
        std::vector<HaloMove> halo;  // what to trade with each peer before a mat-vec
        std::vector<Scalar<T>> part_yw, part_yy;  // per-colour partial slots of the fused dots
        // peer-memory exchange (lsk_halo_move): ONE landing allocation per block and rank, holding the landing buffer of
        // every peer this rank trades with, in the order of `halo`; recv_off / send_off = byte offset of move i's landing
        // buffer inside this rank's / inside the peer's allocation
        DeviceBuffer<char> landing;
        Runtime::Exported landing_peers;
        std::vector<size_t> recv_off, send_off;
        bool landing_ready = false;
        // the same for the REVERSE exchange of a transposed mat-vec (counts swapped: what I send forward I receive there)
        DeviceBuffer<char> rev_landing;
        Runtime::Exported rev_landing_peers;
        std::vector<size_t> rev_recv_off, rev_send_off;
    };

    Runtime *rt;
    std::vector<std::shared_ptr<IndexPartition>> canonical_index_partitions;
    std::vector<PartitionedVector<T>> sol_vectors, rhs_vectors;
    std::vector<std::vector<PartitionedVector<T>>> workspace_vectors;
    std::vector<Block> row_partitioned_matrices;
    std::vector<std::pair<int64_t, int64_t>> space_need;  // ghost range every vector of a space must hold
    std::vector<std::vector<Scalar<T>>> piece_partials;   // [slot set][space-major local piece]
    uint64_t halo_bytes_per_matvec = 0;
    std::set<std::size_t> halo_fresh;  // vector ids whose ghost values are current on every rank
    void mark_dirty(std::size_t vec_idx) { halo_fresh.erase(vec_idx); }

    // the block's halo plan as lsk_halo_move's of vector v (send ranges and ghost regions are v's, landing buffers the block's)
    int fill_moves(const Block &b, const PartitionedVector<T> &v, lsk_halo_move *moves) const {
        int n = 0;
        for (size_t i = 0; i < b.halo.size(); ++i) {
            const HaloMove &m = b.halo[i];
            lsk_halo_move &o = moves[n++];
            o.peer = m.peer;
            o.reserved = 0;
            o.n = m.send_n;
            o.src = m.send_n > 0 ? reinterpret_cast<const double *>(v.ptr(m.send_lo)) : nullptr;
            o.ll_send = b.landing_peers.base[(size_t) m.peer] + b.send_off[i];
            o.recv_n = m.recv_n;
            o.recv_dst = m.recv_n > 0 ? reinterpret_cast<double *>(v.ptr(m.recv_lo)) : nullptr;
            o.ll_recv = b.landing.ptr + b.recv_off[i];
        }
        return n;
    }

    // REVERSE exchange (transposed mat-vec): the contributions this rank accumulated in its ghost elements of v travel to the
    // rows' owners, which add them to their boundary rows -- the forward plan with the two directions swapped
    void reduce_halo(const Block &b, const PartitionedVector<T> &v) {
        if (b.halo.empty()) return;
        if constexpr (std::is_same<T, double>::value) {
            if (!b.landing_ready || b.halo.size() > LSK_MAX_HALO_MOVES) {
                // NCCL: the peers' contributions land in a scratch buffer, then one axpy (alpha = 1) per peer adds them
                if (rt->capturing() || rt->replaying()) rt->fail(LSK_E_INVALID, "rmatvec over NCCL allocates scratch: not inside a trace");
                int64_t total = 0;
                for (const HaloMove &m : b.halo) total += m.send_n;
                DeviceBuffer<T> tmp(rt, (size_t) (total > 0 ? total : 1));
                rt->group_start();
                int64_t off = 0;
                for (const HaloMove &m : b.halo) {
                    if (m.recv_n > 0) rt->send(v.ptr(m.recv_lo), (size_t) m.recv_n * sizeof(T), m.peer);
                    if (m.send_n > 0) rt->recv(tmp.ptr + off, (size_t) m.send_n * sizeof(T), m.peer);
                    off += m.send_n;
                }
                rt->group_end();
                off = 0;
                for (const HaloMove &m : b.halo) {
                    if (m.send_n > 0) {
                        T *const none[4] = {nullptr, nullptr, nullptr, nullptr};
                        const T *x = tmp.ptr + off;
                        T *y = v.ptr(m.send_lo);
                        const int64_t cnt = m.send_n;
                        rt->enqueue("halo reduce (add)", [&] { return VectorKernels<T>::axpy(rt->ctx(), rt->stream(), cnt, 0, none, x, y); });
                    }
                    off += m.send_n;
                }
                rt->fence();  // tmp is released on return
                return;
            }
            lsk_halo_move moves[LSK_MAX_HALO_MOVES];
            int n = 0;
            for (size_t i = 0; i < b.halo.size(); ++i) {
                const HaloMove &m = b.halo[i];
                lsk_halo_move &o = moves[n++];
                o.peer = m.peer;
                o.reserved = 0;
                o.n = m.recv_n;  // my ghost elements owned by the peer
                o.src = m.recv_n > 0 ? reinterpret_cast<const double *>(v.ptr(m.recv_lo)) : nullptr;
                o.ll_send = b.rev_landing_peers.base[(size_t) m.peer] + b.rev_send_off[i];
                o.recv_n = m.send_n;  // my boundary rows the peer holds ghosts of
                o.recv_dst = m.send_n > 0 ? reinterpret_cast<double *>(v.ptr(m.send_lo)) : nullptr;
                o.ll_recv = b.rev_landing.ptr + b.rev_recv_off[i];
            }
            rt->halo_reduce_p2p(moves, n);
        } else {
            rt->fail(LSK_E_INVALID, "the reverse halo exchange is instantiated for fp64");
        }
    }

    void register_space(size_t idx, const PartitionedVector<T> &v) {
        if (canonical_index_partitions.size() > idx) {
            if (!canonical_index_partitions[idx]->same_as(v.partition())) rt->fail(LSK_E_INVALID, "vector partition differs from the canonical one");
        } else {
            if (canonical_index_partitions.size() != idx) rt->fail(LSK_E_INVALID, "add vectors space by space");
            canonical_index_partitions.push_back(v.get_index_partition());
            space_need.emplace_back(v.partition().own_lo(), v.partition().own_hi());
        }
    }

    // COPY the first partial into `out`, ADD the rest in order, then sum across ranks
    void fold(const std::vector<T *> &parts, const Scalar<T> &out) {
        T *o = out.ptr();
        if (parts.empty()) {
            rt->enqueue("fold", [&] { return VectorKernels<T>::fill(rt->ctx(), rt->stream(), 1, (T) 0, o); });
        }
        for (size_t i = 0; i < parts.size(); ++i) {
            const T *p = parts[i];
            if (i == 0) {
                if (p != o) rt->enqueue("fold", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), LSK_OP_COPY, p, nullptr, o); });
            } else {
                rt->enqueue("fold", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), LSK_OP_ADD, o, p, o); });
            }
        }
        if (rt->fused_collectives()) {  // the producing kernel's tail already summed across ranks
            if (parts.size() != 1) rt->fail(LSK_E_INVALID, "fused reductions need exactly one local piece per rank");
            return;
        }
        if constexpr (std::is_same<T, double>::value) rt->allreduce_sum(o, 1);
        else if (rt->nranks() > 1) rt->fail(LSK_E_INVALID, "multi-rank reductions are instantiated for fp64");
    }

    size_t total_local_pieces() const {
        size_t n = 0;
        for (const auto &p : canonical_index_partitions) n += (size_t) (p->end_color - p->first_color);
        return n;
    }
    // per-piece partial slots, allocated once (so that fused passes never allocate inside a trace);
    // when the whole job has a single local piece the output slot itself is used
    std::vector<T *> partial_slots(int set, const Scalar<T> &out) {
        const size_t n = total_local_pieces();
        if (n == 1) return {out.ptr()};
        if (piece_partials.size() <= (size_t) set) piece_partials.resize((size_t) set + 1);
        auto &v = piece_partials[(size_t) set];
        while (v.size() < n) v.emplace_back(rt);
        std::vector<T *> r;
        for (size_t i = 0; i < n; ++i) r.push_back(v[i].ptr());
        return r;
    }

    template <class F>
    void for_each_local_piece(F &&f) {  // f(space, colour, lo, n, flat_index)
        size_t flat = 0;
        for (size_t s = 0; s < canonical_index_partitions.size(); ++s) {
            const IndexPartition &p = *canonical_index_partitions[s];
            for (int c = p.first_color; c < p.end_color; ++c) f(s, c, p.lo[(size_t) c], p.piece_size(c), flat++);
        }
    }

    void exchange_halo(const Block &b, const PartitionedVector<T> &v) {
        if (b.halo.empty()) return;
        if constexpr (std::is_same<T, double>::value) {
            if (b.landing_ready && b.halo.size() <= LSK_MAX_HALO_MOVES) {
                // one kernel: my boundary leaves as packets for the peers' landing buffers, theirs are unpacked into my ghosts
                lsk_halo_move moves[LSK_MAX_HALO_MOVES];
                const int n = fill_moves(b, v, moves);
                rt->halo_exchange_p2p(moves, n);
                return;
            }
        }
        rt->group_start();
        for (const HaloMove &m : b.halo) {
            if (m.send_n > 0) rt->send(v.ptr(m.send_lo), (size_t) m.send_n * sizeof(T), m.peer);
            if (m.recv_n > 0) rt->recv(v.ptr(m.recv_lo), (size_t) m.recv_n * sizeof(T), m.peer);
        }
        rt->group_end();
    }

public:
    explicit SquarePlanner(Runtime *rt_) : rt(rt_) {}

    Runtime *get_runtime() const { return rt; }
    // a caller wrote vector `vec_idx` behind the planner's back (host upload): its ghost copies are stale
    void vector_written(std::size_t vec_idx) { mark_dirty(vec_idx); }

    std::size_t add_sol_vector(const PartitionedVector<T> &v) {
        if (!workspace_vectors.empty()) rt->fail(LSK_E_INVALID, "add vectors before allocate_workspace");
        const std::size_t idx = sol_vectors.size();
        register_space(idx, v);
        sol_vectors.push_back(v);
        return idx;
    }
    std::size_t add_rhs_vector(const PartitionedVector<T> &v) {
        if (!workspace_vectors.empty()) rt->fail(LSK_E_INVALID, "add vectors before allocate_workspace");
        const std::size_t idx = rhs_vectors.size();
        register_space(idx, v);
        rhs_vectors.push_back(v);
        return idx;
    }

    std::size_t get_num_spaces() const { return canonical_index_partitions.size(); }
    std::size_t get_num_blocks() const { return row_partitioned_matrices.size(); }
    const IndexPartition &get_partition(std::size_t space) const { return *canonical_index_partitions[space]; }
    const IntervalPartition &get_kernel_partition(std::size_t block) const { return row_partitioned_matrices[block].kernel_partition; }
    const IntervalPartition &get_ghost_partition(std::size_t block) const { return row_partitioned_matrices[block].ghost_partition; }
    uint64_t get_halo_bytes_per_matvec() const { return halo_bytes_per_matvec; }

    // allocate_workspace (:153-190): `num_vectors` more vectors per space, ids 2..
    void allocate_workspace(std::size_t num_vectors) {
        if (!workspace_vectors.empty()) rt->fail(LSK_E_INVALID, "workspace already allocated");
        if (sol_vectors.size() != get_num_spaces() || rhs_vectors.size() != get_num_spaces())
            rt->fail(LSK_E_INVALID, "every space needs a sol and a rhs vector");
        for (std::size_t j = 0; j < num_vectors; ++j) {
            workspace_vectors.emplace_back();
            for (std::size_t i = 0; i < get_num_spaces(); ++i) {
                workspace_vectors[j].emplace_back(rt, "workspace_" + std::to_string(j) + "_" + std::to_string(i),
                                                  canonical_index_partitions[i]);
                workspace_vectors[j][i].ensure_range(space_need[i].first, space_need[i].second);
            }
        }
        // one space, one piece per rank: every reducing kernel is launched identically on every rank, so
        // the all-reduces can ride in the kernels' tails and the halo push in the producing xpay
        rt->set_fused_collectives(rt->p2p() && get_num_spaces() == 1 && canonical_index_partitions[0]->pieces == rt->nranks());
    }

    // bring the ghost values of vector `vec_idx` up to date on every rank (stand-alone exchange)
    void refresh_halo(std::size_t vec_idx) {
        std::set<size_t> done;
        for (const Block &b : row_partitioned_matrices)
            if (done.insert(b.domain_index).second) exchange_halo(b, get_vector(vec_idx, b.domain_index));
        halo_fresh.insert(vec_idx);
    }

    // Deferred all-reduces of the fused CG step (lsk_ctx_defer_next_allreduce): possible when the reductions are fused into
    // the producing kernels, this rank holds one piece, and cg_direction will be the consumer of r.r
    bool can_defer_allreduce(std::size_t p, std::size_t r) {
        static const bool off = [] { const char *e = getenv("LSK_DEFER_AR"); return e && e[0] == '0'; }();  // developer A/B switch
        if (off || !rt->fused_collectives() || get_num_spaces() != 1 || total_local_pieces() != 1) return false;
        if constexpr (!std::is_same<T, double>::value) {
            return false;
        } else {
            const IndexPartition &part = *canonical_index_partitions[0];
            const int64_t lo = part.own_lo(), cnt = part.own_hi() - part.own_lo() + 1;
            return cnt > 0 && lsk_cg_direction_supported(cnt, get_vector(r, 0).ptr(lo), get_vector(p, 0).ptr(lo)) != 0 && can_fuse_matvec_dot();
        }
    }
    void defer_next_allreduce() {
        rt->enqueue("defer all-reduce", [&] { return lsk_ctx_defer_next_allreduce(rt->ctx()); });
    }

    // true when xpay_halo exchanges the halo itself (then a step that ends with it leaves the ghosts current)
    bool halo_push_is_fused() const {
        return std::is_same<T, double>::value && rt->fused_collectives() && row_partitioned_matrices.size() == 1 &&
               row_partitioned_matrices[0].landing_ready && row_partitioned_matrices[0].halo.size() <= 4;
    }

    // xpay (2 scalars) whose result is the source of the next mat-vec: when the collectives are fused, the
    // same kernel exchanges the halo of the result with the neighbours
    void xpay_halo(std::size_t dst, Scalar<T> numer, Scalar<T> denom, std::size_t src) {
        if constexpr (std::is_same<T, double>::value) {
            const Block *blk = row_partitioned_matrices.size() == 1 ? &row_partitioned_matrices[0] : nullptr;
            if (blk && halo_push_is_fused()) {
                PartitionedVector<T> &y = get_vector(dst, 0);
                const PartitionedVector<T> &x = get_vector(src, 0);
                const IndexPartition &p = *canonical_index_partitions[0];
                lsk_halo_move moves[4];
                const int n = fill_moves(*blk, y, moves);
                const int64_t lo = p.own_lo(), cnt = p.own_hi() - p.own_lo() + 1;
                const T *xs = x.ptr(lo);
                T *ys = y.ptr(lo);
                mark_dirty(dst);
                rt->enqueue("xpay_halo", [&] {
                    return lsk_xpay_halo_f64(rt->ctx(), rt->stream(), cnt, 2, numer.ptr(), denom.ptr(), nullptr, nullptr, xs, ys, moves, n);
                });
                halo_fresh.insert(dst);
                return;
            }
        }
        xpay(dst, numer, denom, src);
    }

    // CGSolver::step lines src/CGSolver.hpp:53-54 in ONE launch when this rank holds one piece of one space:
    // history.push_back(rr_new); P = fma(rr_new/rr_cur, P, R) [+ the halo exchange of P, as xpay_halo];
    // rr_cur <- rr_new.  false = not possible here (the caller issues xpay_halo + push_back instead).
    bool cg_direction(std::size_t p, const Scalar<T> &rr_new, const Scalar<T> &rr_cur, std::size_t r, const ScalarHistory &history) {
        if constexpr (!std::is_same<T, double>::value) {
            return false;
        } else {
            if (get_num_spaces() != 1 || total_local_pieces() != 1) return false;
            const IndexPartition &part = *canonical_index_partitions[0];
            const int64_t lo = part.own_lo(), cnt = part.own_hi() - part.own_lo() + 1;
            PartitionedVector<T> &y = get_vector(p, 0);
            const PartitionedVector<T> &x = get_vector(r, 0);
            if (cnt <= 0 || !lsk_cg_direction_supported(cnt, x.ptr(lo), y.ptr(lo))) return false;
            lsk_halo_move moves[4];
            int n = 0;
            const bool push = halo_push_is_fused();
            if (push) n = fill_moves(row_partitioned_matrices[0], y, moves);
            const T *xs = x.ptr(lo);
            T *ys = y.ptr(lo);
            mark_dirty(p);
            rt->enqueue("cg_direction", [&] {
                return lsk_cg_direction_f64(rt->ctx(), rt->stream(), cnt, rr_cur.ptr(), rr_new.ptr(), xs, ys, n > 0 ? moves : nullptr, n,
                                            history.data(), history.get_capacity(), history.count_ptr());
            });
            if (push) halo_fresh.insert(p);
            return true;
        }
    }

    // CGSolver::step lines src/CGSolver.hpp:50-54 in ONE launch (lsk_cg_tail_f64) when this rank holds one piece of one
    // space and its vectors are small enough to live in the L2 -- the slab of a multi-GPU run.  false = not possible here
    // (the caller issues cg_update + cg_direction).  Must be preceded by a mat-vec whose p.q may be a deferred reduction.
    bool cg_tail_ready(std::size_t sol, std::size_t r, std::size_t p, std::size_t q) {
        if constexpr (!std::is_same<T, double>::value) {
            return false;
        } else {
            if (get_num_spaces() != 1 || total_local_pieces() != 1) return false;
            // Several ranks: opt-in (LSK_CG_TAIL=multi).  It first lost there (2 GPUs: 94.7 vs 77.1 us per iteration) because its
            // unpacking polled for the neighbours' packets before they could have landed, which delays the landing itself;
            // since halo_unpack lets one thread per CTA look first it wins on 2 GPUs too (77.7 vs 79.2 us).  The 8-GPU
            // confirmation is outstanding (profiles/r02_multi_gpu.md); the kernel is exercised on two ranks by
            // tests/test_comm_loopback_gpu.py.
            if (rt->nranks() > 1) {
                static const bool multi = [] { const char *e = getenv("LSK_CG_TAIL"); return e && std::string(e) == "multi"; }();
                if (!multi || !halo_push_is_fused()) return false;
            }
            const IndexPartition &part = *canonical_index_partitions[0];
            const int64_t lo = part.own_lo(), cnt = part.own_hi() - part.own_lo() + 1;
            return cnt > 0 && lsk_cg_tail_supported(rt->ctx(), cnt, get_vector(p, 0).ptr(lo), get_vector(q, 0).ptr(lo), get_vector(sol, 0).ptr(lo),
                                                    get_vector(r, 0).ptr(lo)) != 0;
        }
    }
    void cg_tail(std::size_t sol, std::size_t r, std::size_t p, std::size_t q, const Scalar<T> &rr_cur, const Scalar<T> &pq, const Scalar<T> &rr_new,
                 const ScalarHistory &history) {
        if constexpr (std::is_same<T, double>::value) {
            const IndexPartition &part = *canonical_index_partitions[0];
            const int64_t lo = part.own_lo(), cnt = part.own_hi() - part.own_lo() + 1;
            PartitionedVector<T> &vp = get_vector(p, 0);
            lsk_halo_move moves[4];
            int n = 0;
            const bool push = halo_push_is_fused();
            if (push) n = fill_moves(row_partitioned_matrices[0], vp, moves);
            T *pp = vp.ptr(lo), *xx = get_vector(sol, 0).ptr(lo), *rr = get_vector(r, 0).ptr(lo);
            const T *qq = get_vector(q, 0).ptr(lo);
            mark_dirty(sol);
            mark_dirty(r);
            mark_dirty(p);
            rt->enqueue("cg_tail", [&] {
                return lsk_cg_tail_f64(rt->ctx(), rt->stream(), cnt, rr_cur.ptr(), pq.ptr(), rr_new.ptr(), pp, qq, xx, rr, n > 0 ? moves : nullptr, n,
                                       history.data(), history.get_capacity(), history.count_ptr());
            });
            if (push) halo_fresh.insert(p);
        }
    }

    // add_row_partitioned_matrix (:209-235): kernel partition from the range partition, ghost
    // partition from the kernel partition -- computed once, on the GPU -- plus the halo plan
    void add_row_partitioned_matrix(const AbstractMatrix<T> &matrix, std::size_t domain_index, std::size_t range_index) {
        if (domain_index >= get_num_spaces() || range_index >= get_num_spaces()) rt->fail(LSK_E_INVALID, "block index out of range");
        if (!workspace_vectors.empty()) rt->fail(LSK_E_INVALID, "register matrices before allocate_workspace");
        Block b;
        b.matrix = &matrix;
        b.domain_index = domain_index;
        b.range_index = range_index;
        const IndexPartition &range = *canonical_index_partitions[range_index];
        const IndexPartition &domain = *canonical_index_partitions[domain_index];
        b.kernel_partition = matrix.create_kernel_partition_from_range_partition(range);
        b.ghost_partition = matrix.create_domain_partition_from_kernel_partition(domain.volume, b.kernel_partition, range);
        for (int c = range.first_color; c < range.end_color; ++c) {
            b.part_yw.emplace_back(rt);
            b.part_yy.emplace_back(rt);
        }
        // this rank's ghost interval = bounding interval over its colours
        int64_t g_lo = INT64_MAX, g_hi = INT64_MIN;
        for (int c = range.first_color; c < range.end_color; ++c) {
            if (b.ghost_partition.hi[(size_t) c] < b.ghost_partition.lo[(size_t) c]) continue;
            g_lo = std::min(g_lo, b.ghost_partition.lo[(size_t) c]);
            g_hi = std::max(g_hi, b.ghost_partition.hi[(size_t) c]);
        }
        if (g_hi < g_lo) { g_lo = 0; g_hi = -1; }
        // everybody learns everybody's owned rows and ghost interval of the domain space
        const int R = rt->nranks();
        std::vector<int64_t> all((size_t) R * 4);
        {
            const int64_t mine[4] = {domain.own_lo(), domain.own_hi(), g_lo, g_hi};
            DeviceBuffer<int64_t> send(rt, 4), recv(rt, (size_t) R * 4);
            rt->check_cuda(cudaMemcpyAsync(send.ptr, mine, sizeof(mine), cudaMemcpyHostToDevice, rt->stream()), "halo plan H2D");
            rt->allgather_i64(send.ptr, recv.ptr, 4);
            rt->check_cuda(cudaMemcpyAsync(all.data(), recv.ptr, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, rt->stream()), "halo plan D2H");
            rt->fence();
        }
        // the plan itself is host arithmetic with a C entry point of its own (lsk_halo_plan): the multi-process CPU tests
        // drive exactly this code
        std::vector<int64_t> moves((size_t) R * 5);
        int nmoves = 0;
        if (lsk_halo_plan(rt->rank(), R, all.data(), moves.data(), &nmoves) != 0) rt->fail(LSK_E_INVALID, "lsk_halo_plan");
        for (int i = 0; i < nmoves; ++i) {
            HaloMove m;
            m.peer = (int) moves[(size_t) i * 5];
            m.send_lo = moves[(size_t) i * 5 + 1];
            m.send_n = moves[(size_t) i * 5 + 2];
            m.recv_lo = moves[(size_t) i * 5 + 3];
            m.recv_n = moves[(size_t) i * 5 + 4];
            b.halo.push_back(m);
            halo_bytes_per_matvec += (uint64_t) m.recv_n * sizeof(T);
        }
        // Peer-memory exchange: this rank's landing buffers (one per peer it trades with, in plan order) live in one
        // zeroed allocation, mapped into every peer.  Where MY packets land inside a peer's allocation follows from that
        // peer's own plan, which is the same host arithmetic on the same all-gathered ranges.  COLLECTIVE: every rank
        // exports, halo or not.
        if (std::is_same<T, double>::value && rt->p2p()) {
            // `what` = index of the count a landing buffer is sized for in a plan entry: 4 (recv_n) forward, 2 (send_n) reverse
            auto landing_layout = [&](int r, int what, std::vector<int64_t> &mv, int &nmv, std::vector<size_t> &off) -> size_t {
                mv.assign((size_t) R * 5, 0);
                if (lsk_halo_plan(r, R, all.data(), mv.data(), &nmv) != 0) rt->fail(LSK_E_INVALID, "lsk_halo_plan");
                size_t total = 0;
                off.clear();
                for (int i = 0; i < nmv; ++i) {
                    off.push_back(total);
                    total += (lsk_halo_landing_bytes(mv[(size_t) i * 5 + (size_t) what]) + 255) & ~size_t(255);
                }
                return total;
            };
            auto make_landing = [&](int what, DeviceBuffer<char> &buf, Runtime::Exported &peers, std::vector<size_t> &recv_off, std::vector<size_t> &send_off) {
                std::vector<int64_t> mv;
                std::vector<size_t> off;
                int nmv = 0;
                const size_t bytes = landing_layout(rt->rank(), what, mv, nmv, recv_off);
                buf = DeviceBuffer<char>(rt, bytes + 256);
                rt->check_cuda(cudaMemsetAsync(buf.ptr, 0, bytes + 256, rt->stream()), "landing buffers");
                send_off.assign(b.halo.size(), 0);
                for (size_t i = 0; i < b.halo.size(); ++i) {
                    landing_layout(b.halo[i].peer, what, mv, nmv, off);
                    const int64_t mine = what == 4 ? b.halo[i].send_n : b.halo[i].recv_n;  // what the peer sized its buffer for
                    bool found = false;
                    for (int j = 0; j < nmv; ++j)
                        if ((int) mv[(size_t) j * 5] == rt->rank()) {
                            if (mv[(size_t) j * 5 + (size_t) what] != mine) rt->fail(LSK_E_INVALID, "halo plans of two ranks disagree");
                            send_off[i] = off[(size_t) j];
                            found = true;
                        }
                    if (!found) rt->fail(LSK_E_INVALID, "halo plan: the peer does not list this rank");
                }
                rt->fence();  // the buffers are zero before any peer can learn their address
                peers = rt->export_allocation(buf.ptr, 0, 0);
            };
            make_landing(4, b.landing, b.landing_peers, b.recv_off, b.send_off);
            make_landing(2, b.rev_landing, b.rev_landing_peers, b.rev_recv_off, b.rev_send_off);
            b.landing_ready = true;
        }
        // every vector of the domain space must be able to hold the ghost interval
        if (g_hi >= g_lo) {
            space_need[domain_index].first = std::min(space_need[domain_index].first, g_lo);
            space_need[domain_index].second = std::max(space_need[domain_index].second, g_hi);
        }
        // (ensure_range is collective in multi-rank mode: called on every rank, empty range or not)
        sol_vectors[domain_index].ensure_range(g_lo, g_hi);
        rhs_vectors[domain_index].ensure_range(g_lo, g_hi);
        row_partitioned_matrices.push_back(std::move(b));
    }

    PartitionedVector<T> &get_vector(std::size_t vec_idx, std::size_t space_idx) {
        if (vec_idx == 0) return sol_vectors[space_idx];
        if (vec_idx == 1) return rhs_vectors[space_idx];
        if (vec_idx - 2 >= workspace_vectors.size()) rt->fail(LSK_E_INVALID, "vector id out of range");
        return workspace_vectors[vec_idx - 2][space_idx];
    }

    // ---- the reference's vector-id operations (:248-338) --------------------------------------------------
    void zero_fill(std::size_t vec_idx) {
        mark_dirty(vec_idx);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(vec_idx, i).zero_fill();
    }
    void copy(std::size_t dst, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i) = get_vector(src, i);
    }
    void scal(std::size_t dst, Scalar<T> alpha) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).scal(alpha);
    }
    void axpy(std::size_t dst, Scalar<T> alpha, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).axpy(alpha, get_vector(src, i));
    }
    void axpy(std::size_t dst, Scalar<T> numer, Scalar<T> denom, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).axpy(numer, denom, get_vector(src, i));
    }
    void axpy(std::size_t dst, Scalar<T> n1, Scalar<T> n2, Scalar<T> denom, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).axpy(n1, n2, denom, get_vector(src, i));
    }
    void xpay(std::size_t dst, Scalar<T> alpha, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).xpay(alpha, get_vector(src, i));
    }
    void xpay(std::size_t dst, Scalar<T> numer, Scalar<T> denom, std::size_t src) {
        mark_dirty(dst);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst, i).xpay(numer, denom, get_vector(src, i));
    }
    // dot (:331-338): per-space dots chained with AddScalarTask
    Scalar<T> dot(std::size_t v, std::size_t w) {
        Scalar<T> result = get_vector(v, 0).dot(get_vector(w, 0));
        for (std::size_t i = 1; i < get_num_spaces(); ++i) result = result + get_vector(v, i).dot(get_vector(w, i));
        return result;
    }
    // same value into an existing slot, one all-reduce for all spaces (no allocation: trace-safe)
    void dot_into(std::size_t v, std::size_t w, const Scalar<T> &out) {
        std::vector<T *> parts = partial_slots(0, out);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t flat) {
            const T *a = get_vector(v, s).ptr(lo), *b = get_vector(w, s).ptr(lo);
            T *o = parts[flat];
            rt->enqueue("dot", [&] { return VectorKernels<T>::dot(rt->ctx(), rt->stream(), n, a, b, o); });
        });
        fold(parts, out);
    }

    // matvec (:340-357): zero_fill(dst), then one mat-vec launch per registered block.  CSR blocks
    // overwrite their piece (beta = 0, like the reference's GPU variant), so the 8 B/row fill is
    // skipped for range spaces written by a CSR block; COO blocks accumulate (beta = 1).
    void matvec(std::size_t dst_idx, std::size_t src_idx) { matvec_impl(dst_idx, src_idx, nullptr, nullptr, 0); }

    // rmatvec: dst = A^T src over all registered blocks -- the transposed operator the reference reserves TaskIDs for
    // (CSRRmatvecTask / COORmatvecTask) and never implemented.  zero_fill(dst), then every block accumulates its piece's
    // contribution into the columns it references.  On several ranks the contributions to columns owned by OTHER ranks
    // accumulate in this rank's ghost elements of dst and travel back to their owners in a REVERSE halo exchange
    // (lsk_halo_reduce_f64: the forward plan with the directions swapped, received values are added).
    void rmatvec(std::size_t dst_idx, std::size_t src_idx) {
        mark_dirty(dst_idx);
        for (std::size_t i = 0; i < get_num_spaces(); ++i) get_vector(dst_idx, i).zero_fill();
        for (const Block &b : row_partitioned_matrices) {
            PartitionedVector<T> &dst = get_vector(dst_idx, b.domain_index);
            const bool remote = rt->nranks() > 1 && !b.halo.empty();
            if (remote) dst.zero_ghosts();
            b.matrix->rmatvec(dst, get_vector(src_idx, b.range_index), b.kernel_partition, b.ghost_partition);
            if (remote) reduce_halo(b, dst);
        }
    }

    // SOL-style update x += sum_j y[j] v_j in one pass per piece (GMRES's real update): y = m device doubles,
    // basis = the vector ids of v_0 .. v_{m-1}
    void prepare_basis_table(const std::vector<std::size_t> &basis, DeviceBuffer<const T *> &table) {
        const size_t m = basis.size(), pieces = total_local_pieces();
        if (rt->capturing() || rt->replaying()) rt->fail(LSK_E_INVALID, "the basis pointer table must be built outside a trace");
        std::vector<const T *> host(m * pieces);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t, size_t flat) {
            for (size_t j = 0; j < m; ++j) host[flat * m + j] = get_vector(basis[j], s).ptr(lo);
        });
        table = DeviceBuffer<const T *>(rt, m * pieces);
        rt->check_cuda(cudaMemcpyAsync(table.ptr, host.data(), sizeof(const T *) * host.size(), cudaMemcpyHostToDevice, rt->stream()), "basis table");
        rt->fence();
    }
    void multi_axpy(std::size_t dst, const T *y_dev, const std::vector<std::size_t> &basis, DeviceBuffer<const T *> &table) {
        static_assert(std::is_same<T, double>::value, "fused passes are instantiated for fp64");
        mark_dirty(dst);
        const size_t m = basis.size();
        if (table.count != m * total_local_pieces()) rt->fail(LSK_E_INVALID, "multi_axpy: prepare_basis_table first");
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t flat) {
            T *x = get_vector(dst, s).ptr(lo);
            const T *const *V = table.ptr + flat * m;
            rt->enqueue("multi_axpy", [&] { return lsk_multi_axpy_f64(rt->ctx(), rt->stream(), n, (int) m, y_dev, V, x); });
        });
    }

    // ---- fused fast path (fp64) --------------------------------------------------------------------------------
    bool can_fuse_matvec_dot() const {
        std::set<size_t> seen;
        for (const Block &b : row_partitioned_matrices) {
            if (!b.matrix->overwrites_output() || !seen.insert(b.range_index).second) return false;
        }
        return seen.size() == get_num_spaces();
    }
    // dst = A src and *yw = dst . vec(w_idx) [and *yy = dst . dst] in the same pass
    void matvec_dot(std::size_t dst_idx, std::size_t src_idx, std::size_t w_idx, const Scalar<T> &yw, const Scalar<T> *yy = nullptr) {
        if (!can_fuse_matvec_dot()) {
            matvec(dst_idx, src_idx);
            dot_into(dst_idx, w_idx, yw);
            if (yy) dot_into(dst_idx, dst_idx, *yy);
            return;
        }
        matvec_impl(dst_idx, src_idx, &yw, yy, w_idx);
    }

    // CGSolver::step lines src/CGSolver.hpp:50-52 in one pass per piece
    void cg_update(std::size_t sol, std::size_t r, const Scalar<T> &rr_old, const Scalar<T> &pq, std::size_t p, std::size_t q,
                   const Scalar<T> &rr_new) {
        static_assert(std::is_same<T, double>::value, "fused passes are instantiated for fp64");
        mark_dirty(sol);
        mark_dirty(r);
        std::vector<T *> parts = partial_slots(0, rr_new);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t flat) {
            const T *pp = get_vector(p, s).ptr(lo), *qq = get_vector(q, s).ptr(lo);
            T *xx = get_vector(sol, s).ptr(lo), *rr = get_vector(r, s).ptr(lo), *o = parts[flat];
            rt->enqueue("cg_update", [&] { return lsk_cg_update_f64(rt->ctx(), rt->stream(), n, rr_old.ptr(), pq.ptr(), pp, qq, xx, rr, o); });
        });
        fold(parts, rr_new);
    }
    // dst = fma(alpha, src, dst); *out = dst . vec(w_idx)   (alpha folded from 1-4 terms like get_alpha)
    void axpy_dot(std::size_t dst, std::initializer_list<Scalar<T>> terms, std::size_t src, std::size_t w_idx, const Scalar<T> &out) {
        static_assert(std::is_same<T, double>::value, "fused passes are instantiated for fp64");
        T *f[4] = {nullptr, nullptr, nullptr, nullptr};
        int nt = 0;
        for (const Scalar<T> &t : terms) f[nt++] = t.ptr();
        mark_dirty(dst);
        std::vector<T *> parts = partial_slots(0, out);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t flat) {
            const T *xs = get_vector(src, s).ptr(lo), *ws = get_vector(w_idx, s).ptr(lo);
            T *ys = get_vector(dst, s).ptr(lo), *o = parts[flat];
            rt->enqueue("axpy_dot", [&] { return lsk_axpy_dot_f64(rt->ctx(), rt->stream(), n, nt, f[0], f[1], f[2], f[3], xs, ys, ws, o); });
        });
        fold(parts, out);
    }
    void bicg_p_update(std::size_t p, const Scalar<T> &rho_new, const Scalar<T> &rho_old, const Scalar<T> &alpha, const Scalar<T> &omega,
                       std::size_t v, std::size_t r) {
        static_assert(std::is_same<T, double>::value, "fused passes are instantiated for fp64");
        mark_dirty(p);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t) {
            const T *vv = get_vector(v, s).ptr(lo), *rr = get_vector(r, s).ptr(lo);
            T *pp = get_vector(p, s).ptr(lo);
            rt->enqueue("bicg_p_update", [&] {
                return lsk_bicg_p_update_f64(rt->ctx(), rt->stream(), n, rho_new.ptr(), rho_old.ptr(), alpha.ptr(), omega.ptr(), vv, rr, pp);
            });
        });
    }
    void bicg_tail(std::size_t sol, std::size_t r, const Scalar<T> &alpha, const Scalar<T> &ru, const Scalar<T> &uu, std::size_t p,
                   std::size_t u, std::size_t rt_idx, const Scalar<T> &rho_next) {
        static_assert(std::is_same<T, double>::value, "fused passes are instantiated for fp64");
        mark_dirty(sol);
        mark_dirty(r);
        std::vector<T *> parts = partial_slots(0, rho_next);
        for_each_local_piece([&](size_t s, int, int64_t lo, int64_t n, size_t flat) {
            const T *pp = get_vector(p, s).ptr(lo), *uu_v = get_vector(u, s).ptr(lo), *rtv = get_vector(rt_idx, s).ptr(lo);
            T *xx = get_vector(sol, s).ptr(lo), *rr = get_vector(r, s).ptr(lo), *o = parts[flat];
            rt->enqueue("bicg_tail", [&] {
                return lsk_bicg_tail_f64(rt->ctx(), rt->stream(), n, alpha.ptr(), ru.ptr(), uu.ptr(), pp, uu_v, rtv, xx, rr, o);
            });
        });
        fold(parts, rho_next);
    }

private:
    void matvec_impl(std::size_t dst_idx, std::size_t src_idx, const Scalar<T> *yw, const Scalar<T> *yy, std::size_t w_idx) {
        const size_t S = get_num_spaces();
        mark_dirty(dst_idx);
        const bool src_fresh = halo_fresh.count(src_idx) != 0;
        std::vector<bool> overwritten(S, false);
        for (const Block &b : row_partitioned_matrices)
            if (b.matrix->overwrites_output()) overwritten[b.range_index] = true;
        for (size_t s = 0; s < S; ++s)
            if (!overwritten[s]) get_vector(dst_idx, s).zero_fill();
        std::set<size_t> exchanged;
        std::vector<T *> parts_yw, parts_yy;
        std::vector<bool> csr_written(S, false);  // a range space's FIRST overwriting block overwrites, further ones accumulate
        auto run = [&](const Block &b) {
            const bool accumulate = b.matrix->overwrites_output() && csr_written[b.range_index];
            if (b.matrix->overwrites_output()) csr_written[b.range_index] = true;
            PartitionedVector<T> &src = get_vector(src_idx, b.domain_index);
            if (!src_fresh && exchanged.insert(b.domain_index).second) exchange_halo(b, src);
            if (yw) {
                const IndexPartition &range = *canonical_index_partitions[b.range_index];
                MatvecFusion<T> fz;
                fz.w = &get_vector(w_idx, b.range_index);
                fz.yw.assign((size_t) range.pieces, nullptr);
                if (yy) fz.yy.assign((size_t) range.pieces, nullptr);
                const bool single = (total_local_pieces() == 1);
                for (int c = range.first_color; c < range.end_color; ++c) {
                    const size_t i = (size_t) (c - range.first_color);
                    fz.yw[(size_t) c] = single ? yw->ptr() : b.part_yw[i].ptr();
                    parts_yw.push_back(fz.yw[(size_t) c]);
                    if (yy) {
                        fz.yy[(size_t) c] = single ? yy->ptr() : b.part_yy[i].ptr();
                        parts_yy.push_back(fz.yy[(size_t) c]);
                    }
                }
                b.matrix->matvec(get_vector(dst_idx, b.range_index), src, b.kernel_partition, b.ghost_partition, &fz);
            } else {
                b.matrix->matvec(get_vector(dst_idx, b.range_index), src, b.kernel_partition, b.ghost_partition, nullptr, accumulate);
            }
        };
        // overwriting (CSR) blocks first, accumulating (COO) blocks after
        for (const Block &b : row_partitioned_matrices)
            if (b.matrix->overwrites_output()) run(b);
        for (const Block &b : row_partitioned_matrices)
            if (!b.matrix->overwrites_output()) run(b);
        if (yw) fold(parts_yw, *yw);
        if (yy) fold(parts_yy, *yy);
    }
};

}  // namespace LegionSolvers
