// lsk_solvers.cpp -- C ABI over the C++ host layer (include/lsk_solvers.h).
#include "../../include/lsk_solvers.h"

#include <cstring>
#include <memory>
#include <string>

#include "Solvers.hpp"
#include "StencilGenerator.hpp"

using namespace LegionSolvers;

struct lsk_runtime {
    std::unique_ptr<Runtime> rt;
};
struct lsk_vector {
    Runtime *rt;
    std::unique_ptr<PartitionedVector<double>> v;
};
struct lsk_matrix {
    Runtime *rt;
    std::unique_ptr<CSRMatrix<double>> csr;
    std::unique_ptr<COOMatrix<double>> coo;
    const AbstractMatrix<double> *get() const {
        return csr ? static_cast<const AbstractMatrix<double> *>(csr.get()) : coo.get();
    }
};
struct lsk_planner {
    Runtime *rt;
    std::unique_ptr<SquarePlanner<double>> pl;
};
struct lsk_solver {
    Runtime *rt;
    int kind;
    std::unique_ptr<CGSolver<double>> cg;
    std::unique_ptr<BiCGStabSolver<double>> bicg;
    std::unique_ptr<GMRESSolver<double>> gmres;
};

static thread_local std::string g_last_error;

template <class F>
static int guard(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return LSK_E_INVALID;
    } catch (...) {
        g_last_error = "unknown C++ exception";
        return LSK_E_INVALID;
    }
}

#define REQUIRE(cond)                                               \
    do {                                                            \
        if (!(cond)) {                                              \
            g_last_error = std::string("null / bad argument: ") + #cond; \
            return LSK_E_INVALID;                                   \
        }                                                           \
    } while (0)

extern "C" {

const char *lsk_last_error(void) { return g_last_error.c_str(); }

// ---- halo plan (host arithmetic, no CUDA): what SquarePlanner::add_row_partitioned_matrix derives from the all-gathered
// {owned rows, ghost interval} of every rank -- for each peer, the sub-range of MY owned rows inside ITS ghost interval
// (send) and the sub-range of MY ghost interval that IT owns (recv) ------------------------------------------------------
int lsk_halo_plan(int rank, int nranks, const int64_t *ranges4, int64_t *moves5, int *nmoves) {
    if (rank < 0 || nranks < 1 || rank >= nranks || !ranges4 || !moves5 || !nmoves) return LSK_E_INVALID;
    const int64_t own_lo = ranges4[(size_t) rank * 4], own_hi = ranges4[(size_t) rank * 4 + 1];
    const int64_t g_lo = ranges4[(size_t) rank * 4 + 2], g_hi = ranges4[(size_t) rank * 4 + 3];
    int n = 0;
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) continue;
        const int64_t q_own_lo = ranges4[(size_t) q * 4], q_own_hi = ranges4[(size_t) q * 4 + 1];
        const int64_t q_g_lo = ranges4[(size_t) q * 4 + 2], q_g_hi = ranges4[(size_t) q * 4 + 3];
        const int64_t recv_lo = std::max(g_lo, q_own_lo);
        const int64_t recv_n = std::max<int64_t>(0, std::min(g_hi, q_own_hi) - recv_lo + 1);
        const int64_t send_lo = std::max(q_g_lo, own_lo);
        const int64_t send_n = std::max<int64_t>(0, std::min(q_g_hi, own_hi) - send_lo + 1);
        if (recv_n > 0 || send_n > 0) {
            int64_t *m = moves5 + (size_t) n * 5;
            m[0] = q; m[1] = send_lo; m[2] = send_n; m[3] = recv_lo; m[4] = recv_n;
            ++n;
        }
    }
    *nmoves = n;
    return 0;
}

// ---- runtime ---------------------------------------------------------------------------------------------------
int lsk_rt_create(int device, int rank, int nranks, void *external_stream, lsk_runtime **out) {
    REQUIRE(out);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_runtime>();
        h->rt = std::make_unique<Runtime>(device, rank, nranks, external_stream);
        *out = h.release();
    });
}
int lsk_rt_destroy(lsk_runtime *rt) {
    delete rt;
    return 0;
}
int lsk_rt_unique_id(void *out128) {
    REQUIRE(out128);
    return guard([&] { Runtime::comm_unique_id(out128); });
}
int lsk_rt_comm_init(lsk_runtime *rt, const void *uid128) {
    REQUIRE(rt && uid128);
    return guard([&] { rt->rt->comm_init(uid128); });
}
int lsk_rt_uses_peer_memory(lsk_runtime *rt) {
    if (!rt || !rt->rt->p2p()) return 0;
    return rt->rt->fused_collectives() ? 2 : 1;
}
int lsk_rt_comm_stats(lsk_runtime *rt, uint64_t *out4) {
    REQUIRE(rt && out4);
    return guard([&] {
        out4[0] = out4[1] = out4[2] = out4[3] = 0;
        if (!rt->rt->p2p()) return;
        const int rc = lsk_comm_stats(rt->rt->ctx(), rt->rt->stream(), &rt->rt->peers(), out4);
        if (rc != 0) rt->rt->fail(rc, "lsk_comm_stats");
    });
}
int lsk_rt_comm_error(lsk_runtime *rt, int *out) {
    REQUIRE(rt && out);
    return guard([&] { *out = rt->rt->comm_error(); });
}
lsk_ctx *lsk_rt_ctx(lsk_runtime *rt) { return rt ? rt->rt->ctx() : nullptr; }
void *lsk_rt_stream(lsk_runtime *rt) { return rt ? (void *) rt->rt->stream() : nullptr; }
int lsk_rt_fence(lsk_runtime *rt) {
    REQUIRE(rt);
    return guard([&] { rt->rt->fence(); });
}
uint64_t lsk_rt_kernel_launches(lsk_runtime *rt) { return rt ? rt->rt->kernel_launches() : 0; }
int lsk_rt_begin_trace(lsk_runtime *rt, int id) {
    REQUIRE(rt);
    return guard([&] { rt->rt->begin_trace(id); });
}
int lsk_rt_end_trace(lsk_runtime *rt, int id) {
    REQUIRE(rt);
    return guard([&] { rt->rt->end_trace(id); });
}

// ---- vectors -----------------------------------------------------------------------------------------------------
int lsk_vector_create(lsk_runtime *rt, const char *name, int64_t volume, int pieces, lsk_vector **out) {
    REQUIRE(rt && out && volume >= 0 && pieces > 0);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_vector>();
        h->rt = rt->rt.get();
        h->v = std::make_unique<PartitionedVector<double>>(h->rt, name ? name : "", IndexPartition::equal(*h->rt, volume, pieces));
        *out = h.release();
    });
}
int lsk_vector_destroy(lsk_vector *v) {
    delete v;
    return 0;
}
int lsk_vector_owned_range(lsk_vector *v, int64_t *lo, int64_t *hi) {
    REQUIRE(v && lo && hi);
    *lo = v->v->partition().own_lo();
    *hi = v->v->partition().own_hi();
    return 0;
}
int lsk_vector_constant_fill(lsk_vector *v, double value) {
    REQUIRE(v);
    return guard([&] { v->v->constant_fill(value); });
}
int lsk_vector_assign(lsk_vector *dst, const lsk_vector *src) {
    REQUIRE(dst && src);
    return guard([&] { *dst->v = *src->v; });
}
int lsk_vector_scal(lsk_vector *v, double alpha) {
    REQUIRE(v);
    return guard([&] { v->v->scal(Scalar<double>(v->rt, alpha)); });
}
int lsk_vector_axpy(lsk_vector *y, double alpha, const lsk_vector *x) {
    REQUIRE(y && x);
    return guard([&] { y->v->axpy(alpha, *x->v); });
}
int lsk_vector_xpay(lsk_vector *y, double alpha, const lsk_vector *x) {
    REQUIRE(y && x);
    return guard([&] { y->v->xpay(alpha, *x->v); });
}
int lsk_vector_dot(const lsk_vector *v, const lsk_vector *w, double *out) {
    REQUIRE(v && w && out);
    return guard([&] { *out = v->v->dot(*w->v).get_value(); });
}
int lsk_vector_copy_from_host(lsk_vector *v, const double *global) {
    REQUIRE(v && global);
    return guard([&] {
        v->v->copy_from_host(global);
        v->rt->fence();
    });
}
int lsk_vector_copy_to_host(const lsk_vector *v, double *global) {
    REQUIRE(v && global);
    return guard([&] {
        v->v->copy_to_host(global);
        v->rt->fence();
    });
}

// ---- matrices ----------------------------------------------------------------------------------------------------
int lsk_csr_create(lsk_runtime *rt, int64_t rows, int64_t cols, int64_t nnz_global, int64_t r_lo, int64_t r_hi, int64_t k_lo,
                   int64_t k_hi, const double *entry, const int64_t *col, const lsk_rect *rowptr, lsk_matrix **out) {
    REQUIRE(rt && out);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_matrix>();
        h->rt = rt->rt.get();
        h->csr = std::make_unique<CSRMatrix<double>>(h->rt, rows, cols, nnz_global, r_lo, r_hi, k_lo, k_hi, entry, col, rowptr);
        *out = h.release();
    });
}
int lsk_coo_create(lsk_runtime *rt, int64_t rows, int64_t cols, int64_t nnz_global, int64_t k_lo, int64_t k_hi,
                   const double *entry, const int64_t *row, const int64_t *col, lsk_matrix **out) {
    REQUIRE(rt && out);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_matrix>();
        h->rt = rt->rt.get();
        h->coo = std::make_unique<COOMatrix<double>>(h->rt, rows, cols, nnz_global, k_lo, k_hi, entry, row, col);
        *out = h.release();
    });
}
int lsk_csr_create_stencil(lsk_runtime *rt, const lsk_stencil *stencil, int pieces, lsk_matrix **out) {
    REQUIRE(rt && stencil && out && pieces > 0);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_matrix>();
        h->rt = rt->rt.get();
        h->csr = create_linearized_csr_stencil_matrix(h->rt, *stencil, (std::size_t) pieces);
        *out = h.release();
    });
}
int lsk_benchmark_stencil(int dim_flag, int64_t nx, int64_t ny, int64_t nz, lsk_stencil *out) {
    REQUIRE(out);
    return guard([&] { *out = benchmark_stencil(dim_flag, nx, ny, nz); });
}
int lsk_matrix_destroy(lsk_matrix *m) {
    delete m;
    return 0;
}
int lsk_matrix_info(const lsk_matrix *m, int64_t *o) {
    REQUIRE(m && o);
    if (m->csr) {
        const auto &c = *m->csr;
        const int64_t v[8] = {c.rows(), c.cols(), c.get_kernel_volume(), c.slab_rows_lo(), c.slab_rows_hi(),
                              c.slab_kernel_lo(), c.slab_kernel_hi(), 1};
        std::memcpy(o, v, sizeof(v));
    } else {
        const auto &c = *m->coo;
        const int64_t v[8] = {c.rows(), c.cols(), c.get_kernel_volume(), 0, c.rows() - 1, c.slab_kernel_lo(), c.slab_kernel_hi(), 0};
        std::memcpy(o, v, sizeof(v));
    }
    return 0;
}
int lsk_matrix_slab_to_host(const lsk_matrix *m, double *entry, int64_t *col, void *third) {
    REQUIRE(m && entry && col && third);
    return guard([&] {
        Runtime *rt = m->rt;
        if (m->csr) {
            const auto &c = *m->csr;
            const size_t nk = (size_t) std::max<int64_t>(0, c.slab_kernel_hi() - c.slab_kernel_lo() + 1);
            const size_t nr = (size_t) std::max<int64_t>(0, c.slab_rows_hi() - c.slab_rows_lo() + 1);
            rt->check_cuda(cudaMemcpyAsync(entry, c.entry_ptr(), nk * sizeof(double), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
            rt->check_cuda(cudaMemcpyAsync(col, c.col_ptr(), nk * sizeof(int64_t), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
            rt->check_cuda(cudaMemcpyAsync(third, c.rowptr_ptr(), nr * sizeof(lsk_rect), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
        } else {
            const auto &c = *m->coo;
            const size_t nk = (size_t) std::max<int64_t>(0, c.slab_kernel_hi() - c.slab_kernel_lo() + 1);
            rt->check_cuda(cudaMemcpyAsync(entry, c.entry_ptr(), nk * sizeof(double), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
            rt->check_cuda(cudaMemcpyAsync(col, c.col_ptr(), nk * sizeof(int64_t), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
            rt->check_cuda(cudaMemcpyAsync(third, c.row_ptr(), nk * sizeof(int64_t), cudaMemcpyDeviceToHost, rt->stream()), "D2H");
        }
        rt->fence();
    });
}
int lsk_matrix_device_fields(const lsk_matrix *m, void **entry, void **col, void **third) {
    REQUIRE(m && entry && col && third);
    if (m->csr) {
        *entry = m->csr->entry_ptr();
        *col = m->csr->col_ptr();
        *third = m->csr->rowptr_ptr();
    } else {
        *entry = m->coo->entry_ptr();
        *col = m->coo->col_ptr();
        *third = m->coo->row_ptr();
    }
    return 0;
}

// ---- planner -----------------------------------------------------------------------------------------------------
int lsk_planner_create(lsk_runtime *rt, lsk_planner **out) {
    REQUIRE(rt && out);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_planner>();
        h->rt = rt->rt.get();
        h->pl = std::make_unique<SquarePlanner<double>>(h->rt);
        *out = h.release();
    });
}
int lsk_planner_destroy(lsk_planner *pl) {
    if (!pl) return 0;
    delete pl;
    return 0;
}
int lsk_planner_add_sol_vector(lsk_planner *pl, lsk_vector *v) {
    REQUIRE(pl && v);
    return guard([&] { pl->pl->add_sol_vector(*v->v); });
}
int lsk_planner_add_rhs_vector(lsk_planner *pl, lsk_vector *v) {
    REQUIRE(pl && v);
    return guard([&] { pl->pl->add_rhs_vector(*v->v); });
}
int lsk_planner_add_row_partitioned_matrix(lsk_planner *pl, const lsk_matrix *m, int d, int r) {
    REQUIRE(pl && m && d >= 0 && r >= 0);
    return guard([&] { pl->pl->add_row_partitioned_matrix(*m->get(), (size_t) d, (size_t) r); });
}
int lsk_planner_allocate_workspace(lsk_planner *pl, int n) {
    REQUIRE(pl && n >= 0);
    return guard([&] { pl->pl->allocate_workspace((size_t) n); });
}
int lsk_planner_partition_bounds(lsk_planner *pl, int which, int index, int color, int64_t *lo, int64_t *hi) {
    REQUIRE(pl && lo && hi && index >= 0 && color >= 0);
    return guard([&] {
        auto &P = *pl->pl;
        if (which == 0) {
            if ((size_t) index >= P.get_num_spaces() || color >= P.get_partition((size_t) index).pieces) pl->rt->fail(LSK_E_INVALID, "partition_bounds");
            *lo = P.get_partition((size_t) index).lo[(size_t) color];
            *hi = P.get_partition((size_t) index).hi[(size_t) color];
        } else {
            if ((size_t) index >= P.get_num_blocks()) pl->rt->fail(LSK_E_INVALID, "partition_bounds");
            const IntervalPartition &ip = which == 1 ? P.get_kernel_partition((size_t) index) : P.get_ghost_partition((size_t) index);
            if ((size_t) color >= ip.lo.size()) pl->rt->fail(LSK_E_INVALID, "partition_bounds");
            *lo = ip.lo[(size_t) color];
            *hi = ip.hi[(size_t) color];
        }
    });
}
int lsk_planner_local_colors(lsk_planner *pl, int space, int *first, int *end) {
    REQUIRE(pl && first && end && space >= 0 && (size_t) space < pl->pl->get_num_spaces());
    *first = pl->pl->get_partition((size_t) space).first_color;
    *end = pl->pl->get_partition((size_t) space).end_color;
    return 0;
}
uint64_t lsk_planner_halo_bytes_per_matvec(lsk_planner *pl) { return pl ? pl->pl->get_halo_bytes_per_matvec() : 0; }
int lsk_planner_zero_fill(lsk_planner *pl, int v) {
    REQUIRE(pl && v >= 0);
    return guard([&] { pl->pl->zero_fill((size_t) v); });
}
int lsk_planner_copy(lsk_planner *pl, int dst, int src) {
    REQUIRE(pl && dst >= 0 && src >= 0);
    return guard([&] { pl->pl->copy((size_t) dst, (size_t) src); });
}
int lsk_planner_scal(lsk_planner *pl, int dst, double alpha) {
    REQUIRE(pl && dst >= 0);
    return guard([&] { pl->pl->scal((size_t) dst, Scalar<double>(pl->rt, alpha)); });
}
int lsk_planner_axpy(lsk_planner *pl, int dst, double alpha, int src) {
    REQUIRE(pl && dst >= 0 && src >= 0);
    return guard([&] { pl->pl->axpy((size_t) dst, Scalar<double>(pl->rt, alpha), (size_t) src); });
}
int lsk_planner_xpay(lsk_planner *pl, int dst, double alpha, int src) {
    REQUIRE(pl && dst >= 0 && src >= 0);
    return guard([&] { pl->pl->xpay((size_t) dst, Scalar<double>(pl->rt, alpha), (size_t) src); });
}
int lsk_planner_dot(lsk_planner *pl, int v, int w, double *out) {
    REQUIRE(pl && out && v >= 0 && w >= 0);
    return guard([&] { *out = pl->pl->dot((size_t) v, (size_t) w).get_value(); });
}
int lsk_planner_matvec(lsk_planner *pl, int dst, int src) {
    REQUIRE(pl && dst >= 0 && src >= 0);
    return guard([&] { pl->pl->matvec((size_t) dst, (size_t) src); });
}
int lsk_planner_rmatvec(lsk_planner *pl, int dst, int src) {
    REQUIRE(pl && dst >= 0 && src >= 0);
    return guard([&] { pl->pl->rmatvec((size_t) dst, (size_t) src); });
}
int lsk_planner_matvec_dot(lsk_planner *pl, int dst, int src, int w, double *out_yw, double *out_yy) {
    REQUIRE(pl && out_yw && dst >= 0 && src >= 0 && w >= 0);
    return guard([&] {
        Scalar<double> yw(pl->rt), yy(pl->rt);
        pl->pl->matvec_dot((size_t) dst, (size_t) src, (size_t) w, yw, out_yy ? &yy : nullptr);
        *out_yw = yw.get_value();
        if (out_yy) *out_yy = yy.get_value();
    });
}
int lsk_planner_vector_to_host(lsk_planner *pl, int vec, int space, double *global) {
    REQUIRE(pl && global && vec >= 0 && space >= 0);
    return guard([&] {
        pl->pl->get_vector((size_t) vec, (size_t) space).copy_to_host(global);
        pl->rt->fence();
    });
}
int lsk_planner_vector_from_host(lsk_planner *pl, int vec, int space, const double *global) {
    REQUIRE(pl && global && vec >= 0 && space >= 0);
    return guard([&] {
        pl->pl->vector_written((size_t) vec);  // its ghosts on the other ranks are stale from now on
        pl->pl->get_vector((size_t) vec, (size_t) space).copy_from_host(global);
        pl->rt->fence();
    });
}
int lsk_planner_vector_to_async(lsk_planner *pl, int vec, int space, double *global, void *stream) {
    REQUIRE(pl && global && vec >= 0 && space >= 0);
    return guard([&] {
        pl->pl->get_vector((size_t) vec, (size_t) space).copy_to_async(global, static_cast<cudaStream_t>(stream));
    });
}
int lsk_planner_vector_from_async(lsk_planner *pl, int vec, int space, const double *global, void *stream) {
    REQUIRE(pl && global && vec >= 0 && space >= 0);
    return guard([&] {
        pl->pl->vector_written((size_t) vec);
        pl->pl->get_vector((size_t) vec, (size_t) space).copy_from_async(global, static_cast<cudaStream_t>(stream));
    });
}

// ---- solvers -----------------------------------------------------------------------------------------------------
int lsk_solver_create(lsk_planner *pl, int kind, int restart, int fused, lsk_solver **out) {
    REQUIRE(pl && out);
    *out = nullptr;
    return guard([&] {
        auto h = std::make_unique<lsk_solver>();
        h->rt = pl->rt;
        h->kind = kind;
        if (kind == LSK_SOLVER_CG) h->cg = std::make_unique<CGSolver<double>>(*pl->pl, fused != 0, int64_t(1) << 16);
        else if (kind == LSK_SOLVER_BICGSTAB) h->bicg = std::make_unique<BiCGStabSolver<double>>(*pl->pl, fused != 0);
        else if (kind == LSK_SOLVER_GMRES) {
            if (restart <= 0) pl->rt->fail(LSK_E_INVALID, "GMRES restart");
            h->gmres = std::make_unique<GMRESSolver<double>>(*pl->pl, (size_t) restart, fused != 0);
        } else pl->rt->fail(LSK_E_INVALID, "solver kind");
        *out = h.release();
    });
}
int lsk_solver_destroy(lsk_solver *s) {
    if (!s) return 0;
    delete s;
    return 0;
}
int lsk_solver_step(lsk_solver *s) {
    REQUIRE(s);
    return guard([&] {
        if (s->cg) s->cg->step();
        else if (s->bicg) s->bicg->step();
        else s->gmres->step();
    });
}
int lsk_solver_reset(lsk_solver *s) {
    REQUIRE(s);
    return guard([&] {
        if (s->cg) s->cg->reset();
        else if (s->bicg) s->bicg->reset();
        // GMRESSolver keeps no state between restart cycles: a new solve starts from whatever SOL and RHS hold
    });
}
int lsk_solver_set_option(lsk_solver *s, int option, int value) {
    REQUIRE(s);
    return guard([&] {
        if (option == LSK_OPT_GMRES_REAL_UPDATE && s->gmres) s->gmres->set_real_update(value != 0);
        else s->rt->fail(LSK_E_INVALID, "unknown solver option");
    });
}
int lsk_solver_history_copy_async(lsk_solver *s, int which, double *dst, int64_t n, void *stream) {
    REQUIRE(s && dst && n >= 0);
    return guard([&] {
        const ScalarHistory *h = nullptr;
        if (s->cg) h = &s->cg->residual_norm_squared;
        else if (s->bicg) h = which == 0 ? &s->bicg->rho : which == 1 ? &s->bicg->alpha : &s->bicg->omega;
        else s->rt->fail(LSK_E_INVALID, "GMRESSolver keeps no history");
        if (n > h->get_capacity()) s->rt->fail(LSK_E_INVALID, "history_copy_async: more entries than the history holds");
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : s->rt->stream();
        if (n > 0) s->rt->check_cuda(cudaMemcpyAsync(dst, h->data(), sizeof(double) * (size_t) n, cudaMemcpyDefault, st), "history copy (async)");
    });
}
int lsk_solver_history(lsk_solver *s, int which, double *out, int64_t cap, int64_t *n) {
    REQUIRE(s && n);
    return guard([&] {
        std::vector<double> h;
        if (s->cg) {
            h = s->cg->residual_norm_squared.to_host();
        } else if (s->bicg) {
            h = which == 0 ? s->bicg->rho.to_host() : which == 1 ? s->bicg->alpha.to_host() : s->bicg->omega.to_host();
        } else if (which == 1) {
            h = s->gmres->residual_norm.to_host();
        } else {
            for (auto &row : s->gmres->inner_products)
                for (auto &sc : row) h.push_back(sc.get_value());
        }
        *n = (int64_t) h.size();
        if (out) std::memcpy(out, h.data(), sizeof(double) * (size_t) std::min<int64_t>(cap, *n));
    });
}

}  // extern "C"
