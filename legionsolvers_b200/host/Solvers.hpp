// Solvers.hpp -- CGSolver / BiCGStabSolver / GMRESSolver of the reference (src/CGSolver.hpp,
// src/BiCGStabSolver.hpp, src/GMRESSolver.hpp): the step recurrences, expressed on SquarePlanner
// vector ids.  Each solver has two equivalent bodies:
//   fused = false : the reference's sequence call for call (matvec, dot, axpy, ... one launch each);
//   fused = true  : the same arithmetic per element in the fewest HBM passes the dependency
//                   structure allows: CG 3 passes, BiCGStab 5 passes.
// Scalars live in fixed device slots allocated by the constructor, histories have their length on
// the device, so step() allocates nothing and can be recorded once (Runtime::begin_trace) and
// replayed as a CUDA graph -- the analogue of BenchmarkStencil's Legion traces.
#pragma once

#include "SquarePlanner.hpp"

namespace LegionSolvers {

template <typename T>
class CGSolver {
    static_assert(std::is_same<T, double>::value, "solvers are instantiated for fp64");
    static constexpr std::size_t SOL = 0, RHS = 1, P = 2, Q = 3, R = 4;

public:
    SquarePlanner<T> &planner;
    ScalarHistory residual_norm_squared;
    Scalar<T> negative_one;
    const bool fused;

private:
    Scalar<T> rr_cur, rr_new, p_norm;

public:
    // constructor (src/CGSolver.hpp:32-44): workspace(3); P <- RHS; R <- RHS (x0 = 0 assumed); rr0 = R.R
    explicit CGSolver(SquarePlanner<T> &planner_, bool fused_ = true, int64_t history_capacity = 1 << 16)
        : planner(planner_), residual_norm_squared(planner_.get_runtime(), history_capacity),
          negative_one(planner_.get_runtime(), static_cast<T>(-1)), fused(fused_), rr_cur(planner_.get_runtime()),
          rr_new(planner_.get_runtime()), p_norm(planner_.get_runtime()) {
        planner.allocate_workspace(3);
        reset();
    }

    // start a new solve with the current RHS (and SOL taken as 0, like the constructor)
    void reset() {
        residual_norm_squared.clear();
        planner.copy(P, RHS);
        planner.copy(R, RHS);
        planner.dot_into(R, R, rr_cur);
        residual_norm_squared.push_back(rr_cur);
        // when the xpay exchanges P's halo itself, every step ends with current ghosts; make it START so too,
        // so that the launch sequence of a step is the same from the first one on (traces are replayed)
        if (fused && planner.halo_push_is_fused()) planner.refresh_halo(P);
    }

    // step (src/CGSolver.hpp:46-55)
    void step() {
        if (fused) {
            // several ranks: the two dot products are only SENT by their producers; the next kernel -- their only
            // consumer -- forms the cross-rank sums at its start, so the NVLink flight overlaps the kernel boundary
            const bool defer = planner.can_defer_allreduce(P, R);
            if (defer) planner.defer_next_allreduce();
            planner.matvec_dot(Q, P, P, p_norm);                     // Q = A P and P.Q in one pass
            // an L2-resident slab (several GPUs): both vector passes, the r.r sum and the halo exchange of P in ONE launch
            if (planner.cg_tail_ready(SOL, R, P, Q)) {
                planner.cg_tail(SOL, R, P, Q, rr_cur, p_norm, rr_new, residual_norm_squared);
                return;
            }
            if (defer) planner.defer_next_allreduce();
            planner.cg_update(SOL, R, rr_cur, p_norm, P, Q, rr_new);  // both axpys and R.R in one pass
            // append rr_new, P = R + (rr_new/rr_cur) P (halo of P exchanged by the same kernel), rr_cur <- rr_new: one launch
            if (planner.cg_direction(P, rr_new, rr_cur, R, residual_norm_squared)) return;
            planner.xpay_halo(P, rr_new, rr_cur, R);                  // P's halo is exchanged by the same kernel
        } else {
            planner.matvec(Q, P);
            planner.dot_into(P, Q, p_norm);
            planner.axpy(SOL, rr_cur, p_norm, P);
            planner.axpy(R, negative_one, rr_cur, p_norm, Q);
            planner.dot_into(R, R, rr_new);
            planner.xpay(P, rr_new, rr_cur, R);
        }
        residual_norm_squared.push_back(rr_new, &rr_cur);  // append, and rr_cur <- rr_new for the next step
    }

    CGSolver(const CGSolver &) = delete;
    CGSolver &operator=(const CGSolver &) = delete;
};

template <typename T>
class BiCGStabSolver {
    static_assert(std::is_same<T, double>::value, "solvers are instantiated for fp64");
    static constexpr std::size_t SOL = 0, RHS = 1, P = 2, R = 3, R_TILDE = 4, U = 5, V = 6;

public:
    SquarePlanner<T> &planner;
    ScalarHistory rho, alpha, omega;
    Scalar<T> negative_one, zero, one;
    const bool fused;

private:
    Scalar<T> rho_cur, rho_next, alpha_cur, omega_cur, temp, ru, uu, beta, neg_omega, q1, q2;
    bool have_rho_next = false;

public:
    // constructor (src/BiCGStabSolver.hpp:36-60)
    explicit BiCGStabSolver(SquarePlanner<T> &planner_, bool fused_ = true, int64_t history_capacity = 1 << 16)
        : planner(planner_), rho(planner_.get_runtime(), history_capacity), alpha(planner_.get_runtime(), history_capacity),
          omega(planner_.get_runtime(), history_capacity), negative_one(planner_.get_runtime(), static_cast<T>(-1)),
          zero(planner_.get_runtime(), static_cast<T>(0)), one(planner_.get_runtime(), static_cast<T>(1)), fused(fused_),
          rho_cur(planner_.get_runtime(), static_cast<T>(1)), rho_next(planner_.get_runtime()),
          alpha_cur(planner_.get_runtime(), static_cast<T>(0)), omega_cur(planner_.get_runtime(), static_cast<T>(1)),
          temp(planner_.get_runtime()), ru(planner_.get_runtime()), uu(planner_.get_runtime()), beta(planner_.get_runtime()),
          neg_omega(planner_.get_runtime()), q1(planner_.get_runtime()), q2(planner_.get_runtime()) {
        planner.allocate_workspace(5);
        reset();
    }

    // start a new solve with the current RHS (and SOL taken as 0): the constructor's initialisation (:44-60)
    void reset() {
        rho.clear();
        alpha.clear();
        omega.clear();
        rho_cur.assign_value(one);
        alpha_cur.assign_value(zero);
        omega_cur.assign_value(one);
        planner.copy(R, RHS);
        planner.copy(R_TILDE, RHS);
        rho.push_back(one);
        alpha.push_back(zero);
        omega.push_back(one);
        planner.zero_fill(P);
        planner.zero_fill(V);
        if (fused) {  // the first step's rho = R.R~; later ones come out of the previous step's tail pass
            planner.dot_into(R, R_TILDE, rho_next);
            have_rho_next = true;
        }
    }

    // step (src/BiCGStabSolver.hpp:62-82)
    void step() {
        if (fused) {
            // P = beta (P - omega V) + R with beta = (rho_new/rho_old)(alpha/omega): one pass
            planner.bicg_p_update(P, rho_next, rho_cur, alpha_cur, omega_cur, V, R);
            rho.push_back(rho_next, &rho_cur);
            planner.matvec_dot(V, P, R_TILDE, temp);                 // V = A P and R~.V
            planner.axpy(R, negative_one, rho_cur, temp, V);         // s = r - (rho/temp) v
            alpha_cur.set(LSK_OP_DIV, rho_cur, temp);
            alpha.push_back(alpha_cur);
            planner.matvec_dot(U, R, R, ru, &uu);                    // U = A s, s.U and U.U
            omega_cur.set(LSK_OP_DIV, ru, uu);
            omega.push_back(omega_cur);
            // x += alpha p + omega s; r = s - omega u; next rho = r.R~: one pass
            planner.bicg_tail(SOL, R, alpha_cur, ru, uu, P, U, R_TILDE, rho_next);
        } else {
            planner.dot_into(R, R_TILDE, rho_next);
            q1.set(LSK_OP_DIV, rho_next, rho_cur);
            q2.set(LSK_OP_DIV, alpha_cur, omega_cur);
            beta.set(LSK_OP_MUL, q1, q2);
            rho.push_back(rho_next, &rho_cur);
            neg_omega.set(LSK_OP_NEG, omega_cur);
            planner.axpy(P, neg_omega, V);
            planner.xpay(P, beta, R);
            planner.matvec(V, P);
            planner.dot_into(R_TILDE, V, temp);
            planner.axpy(R, negative_one, rho_cur, temp, V);
            alpha_cur.set(LSK_OP_DIV, rho_cur, temp);
            alpha.push_back(alpha_cur);
            planner.matvec(U, R);
            planner.dot_into(R, U, ru);
            planner.dot_into(U, U, uu);
            omega_cur.set(LSK_OP_DIV, ru, uu);
            omega.push_back(omega_cur);
            planner.axpy(SOL, alpha_cur, P);
            planner.axpy(SOL, omega_cur, R);
            neg_omega.set(LSK_OP_NEG, omega_cur);
            planner.axpy(R, neg_omega, U);
        }
    }
};

template <typename T>
class GMRESSolver {
    static_assert(std::is_same<T, double>::value, "solvers are instantiated for fp64");
    static constexpr std::size_t SOL = 0, RHS = 1;

public:
    SquarePlanner<T> &planner;
    std::size_t restart;
    Scalar<T> negative_one, one;
    std::vector<std::vector<Scalar<T>>> inner_products;  // (restart + 1) x restart, as in the reference
    const bool fused;
    // false (default): the reference's placeholder update SOL += 1 * v_j (DummyTask), for parity with it.
    // true (set_real_update): the finished algorithm -- y = argmin || beta e1 - H y || by Givens rotations, SOL += V y in one pass.
    bool real_update = false;
    ScalarHistory residual_norm;  // real update only: || b - A x || after each cycle (the least-squares minimum)

private:
    Scalar<T> d, scale, neg_h, coeff;
    DeviceBuffer<T> hess;         // the inner_products table, contiguous: (restart + 1) x restart, row-major
    DeviceBuffer<T> ls;           // beta^2, y[0 .. restart), residual
    DeviceBuffer<const T *> basis_table;

public:
    // constructor (src/GMRESSolver.hpp:32-77)
    explicit GMRESSolver(SquarePlanner<T> &planner_, std::size_t restart_, bool fused_ = true)
        : planner(planner_), restart(restart_), negative_one(planner_.get_runtime(), static_cast<T>(-1)),
          one(planner_.get_runtime(), static_cast<T>(1)), fused(fused_), residual_norm(planner_.get_runtime(), 1 << 12),
          d(planner_.get_runtime()), scale(planner_.get_runtime()), neg_h(planner_.get_runtime()), coeff(planner_.get_runtime()),
          hess(planner_.get_runtime(), (restart_ + 1) * restart_), ls(planner_.get_runtime(), restart_ + 2) {
        Runtime *rt = planner.get_runtime();
        if (restart == 0 || restart > 64) rt->fail(LSK_E_INVALID, "GMRES restart must be 1 .. 64");
        planner.allocate_workspace(restart + 1);
        rt->check_cuda(cudaMemsetAsync(hess.ptr, 0, sizeof(T) * hess.count, rt->stream()), "hessenberg init");
        rt->check_cuda(cudaMemsetAsync(ls.ptr, 0, sizeof(T) * ls.count, rt->stream()), "least-squares init");
        for (std::size_t i = 0; i <= restart; ++i) {
            std::vector<Scalar<T>> row;
            for (std::size_t j = 0; j < restart; ++j) row.emplace_back(rt, hess.ptr + i * restart + j);  // views into the table
            inner_products.push_back(std::move(row));
        }
    }

    static constexpr std::size_t krylov_basis(std::size_t i) noexcept { return i + 2; }

    // step (src/GMRESSolver.hpp:83-127) = one restart cycle: residual, Arnoldi with modified Gram-Schmidt, then the
    // update (the reference's placeholder, or the real one).
    void step() {
        Runtime *rt = planner.get_runtime();
        planner.matvec(krylov_basis(0), SOL);
        planner.xpay(krylov_basis(0), negative_one, RHS);
        planner.dot_into(krylov_basis(0), krylov_basis(0), d);
        if (real_update) Scalar<T>(rt, ls.ptr).assign_value(d);  // beta^2 = || b - A x ||^2
        scale.set(LSK_OP_RSQRT, d);
        planner.scal(krylov_basis(0), scale);
        for (std::size_t j = 0; j < restart; ++j) {
            const std::size_t w = krylov_basis(j + 1);
            if (fused) {
                // h_0 rides on the mat-vec; each projection axpy carries the next inner product
                planner.matvec_dot(w, krylov_basis(j), krylov_basis(0), inner_products[0][j]);
                for (std::size_t k = 0; k <= j; ++k) {
                    const bool last = (k == j);
                    planner.axpy_dot(w, {negative_one, inner_products[k][j], one}, krylov_basis(k),
                                     last ? w : krylov_basis(k + 1), last ? d : inner_products[k + 1][j]);
                }
            } else {
                planner.matvec(w, krylov_basis(j));
                for (std::size_t k = 0; k <= j; ++k) {
                    planner.dot_into(krylov_basis(k), w, inner_products[k][j]);
                    neg_h.set(LSK_OP_NEG, inner_products[k][j]);
                    planner.axpy(w, neg_h, krylov_basis(k));
                }
                planner.dot_into(w, w, d);
            }
            inner_products[j + 1][j].set(LSK_OP_SQRT, d);
            if (j + 1 < restart) {
                scale.set(LSK_OP_RSQRT, d);
                planner.scal(w, scale);
            }
        }
        if (!real_update) {
            coeff.set(LSK_OP_DUMMY, one);
            for (std::size_t j = 0; j < restart; ++j) planner.axpy(SOL, coeff, krylov_basis(j));
            return;
        }
        // y = argmin || beta e1 - H y ||: one single-thread kernel; SOL += V y: one pass over the basis
        T *y = ls.ptr + 1, *res = ls.ptr + 1 + restart;
        const T *H = hess.ptr, *beta_sq = ls.ptr;
        const int m = (int) restart;
        rt->enqueue("gmres least squares", [&] { return lsk_gmres_solve_f64(rt->ctx(), rt->stream(), m, H, m, beta_sq, y, res); });
        planner.multi_axpy(SOL, y, basis_ids(), basis_table);
        residual_norm.push_back(Scalar<T>(rt, res));
    }

    // switch between the reference's placeholder update and the finished one (outside any trace)
    void set_real_update(bool on) {
        real_update = on;
        if (on) planner.prepare_basis_table(basis_ids(), basis_table);
    }

private:
    std::vector<std::size_t> basis_ids() const {
        std::vector<std::size_t> basis;
        for (std::size_t j = 0; j < restart; ++j) basis.push_back(krylov_basis(j));
        return basis;
    }
};

}  // namespace LegionSolvers
