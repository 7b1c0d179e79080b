"""Prototypes of the C-ABI groups beyond the leaf kernels: the set-up kernels of include/lsk.h
(stencil generator, dependent partitioning) and the host-layer handles of include/lsk_solvers.h."""
from __future__ import annotations

import ctypes as C

i64, dbl, vp, ci, u64 = C.c_int64, C.c_double, C.c_void_p, C.c_int, C.c_uint64

LSK_MAX_DIM, LSK_MAX_STENCIL = 3, 64


class Stencil(C.Structure):
    """lsk_stencil (include/lsk.h)."""

    _fields_ = [
        ("dim", ci), ("order", ci), ("noff", ci),
        ("shape", i64 * LSK_MAX_DIM),
        ("offsets", (i64 * LSK_MAX_DIM) * LSK_MAX_STENCIL),
        ("values", dbl * LSK_MAX_STENCIL),
    ]


def declare(L) -> None:
    SP = C.POINTER(Stencil)
    L.lsk_stencil_sort.argtypes = [SP]
    L.lsk_stencil_size.argtypes = [SP]
    L.lsk_stencil_size.restype = i64
    L.lsk_stencil_count_f64.argtypes = [vp, vp, SP, i64, i64, vp]
    L.lsk_stencil_fill_csr_f64.argtypes = [vp, vp, SP, i64, i64, i64, vp, vp, vp, vp]
    L.lsk_csr_expand_rows.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    L.lsk_rect_span_i64.argtypes = [vp, vp, i64, vp, vp]
    L.lsk_image_range_flags.argtypes = [vp, vp, i64, vp, i64, i64, vp]
    L.lsk_minmax_i64.argtypes = [vp, vp, i64, vp, vp]
    L.lsk_image_flags.argtypes = [vp, vp, i64, vp, vp, i64, i64, vp]
    L.lsk_preimage_span_i64.argtypes = [vp, vp, i64, vp, i64, i64, i64, vp]
    L.lsk_preimage_flags.argtypes = [vp, vp, i64, vp, i64, i64, vp]
    L.lsk_preimage_range_flags.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp]
    L.lsk_equal_partition.argtypes = [i64, ci, vp, vp]
    L.lsk_shard.argtypes = [i64, i64, i64]

    L.lsk_comm_window_bytes.restype = C.c_size_t
    L.lsk_allreduce_sum_f64.argtypes = [vp, vp, vp, vp, ci]
    L.lsk_halo_exchange_f64.argtypes = [vp, vp, vp, vp, ci]
    L.lsk_halo_reduce_f64.argtypes = [vp, vp, vp, vp, ci]
    L.lsk_comm_error.argtypes = [vp, vp, vp, vp]
    L.lsk_ctx_set_peers.argtypes = [vp, vp]
    L.lsk_ctx_defer_next_allreduce.argtypes = [vp]
    L.lsk_ctx_settle.argtypes = [vp, vp]
    L.lsk_xpay_halo_f64.argtypes = [vp, vp, i64, ci, vp, vp, vp, vp, vp, vp, vp, ci]
    L.lsk_halo_plan.argtypes = [ci, ci, vp, vp, C.POINTER(ci)]
    L.lsk_mm_last_error.restype = C.c_char_p
    L.lsk_mm_read_info.argtypes = [C.c_char_p, vp]
    L.lsk_mm_read_coo_f64.argtypes = [C.c_char_p, i64, vp, vp, vp, C.POINTER(i64)]
    L.lsk_mm_write_coo_f64.argtypes = [C.c_char_p, i64, i64, i64, vp, vp, vp]
    L.lsk_coo_to_csr_f64.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp]
    L.lsk_last_error.restype = C.c_char_p
    L.lsk_rt_create.argtypes = [ci, ci, ci, vp, C.POINTER(vp)]
    L.lsk_rt_destroy.argtypes = [vp]
    L.lsk_rt_unique_id.argtypes = [vp]
    L.lsk_rt_comm_init.argtypes = [vp, vp]
    L.lsk_rt_uses_peer_memory.argtypes = [vp]
    L.lsk_rt_comm_error.argtypes = [vp, C.POINTER(ci)]
    L.lsk_rt_comm_stats.argtypes = [vp, vp]
    L.lsk_comm_stats.argtypes = [vp, vp, vp, vp]
    L.lsk_rt_ctx.argtypes = [vp]
    L.lsk_rt_ctx.restype = vp
    L.lsk_rt_stream.argtypes = [vp]
    L.lsk_rt_stream.restype = vp
    L.lsk_rt_fence.argtypes = [vp]
    L.lsk_rt_kernel_launches.argtypes = [vp]
    L.lsk_rt_kernel_launches.restype = u64
    L.lsk_rt_begin_trace.argtypes = [vp, ci]
    L.lsk_rt_end_trace.argtypes = [vp, ci]

    L.lsk_vector_create.argtypes = [vp, C.c_char_p, i64, ci, C.POINTER(vp)]
    L.lsk_vector_destroy.argtypes = [vp]
    L.lsk_vector_owned_range.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.lsk_vector_constant_fill.argtypes = [vp, dbl]
    L.lsk_vector_assign.argtypes = [vp, vp]
    L.lsk_vector_scal.argtypes = [vp, dbl]
    L.lsk_vector_axpy.argtypes = [vp, dbl, vp]
    L.lsk_vector_xpay.argtypes = [vp, dbl, vp]
    L.lsk_vector_dot.argtypes = [vp, vp, C.POINTER(dbl)]
    L.lsk_vector_copy_from_host.argtypes = [vp, vp]
    L.lsk_vector_copy_to_host.argtypes = [vp, vp]

    L.lsk_csr_create.argtypes = [vp, i64, i64, i64, i64, i64, i64, i64, vp, vp, vp, C.POINTER(vp)]
    L.lsk_coo_create.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp, vp, C.POINTER(vp)]
    L.lsk_csr_create_stencil.argtypes = [vp, SP, ci, C.POINTER(vp)]
    L.lsk_benchmark_stencil.argtypes = [ci, i64, i64, i64, SP]
    L.lsk_matrix_destroy.argtypes = [vp]
    L.lsk_matrix_info.argtypes = [vp, vp]
    L.lsk_matrix_slab_to_host.argtypes = [vp, vp, vp, vp]
    L.lsk_matrix_device_fields.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]

    L.lsk_planner_create.argtypes = [vp, C.POINTER(vp)]
    L.lsk_planner_destroy.argtypes = [vp]
    L.lsk_planner_add_sol_vector.argtypes = [vp, vp]
    L.lsk_planner_add_rhs_vector.argtypes = [vp, vp]
    L.lsk_planner_add_row_partitioned_matrix.argtypes = [vp, vp, ci, ci]
    L.lsk_planner_allocate_workspace.argtypes = [vp, ci]
    L.lsk_planner_partition_bounds.argtypes = [vp, ci, ci, ci, C.POINTER(i64), C.POINTER(i64)]
    L.lsk_planner_local_colors.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    L.lsk_planner_halo_bytes_per_matvec.argtypes = [vp]
    L.lsk_planner_halo_bytes_per_matvec.restype = u64
    L.lsk_planner_zero_fill.argtypes = [vp, ci]
    L.lsk_planner_copy.argtypes = [vp, ci, ci]
    L.lsk_planner_scal.argtypes = [vp, ci, dbl]
    L.lsk_planner_axpy.argtypes = [vp, ci, dbl, ci]
    L.lsk_planner_xpay.argtypes = [vp, ci, dbl, ci]
    L.lsk_planner_dot.argtypes = [vp, ci, ci, C.POINTER(dbl)]
    L.lsk_planner_matvec.argtypes = [vp, ci, ci]
    L.lsk_planner_matvec_dot.argtypes = [vp, ci, ci, ci, C.POINTER(dbl), C.POINTER(dbl)]
    L.lsk_planner_vector_to_host.argtypes = [vp, ci, ci, vp]
    L.lsk_planner_vector_from_host.argtypes = [vp, ci, ci, vp]
    L.lsk_planner_vector_to_async.argtypes = [vp, ci, ci, vp, vp]
    L.lsk_planner_vector_from_async.argtypes = [vp, ci, ci, vp, vp]

    L.lsk_solver_create.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
    L.lsk_solver_destroy.argtypes = [vp]
    L.lsk_solver_step.argtypes = [vp]
    L.lsk_solver_reset.argtypes = [vp]
    L.lsk_solver_history.argtypes = [vp, ci, vp, i64, C.POINTER(i64)]
    L.lsk_solver_history_copy_async.argtypes = [vp, ci, vp, i64, vp]
    L.lsk_solver_set_option.argtypes = [vp, ci, ci]
    L.lsk_planner_rmatvec.argtypes = [vp, ci, ci]
    L.lsk_csr_rspmv_f64.argtypes = [vp, vp, i64, i64, vp, vp, vp, i64, vp, vp, i64, i64]
    L.lsk_coo_rspmv_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, i64, i64, i64, i64]
    L.lsk_gmres_solve_f64.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp]
    L.lsk_multi_axpy_f64.argtypes = [vp, vp, i64, ci, vp, vp, vp]
