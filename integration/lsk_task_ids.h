/*
 * lsk_task_ids.h -- the TaskID arithmetic of dzhang314/LegionSolvers, restated in plain C.
 *
 * A drop-in GPU variant is registered under the SAME TaskID as the reference's CPU body
 * (TaskTDI::preregister / TaskTDDDIII::preregister register `cuda_task_body` for Processor::TOC_PROC with the id
 * of `task_body`, src/TaskBaseClasses.hpp:211-223, 304-317), and the launchers compute that id at run time from the
 * index spaces' dimension and type tag (src/TaskBaseClasses.hpp:236-249, 330-365).  This header reproduces the
 * numbering so that a binding (or a test) can name the ids without a Legion build:
 *
 *   origin            LEGION_SOLVERS_TASK_ID_ORIGIN = 500000                       src/LibraryOptions.hpp:24-26
 *   meta tasks        LOAD_CUDA_LIBS = origin + 0; NUM_META_TASK_IDS = 1           src/TaskIDs.hpp:9-14
 *   block size        NUM_ENTRY_TYPES * MAX_DIM^3 * NUM_INDEX_TYPES^3 = 3*27*64    src/TaskBaseClasses.hpp:61-86
 *                     (the `void` list terminators are counted: entry types {float, double, void} -> 3,
 *                      index types {int, unsigned, long long, void} -> 4;           src/TaskBaseClasses.hpp:17-59)
 *   block order       enum TaskBlockID                                             src/TaskIDs.hpp:17-53
 *   within a block    TaskT:      + entry_index                                     src/TaskBaseClasses.hpp:88-103
 *                     TaskTDI:    + NUM_ENTRY*MAX_DIM*index_index + NUM_ENTRY*(dim-1) + entry_index      :196-209
 *                     TaskTDDDIII: mixed-radix over (N1,N2,N3,I1,I2,I3,entry)                            :262-285
 */
#ifndef LSK_TASK_IDS_H
#define LSK_TASK_IDS_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    LSK_TASK_ID_ORIGIN = 500000,
    LSK_NUM_META_TASK_IDS = 1,
    LSK_LOAD_CUDA_LIBS_TASK_ID = LSK_TASK_ID_ORIGIN + 0,
    LSK_NUM_ENTRY_TYPES = 3, /* float, double, void */
    LSK_NUM_INDEX_TYPES = 4, /* int, unsigned, long long, void */
    LSK_MAX_DIM_IDS = 3,
    LSK_TASK_BLOCK_SIZE = LSK_NUM_ENTRY_TYPES * 27 * 64 /* 5184 */
};

/* indices in the reference's type lists */
enum { LSK_ENTRY_F32 = 0, LSK_ENTRY_F64 = 1 };
enum { LSK_INDEX_S32 = 0, LSK_INDEX_U32 = 1, LSK_INDEX_S64 = 2 };

/* enum TaskBlockID, src/TaskIDs.hpp:17-53 (order matters) */
enum lsk_task_block {
    LSK_BLOCK_PRINT_SCALAR = 0,
    LSK_BLOCK_NEGATE_SCALAR,
    LSK_BLOCK_ADD_SCALAR,
    LSK_BLOCK_SUBTRACT_SCALAR,
    LSK_BLOCK_MULTIPLY_SCALAR,
    LSK_BLOCK_DIVIDE_SCALAR,
    LSK_BLOCK_SQRT_SCALAR,
    LSK_BLOCK_RSQRT_SCALAR,
    LSK_BLOCK_DUMMY,
    LSK_BLOCK_PRINT_INDEX,
    LSK_BLOCK_RANDOM_FILL,
    LSK_BLOCK_SCAL,
    LSK_BLOCK_AXPY,
    LSK_BLOCK_XPAY,
    LSK_BLOCK_DOT,
    LSK_BLOCK_COO_MATVEC,
    LSK_BLOCK_COO_RMATVEC,
    LSK_BLOCK_COO_PRINT,
    LSK_BLOCK_CSR_MATVEC,
    LSK_BLOCK_CSR_RMATVEC,
    LSK_BLOCK_CSR_PRINT,
    LSK_BLOCK_FILL_COO_NEGATIVE_LAPLACIAN,
    LSK_BLOCK_FILL_CSR_NEGATIVE_LAPLACIAN,
    LSK_BLOCK_FILL_CSR_NEGATIVE_LAPLACIAN_ROWPTR,
    LSK_BLOCK_FILL_COO_STENCIL,
    LSK_BLOCK_FILL_CSR_STENCIL,
    LSK_BLOCK_FILL_LINEARIZED_COO_STENCIL,
    LSK_BLOCK_FILL_LINEARIZED_CSR_STENCIL
};

static inline int lsk_task_block_base(int block) {
    return LSK_TASK_ID_ORIGIN + LSK_NUM_META_TASK_IDS + LSK_TASK_BLOCK_SIZE * block;
}
/* TaskT<BLOCK, Class, T>: the scalar tasks */
static inline int lsk_task_id_t(int block, int entry_index) { return lsk_task_block_base(block) + entry_index; }
/* TaskTDI<BLOCK, Class, T, N, I>: Scal / Axpy / Xpay / Dot on an N-dimensional index space with coordinate type I */
static inline int lsk_task_id_tdi(int block, int entry_index, int dim, int index_index) {
    return lsk_task_block_base(block) + LSK_NUM_ENTRY_TYPES * LSK_MAX_DIM_IDS * index_index + LSK_NUM_ENTRY_TYPES * (dim - 1) + entry_index;
}
/* TaskTDDDIII<BLOCK, Class, T, N1, N2, N3, I1, I2, I3>: the mat-vec tasks (kernel, domain, range spaces) */
static inline int lsk_task_id_tdddiii(int block, int entry_index, int n1, int n2, int n3, int i1, int i2, int i3) {
    const int I3 = LSK_NUM_INDEX_TYPES * LSK_NUM_INDEX_TYPES * LSK_NUM_INDEX_TYPES, I2 = LSK_NUM_INDEX_TYPES * LSK_NUM_INDEX_TYPES,
              I1 = LSK_NUM_INDEX_TYPES;
    return lsk_task_block_base(block) + I3 * LSK_NUM_ENTRY_TYPES * 9 * (n1 - 1) + I3 * LSK_NUM_ENTRY_TYPES * 3 * (n2 - 1) +
           I3 * LSK_NUM_ENTRY_TYPES * 1 * (n3 - 1) + I2 * LSK_NUM_ENTRY_TYPES * i1 + I1 * LSK_NUM_ENTRY_TYPES * i2 + 1 * LSK_NUM_ENTRY_TYPES * i3 +
           entry_index;
}

/* the ids of the hot path for the configuration every BASELINE config uses: fp64, 1-D spaces, long long coordinates */
#define LSK_TID_F64_1D_S64(block) lsk_task_id_tdi((block), LSK_ENTRY_F64, 1, LSK_INDEX_S64)
#define LSK_TID_MATVEC_F64_1D_S64(block) lsk_task_id_tdddiii((block), LSK_ENTRY_F64, 1, 1, 1, LSK_INDEX_S64, LSK_INDEX_S64, LSK_INDEX_S64)

#ifdef __cplusplus
}
#endif
#endif /* LSK_TASK_IDS_H */
