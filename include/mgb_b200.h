/*
 * mgb_b200.h - C ABI of libmgb_b200.so: B200-native (sm_100a) Newton-step assembly for the
 * MultiGridBarrier(MPI).jl hot path.
 *
 * The reference has no FFI boundary of its own: it plugs into the solver through Julia multiple
 * dispatch (reference src/MultiGridBarrierMPI.jl:62-192).  Every entry point below therefore cites
 * the Julia method / upstream body whose work it replaces; the `ccall` stubs a maintainer would add
 * are in INTEGRATION.md and multigridbarriermpi.jl_b200/julia/MGBB200.jl.
 *
 * Conventions
 *   - plain C types only; all indices int32 (HPCSparseMatrix default Ti=Int32, src:260), values double;
 *   - CSR inputs are 0-based or 1-based (index_base), i.e. the local CSR view of an HPCSparseMatrix
 *     (rowptr/colval of src:216-221, a12 in SURVEY.md) can be passed zero-copy from Julia (base 1);
 *   - dense n x k matrices are column-major (Julia layout; column c is contiguous);
 *   - every function returns 0 on success, non-zero on error; mgb_last_error() gives the message of
 *     the last failing call on the calling thread.  Nothing here calls exit()/abort().
 *   - NaN/Inf in an iterate is data, not an error: it is reported through `all_finite` like
 *     amgb_all_isfinite (src:121-133).
 *   - *_dev pointers are device pointers on the context's GPU; *_host pointers are host memory.
 */
#ifndef MGB_B200_H
#define MGB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgb_ctx mgb_ctx;
typedef struct mgb_plan mgb_plan;

/* Local CSR block of a sparse operator (host memory). */
typedef struct {
    int64_t nrows, ncols, nnz;
    const int32_t* rowptr; /* nrows+1 */
    const int32_t* colidx; /* nnz */
    const double* vals;    /* nnz */
    int32_t index_base;    /* 0 (C) or 1 (Julia) */
} mgb_csr;

/* Pointwise convex set / barrier (upstream convex_Euclidian_power): on the Dz row y,
 * F = -log(s^(2/p) - |q|^2) - mu(p) log s with (q..., s) = y[idx] (0-based idx, last entry = s);
 * slack != 0 adds the last Dz column to s (feasibility phase). */
typedef struct {
    int32_t kind; /* MGB_BARRIER_EUCLIDIAN_POWER */
    int32_t nidx;
    int32_t idx[8];
    double p;
    int32_t slack;
    /* optional second cone (intersection of two convex sets, e.g. upstream parabolic_solve:
     * {s1 >= u^2} with idx2 = (u.id, s1.id), p2 = 2, next to {s2 >= |grad u|^p}); nidx2 = 0: none */
    int32_t nidx2;
    int32_t idx2[8];
    double p2;
} mgb_barrier;

#define MGB_BARRIER_EUCLIDIAN_POWER 1

/* what an assemble call computes */
#define MGB_WANT_F0 1   /* objective            (upstream f0)                                   */
#define MGB_WANT_GRAD 2 /* gradient  R' sum_k D_k' (w .* (F1_k + t c_k))       (upstream f1)    */
#define MGB_WANT_HESS 4 /* Hessian values on the fixed pattern, R' (sum D_j' diag D_k) R (f2)   */
#define MGB_STORE_DZ 8  /* also store Dz = Dz0 + (D R) s  (n x nD, apply_D)                     */

/* plan kinds reported by mgb_plan_info */
#define MGB_PATH_ELEMENT 1 /* fused element-block kernels (broken-element operators detected) */
#define MGB_PATH_CSR 2     /* general CSR kernels                                              */
/* OR-ed into force_path: build no Hessian pattern / replay lists (operator-only plan, e.g. Dz0 = D z) */
#define MGB_PLAN_NO_HESSIAN 16
/* OR-ed into force_path: accepted for compatibility (the two-stage element path - element_kernel +
 * gather_kernel - is the only element path) */
#define MGB_PLAN_TWO_STAGE 32

const char* mgb_last_error(void);
int mgb_version(void);

/* Context = one GPU + one stream (a cudaStream_t owned by the caller).  stream==NULL selects the
 * legacy default stream, which orders with the caller's other default-stream work.  Not re-entrant per ctx
 * (one Julia thread drives a rank: SURVEY.md 8b). */
int mgb_ctx_create(int device, void* stream, mgb_ctx** out);
int mgb_ctx_destroy(mgb_ctx* ctx);
int mgb_ctx_sync(mgb_ctx* ctx);

/* Symbolic phase, once per multigrid level.  Replaces the per-Newton-step structural work of the
 * reference: amgb_diag's spdiagm + structural hash (src:137-147, tools/profile_hash.jl:41-66), the
 * SpGEMM plans of D_j' * diag * D_k and R' * H * R (test/test_map_rows_compare.jl:102-123,165-171).
 *   D[k]  : nD operators, each n x N  (upstream D0[L,k] = hcat(Z.., op, ..Z), test/test_d0_construction.jl:91-100)
 *   R     : N x m  (upstream R_fine[l] = blockdiag(...), test/test_d0_construction.jl:81-83)
 *   x     : n x dim column-major nodes, w : n quadrature weights (host)
 *   row0,row1 : this rank's block of quadrature rows [row0,row1) (HPCSparseArrays row partition, a13);
 *               pass 0,n for a single rank.
 *   force_path: 0 = auto, MGB_PATH_ELEMENT / MGB_PATH_CSR to force one.
 *   ctx == NULL builds a symbolic-only plan (mgb_plan_info / mgb_plan_pattern work; every numeric
 *   entry point fails: there is no CPU numeric path). */
int mgb_plan_create(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R,
                    int32_t dim, const double* x_host, const double* w_host,
                    const mgb_barrier* barrier, int64_t row0, int64_t row1, int32_t force_path,
                    mgb_plan** out);
/* The same symbolic phase fed with this rank's HPCSparseMatrix blocks exactly as the reference stores them
 * (constructor argument order src/MultiGridBarrierMPI.jl:216-221; older field names
 * test/test_dump_matrices.jl:62-71): the local rows as CSC-of-the-transpose, i.e. `colptr` indexes the
 * LOCAL rows, `rowval` holds COMPRESSED local column ids and `col_indices[c]` is the global column of
 * compressed column c.  All four arrays can be passed zero-copy from Julia (index_base = 1).
 *   row0          : first global quadrature row of the block (0-based; row_partition[rank]-1, a13)
 *   x_local/w_local: the rank's rows of x and w (HPCMatrix.A / HPCVector.v local storage, src:176)
 *   R             : replicated N x m (every rank builds the same native geometry, src:239-240)
 * The plan is identical to mgb_plan_create(..., row0, row0 + nrows_local, ...) on the global operators. */
typedef struct {
    int64_t nrows_local, ncols_compressed, ncols_global, row0;
    const int32_t* colptr;      /* nrows_local + 1 */
    const int32_t* rowval;      /* nnz, compressed column ids */
    const double* nzval;        /* nnz */
    const int32_t* col_indices; /* ncols_compressed, global column ids */
    int32_t index_base;         /* 0 (C) or 1 (Julia) */
} mgb_hpc_block;
int mgb_plan_create_local(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_hpc_block* D, const mgb_csr* R,
                          int32_t dim, const double* x_local_host, const double* w_local_host,
                          const mgb_barrier* barrier, int32_t force_path, mgb_plan** out);
int mgb_plan_destroy(mgb_plan* plan);

/* sizes: info[0]=path, [1]=n_local, [2]=nD, [3]=m, [4]=nnzH, [5]=elements, [6]=nodes/element,
 * [7]=local cols/var, [8]=slots/element, [9]=Hessian contributions, [10]=gradient contributions,
 * [11]=plan device bytes, [12]=N, [13]=nu, [14]=algorithmic bytes per assembly (SURVEY 8d formula),
 * [15]=Hessian contributions stored incl. slice padding (CSR path) */
int mgb_plan_info(const mgb_plan* plan, int64_t* info, int32_t ninfo);

/* Fixed sparsity pattern of R' H R (CSR, 0-based, sorted columns) -> host buffers. */
int mgb_plan_pattern(const mgb_plan* plan, int32_t* rowptr_host, int32_t* colidx_host);

/* Numeric phase (per Newton step / line-search point); everything on the ctx stream, asynchronous.
 *   s_dev   : m      Newton unknown on this level (z = z0 + R s)
 *   Dz0_dev : n_local x nD column-major, D*z0 for the local rows (NULL = 0)
 *   c_dev   : n_local x nD column-major linear term (upstream c), scaled by t inside
 *   flags   : MGB_WANT_* | MGB_STORE_DZ
 * outputs (device; may be NULL when not requested):
 *   scal_dev: 4 doubles {f0, all_finite (1.0/0.0), <c,Dz>_w, reserved}
 *   grad_dev: m,  hval_dev: nnzH (same order as mgb_plan_pattern),  Dz_dev: n_local x nD.
 * Replaces upstream f0/f1/f2 (map_rows at src:161-170 + apply_D test/test_apply_d.jl:44 +
 * the triple-product loop test/test_map_rows_compare.jl:102-123 + R'..R :165-171). */
int mgb_assemble(mgb_plan* plan, const double* s_dev, const double* Dz0_dev, const double* c_dev,
                 double t, int32_t flags, double* scal_dev, double* grad_dev, double* hval_dev,
                 double* Dz_dev);

/* CUDA graphs.  mgb_assemble replays an instantiated graph of its two to four kernel launches (captured once per
 * distinct argument tuple, programmatic dependent-launch edges included; a small LRU cache per plan) whenever the
 * context runs on a real stream; MGB_GRAPH=0 in the environment keeps plain launches.  For a longer fixed sequence -
 * e.g. one assembly on every level of a hierarchy - bracket the calls with mgb_graph_begin / mgb_graph_end: between
 * the two, the library's launches on this context are recorded instead of executed; mgb_graph_launch replays them
 * with one launch (same buffers, same scalars). */
typedef struct mgb_graph mgb_graph;
int mgb_graph_begin(mgb_ctx* ctx);
int mgb_graph_end(mgb_ctx* ctx, mgb_graph** out);
int mgb_graph_launch(mgb_graph* graph);
int mgb_graph_destroy(mgb_graph* graph);
/* {state (1 graphs in use, 0 not yet decided, -1 off: MGB_GRAPH=0, legacy default stream or capture refused),
 *  graphs captured, graph launches} of a plan's mgb_assemble cache */
int mgb_graph_stats(const mgb_plan* plan, int64_t* stats3);

/* Same call through HOST buffers (what a CPU-array caller of f1/f2 sees): copies s (and, when
 * upload_inputs!=0, Dz0 and c) to the device, runs mgb_assemble, copies the requested results back
 * and synchronises. */
int mgb_assemble_host(mgb_plan* plan, const double* s_host, const double* Dz0_host,
                      const double* c_host, int32_t upload_inputs, double t, int32_t flags,
                      double* scal_host, double* grad_host, double* hval_host, double* Dz_host);

/* Separately callable pieces (the reference's own unit seams). */
/* apply_D: Dz = Dz0 + (D R) s   (test/test_apply_d.jl:44-49) */
int mgb_apply_D(mgb_plan* plan, const double* s_dev, const double* Dz0_dev, double* Dz_dev);
/* map_rows of the barrier over (x, Dz): which = 0 -> F (n), 1 -> F1 (n x nD), 2 -> F2 (n x nD^2,
 * column (j*nD+k), test/test_map_rows_compare.jl:62-73).  (src:161-170) */
int mgb_map_barrier(mgb_ctx* ctx, const mgb_barrier* barrier, int32_t nD, int64_t n, const double* Dz_dev,
                    int32_t which, double* out_dev);
/* amgb_all_isfinite (src:121-133): *flag_host = 1 if every entry finite. */
int mgb_all_isfinite(mgb_ctx* ctx, const double* v_dev, int64_t len, int32_t* flag_host);
/* Local part of the HPCVector reductions the Newton loop and the line search use (dot, sum, norm: reference
 * tools/profile_scaling.jl:89-109, tools/profile_barrier.jl:95-118): out = <x,y>, sum(x), |x|_2^2 or
 * max|x| over `len` device entries.  Deterministic (fixed grid, fixed fold order).  The result goes to out_dev
 * (1 double, may be NULL) and/or out_host (may be NULL; non-NULL synchronises the stream).  Across ranks the caller
 * all-reduces the scalar (MPI/NCCL), as HPCSparseArrays does. */
#define MGB_REDUCE_DOT 0
#define MGB_REDUCE_SUM 1
#define MGB_REDUCE_NORM2SQ 2
#define MGB_REDUCE_MAXABS 3
int mgb_reduce(mgb_ctx* ctx, int32_t op, const double* x_dev, const double* y_dev, int64_t len, double* out_dev,
               double* out_host);
/* amgb_diag (src:137-147) as a device map: out = w .* y[:,col] (the diagonal the reference wraps in a sparse matrix) */
int mgb_diag_scale(mgb_ctx* ctx, const double* w_dev, const double* y_dev, int64_t n, int64_t ld,
                   int32_t col, double* out_dev);

/* ---- small device-side sparse-matrix handle: the HPCSparseMatrix * HPCVector products the Newton
 * driver needs outside the fused assembly (z = z0 + R s: reference test/test_nonsquare.jl:45-55;
 * R' v: :57-72).  A' x uses a stored transpose (gather, no atomics). */
typedef struct mgb_spmat mgb_spmat;
int mgb_spmat_create(mgb_ctx* ctx, const mgb_csr* A, mgb_spmat** out);
int mgb_spmat_destroy(mgb_spmat* A);
/* y = alpha * op(A) x + beta * y0  (y0 may be NULL = 0; y may alias y0); trans != 0 -> op(A) = A' */
int mgb_spmat_mv(mgb_spmat* A, int32_t trans, double alpha, const double* x_dev, double beta,
                 const double* y0_dev, double* y_dev);

/* ---- index pack / unpack used by the multi-GPU interface exchange of Hessian rows and gradient
 * entries: out[k] = src[idx[k]]  and  dst[idx[k]] += src[k] (idx unique per call). */
int mgb_gather_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* idx_dev, int64_t count, double* out_dev);
int mgb_scatter_add_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* idx_dev, int64_t count, double* dst_dev);
/* owner-side reduction: dst[k] = sum of src[idx[r]] for r in [ptr[k], ptr[k+1]) in list order (deterministic) */
int mgb_segsum_idx(mgb_ctx* ctx, const double* src_dev, const int32_t* ptr_dev, const int32_t* idx_dev, int64_t nout,
                   double* dst_dev);

/* Page-lock caller-owned host arrays (e.g. the Julia Vector that receives nzval) once, so the D2H copies of
 * mgb_assemble_host run at PCIe DMA rate instead of the pageable-memory rate; undo before the array is freed. */
int mgb_host_register(void* ptr_host, int64_t bytes);
int mgb_host_unregister(void* ptr_host);

/* stream-ordered device -> host copy on the ctx stream, then synchronise (results that live in library-owned
 * device memory, e.g. the exchange window of mgb_dist_end) */
int mgb_copy_to_host(mgb_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);

/* ---- multi-GPU: one process per GPU, owner-computes sharding -----------------------------
 * Replaces, for the sharded assembly, the distributed SpGEMM / transpose exchanges HPCSparseArrays runs inside
 * D_j' * diag * D_k and R' * H * R (reference test/test_map_rows_compare.jl:102-123,165-171 on HPC types;
 * collectives listed in SURVEY.md 2.2).  The outputs - gradient entries and rows of R'HR - are split by
 * `out_part` (HPCSparseArrays row partition, a13; offsets 0-based, length nranks+1).  Every rank evaluates the
 * quadrature rows of ITS block of `row_part` (whole broken elements; the reference's partition of x / w) plus the
 * halo elements that touch one of its output rows, so every owned row is completed locally and no Hessian or
 * gradient value crosses NVLink; only the objective scalars are summed across ranks, by peer-memory words with
 * embedded epoch tags written from inside the gather kernel (no NCCL call, no fence, no host round trip).
 * Numbering of the unknowns: the library only sees "rank r owns the contiguous block out_part[r] .. out_part[r+1]".
 * With R = blockdiag(R_u, R_s) in the reference's stacked numbering such a block is all-u or all-s, and a rank ends up
 * evaluating every element that touches its u rows plus every element that touches its s rows (~2E/P, all of them
 * at P=2).  Callers should therefore permute the columns of R rank-major (block r of u, then block r of s, then block
 * r+1 of u ...: Python mgb_b200.dist.colocated_partition, Julia MGBB200.colocated_partition), pass s in that numbering
 * and read g and the rows / columns of R'HR in it: E/P elements plus a halo per rank.
 *
 *   1. mgb_dist_plan_create on every rank (same arguments except `rank`; ctx==NULL: symbolic only)
 *   2. mgb_dist_export -> exchange the 64-byte handles between processes (MPI/NCCL/any) -> mgb_dist_attach
 *      (or mgb_dist_attach_local with raw device pointers when all ranks live in one process)
 *   3. per Newton step, collectively and in the same order on every rank: mgb_dist_assemble
 *      (= mgb_dist_begin + mgb_dist_end).  Dz0 / c are n_local x nD blocks over the plan's rows (mgb_dist_rows
 *      order); s is the replicated Newton unknown (m).  Results stay valid until the next assemble call.
 * A peer that never delivers its scalars makes the call return NaN scalars with all_finite = 0 after
 * MGB_DIST_TIMEOUT_S seconds (default 30) and sets the error flag of mgb_dist_info - never a silent partial sum.
 */
typedef struct { unsigned char bytes[64]; } mgb_ipc_handle;

/* General form of mgb_plan_create: the plan covers the quadrature rows rows_sel[0..nrows_sel) (global 0-based ids,
 * whole broken elements, any order) and the output rows (unknowns) [out0, out1); only the first n_primary listed
 * rows count in the objective scalars.  mgb_plan_create(.., row0, row1, ..) = rows row0..row1-1, all outputs. */
int mgb_plan_create_rows(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R, int32_t dim,
                         const double* x_host, const double* w_host, const mgb_barrier* barrier, int64_t nrows_sel,
                         const int64_t* rows_sel, int64_t n_primary, int64_t out0, int64_t out1, int32_t force_path,
                         mgb_plan** out);
/* Fails with "sharded plans need the element path: ..." when the level cannot be sharded (operators without
 * broken-element structure, element type not instantiated, coarse level): assemble such a level redundantly. */
int mgb_dist_plan_create(mgb_ctx* ctx, int64_t n, int32_t nD, const mgb_csr* D, const mgb_csr* R,
                         int32_t dim, const double* x_host, const double* w_host,
                         const mgb_barrier* barrier, int32_t rank, int32_t nranks,
                         const int64_t* row_part, const int64_t* out_part, mgb_plan** out);
/* info[0]=rank, [1]=nranks, [2]=owned Hessian entries, [3]=owned unknowns, [4],[5]=owned unknown range,
 * [6]=local quadrature rows (primary block + halo), [7]=primary rows, [8]=local elements, [11]=window words,
 * [12]=epoch, [13]=error flag (1: a peer's scalars timed out) */
int mgb_dist_info(const mgb_plan* plan, int64_t* info, int32_t ninfo);
/* global ids of the plan's quadrature rows, in plan order (info[6] entries): row k of the Dz0 / c blocks */
int mgb_dist_rows(const mgb_plan* plan, int64_t* rows_host);
/* owned rows of the global pattern: rowptr (owned rows + 1, 0-based, relative), colidx (global ids) */
int mgb_dist_pattern(const mgb_plan* plan, int32_t* rowptr_host, int32_t* colidx_host);
int mgb_dist_window(mgb_plan* plan, void** window_dev, int64_t* bytes);
int mgb_dist_export(mgb_plan* plan, mgb_ipc_handle* handle);
int mgb_dist_attach(mgb_plan* plan, const mgb_ipc_handle* handles /* nranks */);
int mgb_dist_attach_local(mgb_plan* plan, void* const* windows_dev /* nranks */);
/* element kernel + gather kernel; the gather's scalar block publishes this rank's partial sums (asynchronous) */
int mgb_dist_begin(mgb_plan* plan, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t,
                   int32_t flags);
/* collects every rank's partial sums (small kernel).  Returns device pointers (library-owned): Hessian values of
 * the owned rows (mgb_dist_pattern order), owned gradient block, scalars {f0, all_finite, <c,Dz>_w, infeasible
 * count} (global sums, identical bits on every rank). */
int mgb_dist_end(mgb_plan* plan, double t, int32_t flags, const double** hval_own_dev,
                 const double** grad_own_dev, const double** scal_dev);
/* Row-distributed Newton unknown (the reference's s is an HPCVector: every rank holds only its block, the
 * [out_part[rank], out_part[rank+1]) entries).  mgb_dist_s_publish stores this rank's block into every rank's copy
 * of the whole vector over NVLink peer memory and raises this rank's epoch flag everywhere; mgb_dist_s_wait enqueues
 * the wait for all ranks' flags and returns the local copy (valid until the second-next publish); mgb_dist_assemble_s
 * = publish + wait + mgb_dist_assemble on the gathered vector.  This is the halo exchange of apply_D (R*s reads
 * entries of s that other ranks own), done once per assembly inside the library. */
int mgb_dist_s_publish(mgb_plan* plan, const double* s_own_dev);
int mgb_dist_s_wait(mgb_plan* plan, const double** s_full_dev);
int mgb_dist_assemble_s(mgb_plan* plan, const double* s_own_dev, const double* Dz0_dev, const double* c_dev, double t,
                        int32_t flags, const double** hval_own_dev, const double** grad_own_dev,
                        const double** scal_dev);
/* the same in one call; the gather kernel itself waits for the peers' words (one process per GPU only) */
int mgb_dist_assemble(mgb_plan* plan, const double* s_dev, const double* Dz0_dev, const double* c_dev, double t,
                      int32_t flags, const double** hval_own_dev, const double** grad_own_dev,
                      const double** scal_dev);

/* timing helper: runs `reps` assemblies back to back on the ctx stream, returns average ms measured
 * with CUDA events on that stream (bench.py uses it so the events sit on the launching stream). */
int mgb_time_assemble(mgb_plan* plan, const double* s_dev, const double* Dz0_dev, const double* c_dev,
                      double t, int32_t flags, double* scal_dev, double* grad_dev, double* hval_dev,
                      int32_t reps, int32_t flush_l2, float* ms_total, float* ms_kernel_element,
                      float* ms_kernel_gather);

/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t mgb_launch_count(void);
/* 1 while dependent kernels are launched as programmatic dependent launches (PDL); 0 after MGB_NO_PDL=1 or after the
 * driver refused the launch attribute once (the library then keeps plain launches for the rest of the process) */
int mgb_pdl_active(void);

#ifdef __cplusplus
}
#endif
#endif /* MGB_B200_H */
