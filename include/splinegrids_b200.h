/*
 * splinegrids_b200.h -- C ABI of libsplinegrids_b200.so
 *
 * Hand-written sm_100a CUDA kernels for the grid-evaluation hot path of SplineGrids.jl
 * (per-dimension basis tables, evaluate!, evaluate_adjoint!, refinement-matrix application).
 * Every entry point replaces one KernelAbstractions kernel launch of the reference; the
 * kernel's argument list is the operator interface being mirrored (cited as path:line in the
 * reference checkout).  INTEGRATION.md shows the Julia `ccall` binding for each.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers owned by the caller, in Julia's column-major layout
 *    (first index fastest): eval (n_1..n_D, Nout), control points (c_1..c_D, Nout),
 *    basis tables (n, p+1, mdo+1).  Pointer/size ARRAYS (e.g. `tables`, `n_samples`) are HOST arrays
 *    of length nin.
 *  - Index arrays (sample_indices, row_pointer, column_start, refinement_indices) hold 1-BASED
 *    Int32 values exactly as the reference stores them in its user-visible fields.
 *  - `_f32` / `_f64` suffix = Float32 / Float64 (`Tv`).  `Ti` is Int32 (the reference's default,
 *    src/spline_dimension.jl:105); a host shim converts Int64 index arrays.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous on that stream; the reference synchronises after every launch
 *    (`synchronize(backend)`), which the host shim reproduces with one stream synchronize.
 *  - Return value: 0 = OK; >0 = cudaError_t; <0 = sg_status below.  Nothing throws, nothing
 *    allocates caller-visible memory.  The COMPUTE entry points keep no state between calls and may be
 *    called concurrently from several host threads on different streams (per-call stream, per-call
 *    workspace, per-handle plan).  Process-wide state exists only in the TEST / BENCHMARK hooks of the
 *    "library info" section below -- the launch counter (atomic), sg_set_kernel_policy, sg_last_variant and
 *    sg_profile_adjoint_main -- which are not meant for concurrent use, and in the lazily resolved NCCL
 *    function table (initialised once, thread-safe).
 *  - Every exported call opens an NVTX range named after the entry point (visible in Nsight Systems).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef SPLINEGRIDS_B200_H
#define SPLINEGRIDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_MAX_DIMS 8     /* maximum number of input dimensions Nin / array rank */
#define SG_MAX_DEGREE 15  /* maximum spline degree per dimension */

typedef enum {
    SG_OK = 0,
    SG_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, size <= 0, degree/derivative out of range */
    SG_ERR_UNSUPPORTED = -2,      /* nin > SG_MAX_DIMS, degree > SG_MAX_DEGREE */
    SG_ERR_WORKSPACE = -3,        /* workspace too small / misaligned */
    SG_ERR_NCCL = -4              /* NCCL unavailable or returned an error */
} sg_status;

/* ---- library info ---------------------------------------------------------------------- */
int sg_version(void);                       /* major*10000 + minor*100 + patch */
const char *sg_status_string(int status);   /* static string for a return value */
/* Number of kernels this library has launched since the last reset (process-wide, atomic). */
int64_t sg_launch_count(void);
void sg_launch_count_reset(void);
/* Force a kernel variant: 0 = automatic, 1 = generic kernels only, 2 = prefer tiled fast paths even
 * below the size threshold.  Test/benchmark hook (thread-unsafe, process-wide). */
void sg_set_kernel_policy(int policy);
/* Name of the kernel variant chosen by the last sg_evaluate / sg_evaluate_adjoint call. */
const char *sg_last_variant(void);
/* Benchmark hook: when enabled, sg_evaluate_adjoint records a CUDA event on the caller's stream right before and right
 * after its dominant kernel (the TMA-fed double march over the sample array); sg_profile_adjoint_main_ms waits for
 * the second event and returns the kernel's duration in milliseconds (< 0: nothing recorded).  Process-wide. */
void sg_profile_adjoint_main(int enable);
float sg_profile_adjoint_main_ms(void);

/* ---- K9 expand_knot_vector_kernel -- src/util_kernels.jl:1-20, launcher src/knot_vector.jl:29-37
 * knots_all[sum(mult)] <- knot_values[i] repeated multiplicities[i] times. */
int sg_expand_knot_vector_f32(float *knots_all, const float *knot_values, const int32_t *multiplicities,
                              int64_t n_values, void *stream);
int sg_expand_knot_vector_f64(double *knots_all, const double *knot_values, const int32_t *multiplicities,
                              int64_t n_values, void *stream);

/* ---- K1 set_sample_indices_kernel -- src/util_kernels.jl:22-49, launcher src/utils.jl:19-29
 * sample_indices[i] = clamp(#{k : !(t_i < knots_all[k]) for the leading run}, p+1, n_knots-p-1);
 * binary search instead of the reference's linear scan, bit-identical result. */
int sg_span_indices_f32(int32_t *sample_indices, const float *sample_points, int64_t n,
                        const float *knots_all, int64_t n_knots, int degree, void *stream);
int sg_span_indices_f64(int32_t *sample_indices, const double *sample_points, int64_t n,
                        const double *knots_all, int64_t n_knots, int degree, void *stream);

/* ---- K2 spline_dimension_kernel -- src/spline_dimension.jl:160-215, launcher :231-242
 * eval[n, p+1, mdo+1] <- non-zero basis values and derivatives (Cox-de Boor, in registers; the
 * reference's eval_prev scratch array is not needed).  Operation order follows the reference
 * and contraction is disabled, so tables are bit-identical to an IEEE evaluation of :169-214. */
int sg_basis_tables_f32(float *eval, const float *knots_all, int64_t n_knots, const float *sample_points,
                        const int32_t *sample_indices, int64_t n, int degree, int max_derivative_order,
                        void *stream);
int sg_basis_tables_f64(double *eval, const double *knots_all, int64_t n_knots, const double *sample_points,
                        const int32_t *sample_indices, int64_t n, int degree, int max_derivative_order,
                        void *stream);

/* ---- K1+K2 fused (one launch): what SplineDimension(...) does at src/spline_dimension.jl:153-154 */
int sg_dimension_build_f32(int32_t *sample_indices, float *eval, const float *knots_all, int64_t n_knots,
                           const float *sample_points, int64_t n, int degree, int max_derivative_order,
                           void *stream);
int sg_dimension_build_f64(int32_t *sample_indices, double *eval, const double *knots_all, int64_t n_knots,
                           const double *sample_points, int64_t n, int degree, int max_derivative_order,
                           void *stream);

/* ---- K10 decompress_basis_function_eval_kernel -- src/util_kernels.jl:51-67, src/spline_dimension.jl:251-272
 * out[n, n_basis] (caller zero-fills) <- eval[:, :, derivative_order+1] scattered to columns l-p..l. */
int sg_decompress_f32(float *out, const float *eval, const int32_t *sample_indices, int64_t n,
                      int64_t n_basis, int degree, int derivative_order, void *stream);
int sg_decompress_f64(double *out, const double *eval, const int32_t *sample_indices, int64_t n,
                      int64_t n_basis, int degree, int derivative_order, void *stream);

/* ---- K11 insert_kernel -- src/util_kernels.jl:69-79, src/utils.jl:58-64
 * out[len+1] <- v with x inserted at 1-based position i_insert. */
int sg_insert_f32(float *out, const float *v, int64_t len, int64_t i_insert, float x, void *stream);
int sg_insert_f64(double *out, const double *v, int64_t len, int64_t i_insert, double x, void *stream);
int sg_insert_i32(int32_t *out, const int32_t *v, int64_t len, int64_t i_insert, int32_t x, void *stream);

/* ---- K12 collect_indices_kernel -- src/util_kernels.jl:81-88, src/utils.jl:187-196
 * indices[n, nin] (column-major Int32) <- array of n CartesianIndex{nin} (nin Int64 each, AoS). */
int sg_collect_indices_i32(int32_t *indices, const int64_t *cartesian, int64_t n, int nin, void *stream);

/* ---- K3 spline_eval_kernel -- src/spline_grid.jl:119-183, launcher evaluate! :200-230
 * eval[J,o] = sum_I prod_d B_d[J_d, I_d, der_d+1] * cp[base(J)+I, o]  (optionally rational: weights != NULL
 * multiplies each term by w[base+I] and divides by the weighted sum).  eval need not be initialised.
 * tables[d]: (n_samples[d], degree[d]+1, mdo[d]+1);  indices[d]: n_samples[d] 1-based spans. */
int sg_evaluate_f32(float *eval, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                    const float *const *tables, const int32_t *const *indices, const int *degree,
                    const int *mdo, const int *der, const float *cp, const float *weights_or_null,
                    void *stream);
int sg_evaluate_f64(double *eval, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                    const double *const *tables, const int32_t *const *indices, const int *degree,
                    const int *mdo, const int *der, const double *cp, const double *weights_or_null,
                    void *stream);

/* ---- K3 for several derivative orders in one call (row f1 of the scope table): what the reference's callers do with
 * back-to-back evaluate! calls (docs/src/examples_optics.md:189-191: u, d1 u, d2 u; docs/src/examples_pde.md:69-72).
 * evals: HOST array of n_der device pointers, each (n_1..n_D, Nout); ders: HOST array n_der x nin (tuple q = ders[q*nin ..]).
 * 2-D grids of uniform degree 1..3 with 2..4 tuples run ONE fused kernel (control points and span bookkeeping shared by
 * the tuples); every other case runs one launch per tuple inside this call.  Results are those of n_der sg_evaluate calls. */
int sg_evaluate_multi_f32(float *const *evals, int n_der, const int *ders, int nin, const int64_t *n_samples,
                          const int64_t *n_cp, int nout, const float *const *tables, const int32_t *const *indices,
                          const int *degree, const int *mdo, const float *cp, const float *weights_or_null, void *stream);
int sg_evaluate_multi_f64(double *const *evals, int n_der, const int *ders, int nin, const int64_t *n_samples,
                          const int64_t *n_cp, int nout, const double *const *tables, const int32_t *const *indices,
                          const int *degree, const int *mdo, const double *cp, const double *weights_or_null, void *stream);

/* ---- K4 spline_eval_adjoint_kernel -- src/adjoint.jl:1-40, launcher evaluate_adjoint! :52-83
 * cp[i,o] = sum_{J : i in window(J)} prod_d B_d[..] * eval[J,o]; the callee zero-fills cp first
 * (src/adjoint.jl:61).  No global atomics when the per-dimension span indices are non-decreasing
 * (always true for the reference's constructor, src/spline_dimension.jl:121-128); otherwise an
 * atomic scatter kernel is used.  `weights_or_null` != NULL selects the transpose of the
 * fixed-weights rational map (an extension: the reference has no NURBS adjoint, src/adjoint.jl:52-57).
 * `workspace` may be NULL (the callee then uses cudaMallocAsync on `stream`), else at least
 * sg_evaluate_adjoint_workspace_bytes(...) bytes, 256-byte aligned (`rational` = weights will be passed). */
size_t sg_evaluate_adjoint_workspace_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                           const int *degree, int elem_size /* 4 or 8 */, int rational /* 0|1 */);
int sg_evaluate_adjoint_f32(float *cp, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                            const float *const *tables, const int32_t *const *indices, const int *degree,
                            const int *mdo, const int *der, const float *eval, const float *weights_or_null,
                            void *workspace, size_t workspace_bytes, void *stream);
int sg_evaluate_adjoint_f64(double *cp, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                            const double *const *tables, const int32_t *const *indices, const int *degree,
                            const int *mdo, const int *der, const double *eval, const double *weights_or_null,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- adjoint plan (the "plan/workspace handle for the adjoint's inverse sample map" of the operator interface) ----------
 * sg_evaluate_adjoint rebuilds the inverse sample map (first sample of every knot span, gather tables) on every call
 * and decides on device whether the span indices are monotone, so it must also launch its fallback kernels.  A fitting
 * loop calls evaluate_adjoint! thousands of times on the same SplineDimensions: a plan runs that preparation ONCE, keeps
 * the tables in device memory owned by the handle, reads the few decision flags back (this call synchronises `stream`)
 * and lets sg_evaluate_adjoint_planned launch exactly the kernels that are needed: for an eligible 3-D grid the fused
 * double march (all three contractions in one pass over the sample array + one halo-sum kernel, 2 launches).
 * The plan captures the ARGUMENTS of sg_evaluate_adjoint (sizes, table / index pointers, derivative_order); it is valid
 * as long as those device arrays are unchanged -- re-create it after evaluate!(spline_dimension) (src/spline_dimension.jl:231)
 * or when derivative_order changes.  Results are identical to sg_evaluate_adjoint up to summation order.
 * peer_stage_or_null != NULL: fused gradient push as in sg_evaluate_adjoint_push below.  multicast_stage_or_null != NULL: an
 * NVLS multicast address that maps the staging buffers of ALL ranks (CUDA multicast object / symmetric-memory multicast
 * pointer): the kernel then stores every finished control plane ONCE and the NVSwitch replicates it to every rank, so the
 * sender's NVLink egress is 1/world of the peer-pointer loop (which stays the fallback when no multicast mapping exists). */
typedef struct sg_adjoint_plan sg_adjoint_plan;
int sg_adjoint_plan_create_f32(sg_adjoint_plan **plan, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                               const float *const *tables, const int32_t *const *indices, const int *degree,
                               const int *mdo, const int *der, int rational, void *stream);
int sg_adjoint_plan_create_f64(sg_adjoint_plan **plan, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                               const double *const *tables, const int32_t *const *indices, const int *degree,
                               const int *mdo, const int *der, int rational, void *stream);
int sg_adjoint_plan_destroy(sg_adjoint_plan *plan);
/* what the plan found: span indices monotone in every dimension; column-block tables of the fused pipeline fit; largest
 * number of samples in one knot span of dimension 2 (any pointer may be NULL) */
int sg_adjoint_plan_info(const sg_adjoint_plan *plan, int *monotone, int *fused_tables_fit, int *rows2_max);
int sg_evaluate_adjoint_planned_f32(const sg_adjoint_plan *plan, float *cp, const float *eval, const float *weights_or_null,
                                    void *workspace, size_t workspace_bytes, void *const *peer_stage_or_null, int world,
                                    int my_rank, int64_t k0, int64_t np, int64_t max_planes, int keep_local,
                                    void *multicast_stage_or_null, void *stream);
int sg_evaluate_adjoint_planned_f64(const sg_adjoint_plan *plan, double *cp, const double *eval, const double *weights_or_null,
                                    void *workspace, size_t workspace_bytes, void *const *peer_stage_or_null, int world,
                                    int my_rank, int64_t k0, int64_t np, int64_t max_planes, int keep_local,
                                    void *multicast_stage_or_null, void *stream);

/* ---- K5 refinement_matrix_array_mul_kernel -- src/refinement_matrix.jl:365-403, mult! :421-445
 * (row window helpers src/refinement_matrix.jl:103-125, src/utils.jl:204-235)
 * Y[I] = sum_{J in window(I)} B[J] * prod_{d refined} A_d[I_d, J_d].  ndims = rank of Y and B (incl. Nout).
 * dims[r] (1-based) is the array dimension refined by matrix r; row_ptr[r]/col_start[r] have
 * sizeY[dims[r]-1] entries, nzval[r] has nnz[r]. */
int sg_refmat_mul_f32(float *Y, const float *B, int ndims, const int64_t *sizeY, const int64_t *sizeB, int n_ref,
                      const int *dims, const int32_t *const *row_ptr, const int32_t *const *col_start,
                      const float *const *nzval, const int64_t *nnz, void *stream);
int sg_refmat_mul_f64(double *Y, const double *B, int ndims, const int64_t *sizeY, const int64_t *sizeB, int n_ref,
                      const int *dims, const int32_t *const *row_ptr, const int32_t *const *col_start,
                      const double *const *nzval, const int64_t *nnz, void *stream);

/* ---- K6 refinement_matrix_array_mul_adjoint_kernel -- src/adjoint.jl:85-125, mult_adjoint! :127-152
 * B = transpose-apply of K5 to Y; the callee zero-fills B first (src/adjoint.jl:135). */
int sg_refmat_mul_adjoint_f32(float *B, const float *Y, int ndims, const int64_t *sizeY, const int64_t *sizeB,
                              int n_ref, const int *dims, const int32_t *const *row_ptr,
                              const int32_t *const *col_start, const float *const *nzval, const int64_t *nnz,
                              void *stream);
int sg_refmat_mul_adjoint_f64(double *B, const double *Y, int ndims, const int64_t *sizeY, const int64_t *sizeB,
                              int n_ref, const int *dims, const int32_t *const *row_ptr,
                              const int32_t *const *col_start, const double *const *nzval, const int64_t *nnz,
                              void *stream);

/* ---- K7 local_refinement_kernel -- src/control_points.jl:296-311 (loop :339-347)
 * cp[idx[i,:], o] = values[i, o];  idx (n_active, nin) and values (n_active, nout) column-major. */
int sg_scatter_active_f32(float *cp, int nin, const int64_t *n_cp, int nout, const int32_t *refinement_indices,
                          const float *refinement_values, int64_t n_active, void *stream);
int sg_scatter_active_f64(double *cp, int nin, const int64_t *n_cp, int nout, const int32_t *refinement_indices,
                          const double *refinement_values, int64_t n_active, void *stream);

/* ---- K8 local_refinement_adjoint_kernel -- src/adjoint.jl:154-170 (loop :185-193)
 * values[i, o] = cp[idx[i,:], o]; cp[idx[i,:], o] = 0. */
int sg_gather_zero_active_f32(float *refinement_values, float *cp, int nin, const int64_t *n_cp, int nout,
                              const int32_t *refinement_indices, int64_t n_active, void *stream);
int sg_gather_zero_active_f64(double *refinement_values, double *cp, int nin, const int64_t *n_cp, int nout,
                              const int32_t *refinement_indices, int64_t n_active, void *stream);

/* ---- construction side of local refinement on the device (scope row f3) --------------------------------------------
 * K13 build_refinement_matrix_kernel -- src/refinement.jl:3-36 (launcher :53-88): Boehm's single-knot insertion matrix,
 * (n_basis_old + 1) rows, n_basis_old + degree + 1 non-zeros; knot_span_index as returned by insert_knot (:114-131). */
int sg_boehm_matrix_f32(int32_t *row_pointer, int32_t *column_start, float *nzval, const float *knots_all_old,
                        int64_t n_basis_old, int64_t knot_span_index, float knot_new, int degree, void *stream);
int sg_boehm_matrix_f64(int32_t *row_pointer, int32_t *column_start, double *nzval, const double *knots_all_old,
                        int64_t n_basis_old, int64_t knot_span_index, double knot_new, int degree, void *stream);
/* K14 validate_refinement_matrix_kernel -- src/refinement_matrix.jl:134-181: valid_row[i] (uint8) per row. */
int sg_refmat_validate_i32(uint8_t *valid_row, const int32_t *row_pointer, const int32_t *column_start, int64_t m, int64_t nnz,
                           int64_t n_columns, void *stream);
/* K16 refinement_matrix_mul_nonzeros_kernel -- src/refinement_matrix.jl:229-271: per row of C = A * B the number of
 * non-zeros and the first column.  sg_row_pointer_from_counts: row_pointer = 1 + cumsum (m + 1 entries of scratch, the
 * last one is the total), total copied to the host (synchronises) -- :300-303.  K15 refinement_matrix_multiplication_kernel
 * -- :184-227: the values (nzval_C is zero-filled by the callee). */
int sg_refmat_mul_nonzeros_i32(int32_t *n_nonzero_C, int32_t *column_start_C, const int32_t *row_pointer_A, const int32_t *column_start_A,
                               int64_t m_A, int64_t nnz_A, const int32_t *row_pointer_B, const int32_t *column_start_B, int64_t m_B,
                               int64_t nnz_B, int64_t n_columns_B, void *stream);
int sg_row_pointer_from_counts(int32_t *row_pointer, int64_t *total_host, const int32_t *counts, int64_t m, void *stream);
int sg_refmat_mul_values_f32(float *nzval_C, int64_t nnz_C, const int32_t *row_pointer_C, const int32_t *column_start_C,
                             const int32_t *row_pointer_A, const int32_t *column_start_A, const float *nzval_A, int64_t m_A, int64_t nnz_A,
                             const int32_t *row_pointer_B, const int32_t *column_start_B, const float *nzval_B, int64_t m_B, int64_t nnz_B,
                             void *stream);
int sg_refmat_mul_values_f64(double *nzval_C, int64_t nnz_C, const int32_t *row_pointer_C, const int32_t *column_start_C,
                             const int32_t *row_pointer_A, const int32_t *column_start_A, const double *nzval_A, int64_t m_A, int64_t nnz_A,
                             const int32_t *row_pointer_B, const int32_t *column_start_B, const double *nzval_B, int64_t m_B, int64_t nnz_B,
                             void *stream);
/* K17 collect_refinement_matrix_kernel -- src/refinement_matrix.jl:329-363: dense (m, n) column-major, zero-filled by the callee. */
int sg_refmat_collect_f32(float *out, const int32_t *row_pointer, const int32_t *column_start, const float *nzval, int64_t m, int64_t n,
                          int64_t nnz, void *stream);
int sg_refmat_collect_f64(double *out, const int32_t *row_pointer, const int32_t *column_start, const double *nzval, int64_t m, int64_t n,
                          int64_t nnz, void *stream);
/* K18 refinement_values_new_kernel -- src/control_points.jl:427-456: rows < n_old copy the old values, the others read the
 * refined control points at their (1-based) indices. */
int sg_refinement_values_new_f32(float *values_new, const float *values_old, int64_t n_old, const float *control_points_refined, int nin,
                                 const int64_t *n_cp, int nout, const int32_t *refinement_indices_new, int64_t n_new, void *stream);
int sg_refinement_values_new_f64(double *values_new, const double *values_old, int64_t n_old, const double *control_points_refined, int nin,
                                 const int64_t *n_cp, int nout, const int32_t *refinement_indices_new, int64_t n_new, void *stream);
/* Flag (Bool) variants for deactivate_overwritten_control_points! -- src/control_points.jl:584-680, src/utils.jl:237-247:
 * K7 on a Flag array (set the active entries to `value`), K6 (src/adjoint.jl:117-121: B .= false, B[J] = true for every J in
 * the structural window of a true Y[I]), K8 (read the flag at the active entries).  Flags are uint8 0/1. */
int sg_scatter_active_flag(uint8_t *cp_flags, int nin, const int64_t *n_cp, const int32_t *refinement_indices, int64_t n_active,
                           int value, void *stream);
int sg_refmat_mul_adjoint_flag(uint8_t *B, const uint8_t *Y, int ndims, const int64_t *sizeY, const int64_t *sizeB, int n_ref,
                               const int *dims, const int32_t *const *row_ptr, const int32_t *const *col_start, const int64_t *nnz,
                               void *stream);
int sg_gather_active_flag(uint8_t *values, const uint8_t *cp_flags, int nin, const int64_t *n_cp, const int32_t *refinement_indices,
                          int64_t n_active, void *stream);
/* findall(flags) (invert = 0) / findall(.!flags) (invert = 1): ascending 0-based positions; count on the host (synchronises).
 * scratch: n + 1 Int32. */
int sg_compact_flags(int32_t *out_idx, int64_t *count_host, const uint8_t *flags, int64_t n, int invert, int32_t *scratch, void *stream);
/* out[r, c] = in[row_idx[r], c] for column-major (n_in, ncols) matrices (row selection `A[where, :]`, :669-670). */
int sg_gather_rows_f32(float *out, const float *in, const int32_t *row_idx, int64_t n_in, int64_t n_out, int ncols, void *stream);
int sg_gather_rows_f64(double *out, const double *in, const int32_t *row_idx, int64_t n_in, int64_t n_out, int ncols, void *stream);
int sg_gather_rows_i32(int32_t *out, const int32_t *in, const int32_t *row_idx, int64_t n_in, int64_t n_out, int ncols, void *stream);
/* error_informed_local_refinement! -- src/control_points.jl:556-566: grid_err = sum over outputs, threshold =
 * threshold_factor * mean(grid_err) (fixed summation order), flags = grid_err > threshold.  scratch: ceil(cp_total / 256) T. */
int sg_error_flags_f32(uint8_t *flags, float *grid_err, float *threshold_out_or_null, const float *cp_err, int64_t cp_total, int nout,
                       float threshold_factor, float *scratch_block_sums, void *stream);
int sg_error_flags_f64(uint8_t *flags, double *grid_err, double *threshold_out_or_null, const double *cp_err, int64_t cp_total, int nout,
                       double threshold_factor, double *scratch_block_sums, void *stream);
/* (n, nin) 1-based Int32 index matrix of 0-based linear positions in a column-major (n_cp...) grid (findall + collect_indices). */
int sg_indices_from_linear_i32(int32_t *indices, const int32_t *linear0, int64_t n, int nin, const int64_t *n_cp, void *stream);
/* unique(vcat(old, new); dims = 1) keeping first occurrences (src/control_points.jl:482-494, on the CPU in the reference):
 * keep_new[i] = 1 iff row i of new_indices is the first row naming its control point.  first_row_scratch: prod(n_cp) Int32. */
int sg_unique_new_rows_i32(uint8_t *keep_new, const int32_t *old_indices, int64_t n_old, const int32_t *new_indices, int64_t n_new,
                           int nin, const int64_t *n_cp, int32_t *first_row_scratch, void *stream);

/* ---- multi-GPU: gradient all-reduce (new; the reference is single-device) -------------------
 * One communicator per process/GPU.  `unique_id` is a 128-byte ncclUniqueId produced by
 * sg_comm_unique_id on rank 0 and distributed by the host (MPI, torch.distributed store, files).
 * NCCL is resolved at run time (dlopen of libnccl.so.2); SG_ERR_NCCL if it is not available. */
typedef struct sg_comm sg_comm;
int sg_comm_unique_id(void *unique_id_128_bytes);
int sg_comm_create(sg_comm **comm, int world_size, int rank, const void *unique_id_128_bytes);
int sg_comm_destroy(sg_comm *comm);
int sg_allreduce_sum_f32(float *buf, int64_t count, sg_comm *comm, void *stream);
int sg_allreduce_sum_f64(double *buf, int64_t count, sg_comm *comm, void *stream);

/* ---- multi-GPU: gradient exchange over NVLink peer memory (new; replaces the all-reduce when the ranks'
 * staging buffers are peer-mapped, e.g. CUDA IPC / symmetric memory) ------------------------------------
 * A slab's partial gradient is non-zero only on control planes [k0, k0+np) of the slowest control axis.
 * sg_exchange_push:   copy those planes of `grad` (plane_elems, c_last, nout) into slot `my_rank` of EVERY
 *                     rank's staging buffer with peer-to-peer stores; peer_stage is a HOST array of `world`
 *                     device pointers (this rank's own buffer included).  Staging layout per rank:
 *                     [world][nout][max_planes][plane_elems].
 * (all ranks then synchronise: sg_exchange_signal + sg_exchange_wait_reduce below, or a host-framework barrier)
 * sg_exchange_reduce: grad[:, k, o] = sum over ranks r whose support covers plane k of stage[r][o][k-k0_r][:],
 *                     summed in rank order (deterministic); k0s / nps are HOST arrays of length world. */
/* sg_evaluate_adjoint_push: sg_evaluate_adjoint followed by sg_exchange_push of the result, as ONE call.  When the
 * 3-D double-march pipeline runs, its last kernel stores every finished control plane of the support straight into the
 * peers' staging slots (peer-to-peer stores overlapping the computation; no separate push kernel); other pipelines
 * run the push kernel afterwards.  Either way the staging buffers hold what sg_exchange_push would have written.
 * keep_local = 1: `control_points` holds the local partial gradient on return (as after sg_evaluate_adjoint);
 * keep_local = 0: its contents are unspecified (the caller is about to overwrite it with the reduce; the fused pipeline
 * then skips its own 16.8 MB of zero / result writes on C3). */
int sg_evaluate_adjoint_push_f32(float *control_points, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                 const float *const *tables, const int32_t *const *sample_indices, const int *degree,
                                 const int *max_derivative_order, const int *derivative_order, const float *eval,
                                 const float *weights, void *workspace, size_t workspace_bytes, void *const *peer_stage,
                                 int world, int my_rank, int64_t k0, int64_t np, int64_t max_planes, int keep_local, void *stream);
int sg_evaluate_adjoint_push_f64(double *control_points, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                 const double *const *tables, const int32_t *const *sample_indices, const int *degree,
                                 const int *max_derivative_order, const int *derivative_order, const double *eval,
                                 const double *weights, void *workspace, size_t workspace_bytes, void *const *peer_stage,
                                 int world, int my_rank, int64_t k0, int64_t np, int64_t max_planes, int keep_local, void *stream);
int sg_exchange_push_f32(const float *grad, void *const *peer_stage, int world, int my_rank, int64_t plane_elems,
                         int64_t c_last, int nout, int64_t k0, int64_t np, int64_t max_planes, void *stream);
int sg_exchange_push_f64(const double *grad, void *const *peer_stage, int world, int my_rank, int64_t plane_elems,
                         int64_t c_last, int nout, int64_t k0, int64_t np, int64_t max_planes, void *stream);
int sg_exchange_reduce_f32(float *grad, const float *stage, int world, const int64_t *k0s, const int64_t *nps,
                           int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream);
int sg_exchange_reduce_f64(double *grad, const double *stage, int world, const int64_t *k0s, const int64_t *nps,
                           int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream);


/* ---- flag-synchronised exchange: the barrier between push and reduce inside the C ABI (no host framework) ----------
 * flags      peer-mapped array of `world` uint64 per rank, zero-initialised once: flags[r] on rank q counts the exchanges
 *            whose push from rank r has completely landed in q's staging buffer (monotone, never reset).
 * local_sync SG_EXCHANGE_SYNC_BYTES of LOCAL device memory, zero-initialised once: the number of exchanges this rank has
 *            completed lives there, on the device, so the calls can be captured in a CUDA graph and replayed.
 * sg_exchange_signal       stream-ordered after this rank's push: release-store of (epoch + 1) into flags[my_rank] of every
 *                          rank (peer_flags: HOST array of `world` device pointers to the ranks' flag arrays).
 * sg_exchange_wait_reduce  waits ON THE DEVICE until all `world` flags of this rank reached (epoch + 1), then reduces like
 *                          sg_exchange_reduce and publishes the new epoch.  peer_flags_or_null != NULL: the signal is
 *                          issued by the same kernel first (one launch for signal + barrier + reduce).
 * Consecutive exchanges must alternate between TWO staging buffers (they may share flags and local_sync).  All ranks must
 * have their kernels in flight concurrently (one process or thread per GPU); a peer that never signals makes the wait give
 * up after 20 s and raises the `timed_out` flag that sg_exchange_status reports (blocking: it synchronises `stream`). */
#define SG_EXCHANGE_SYNC_BYTES 64
int sg_exchange_signal(void *const *peer_flags, int world, int my_rank, const void *local_sync, void *stream);
int sg_exchange_status(const void *local_sync, unsigned long long *epoch, int *timed_out, void *stream);
int sg_exchange_wait_reduce_f32(float *grad, const float *stage, const void *my_flags, void *local_sync,
                                void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s, const int64_t *nps,
                                int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream);
int sg_exchange_wait_reduce_f64(double *grad, const double *stage, const void *my_flags, void *local_sync,
                                void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s, const int64_t *nps,
                                int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream);

/* ---- support-plane exchange: the halo variant of the two calls above (new; the reference is single-device) -----------
 * evaluate! on a slab reads only the control planes of the slab's support [k0s[r], k0s[r] + nps[r]) (0-based planes of the
 * slowest control axis), so a fitting loop evaluate! -> evaluate_adjoint! -> update needs the SUMMED gradient on those planes
 * only.  sg_evaluate_adjoint_planned_support pushes every finished plane to the ranks whose support contains it (the p halo
 * planes a slab shares with each neighbour) instead of to all ranks; sg_exchange_wait_reduce_support signals and waits for
 * those neighbours only -- no global barrier -- and writes grad[:, k, :] for the planes k of its own support (summed in rank
 * order: bit-identical on every rank that holds the plane and to sg_exchange_wait_reduce); all other planes of `grad` are left
 * untouched.  Per-rank NVLink ingress drops from (world - 1)/world of the gradient to the halo planes.  Same staging layout,
 * flags, local_sync and two-buffer rule as above; every rank must use the same variant in a given exchange.  k0s / nps: HOST
 * arrays of `world` entries.  Pipelines without the fused push fall back to a push of all planes to all ranks (correct). */
/* host-only helper (no device needed): which LOCAL planes [dst_lo[r], dst_hi[r]) of rank my_rank's support go to rank r */
int sg_exchange_support_ranges(int world, int my_rank, const int64_t *k0s, const int64_t *nps, int64_t max_planes,
                               int *dst_lo, int *dst_hi);
int sg_evaluate_adjoint_planned_support_f32(const sg_adjoint_plan *plan, float *cp, const float *eval, const float *weights_or_null,
                                            void *workspace, size_t workspace_bytes, void *const *peer_stage, int world,
                                            int my_rank, const int64_t *k0s, const int64_t *nps, int64_t max_planes,
                                            int keep_local, void *stream);
int sg_evaluate_adjoint_planned_support_f64(const sg_adjoint_plan *plan, double *cp, const double *eval, const double *weights_or_null,
                                            void *workspace, size_t workspace_bytes, void *const *peer_stage, int world,
                                            int my_rank, const int64_t *k0s, const int64_t *nps, int64_t max_planes,
                                            int keep_local, void *stream);
int sg_exchange_wait_reduce_support_f32(float *grad, const float *stage, const void *my_flags, void *local_sync,
                                        void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s,
                                        const int64_t *nps, int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes,
                                        void *stream);
int sg_exchange_wait_reduce_support_f64(double *grad, const double *stage, const void *my_flags, void *local_sync,
                                        void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s,
                                        const int64_t *nps, int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes,
                                        void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPLINEGRIDS_B200_H */
