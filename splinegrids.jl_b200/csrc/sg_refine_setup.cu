// sg_refine_setup.cu -- construction side of local refinement on the device (row f3 of the scope table):
//   K13 build_refinement_matrix_kernel          src/refinement.jl:3-36
//   K14 validate_refinement_matrix_kernel       src/refinement_matrix.jl:134-181
//   K16 refinement_matrix_mul_nonzeros_kernel   src/refinement_matrix.jl:229-271
//   K15 refinement_matrix_multiplication_kernel src/refinement_matrix.jl:184-227
//   K17 collect_refinement_matrix_kernel        src/refinement_matrix.jl:329-347
//   K18 refinement_values_new_kernel            src/control_points.jl:427-456
//   Flag (Bool) variants of K7 / K6 / K8 used by deactivate_overwritten_control_points!
//                                               src/control_points.jl:584-680, src/utils.jl:237-247, src/adjoint.jl:117-121
//   the reduce + threshold + findall of error_informed_local_refinement!  src/control_points.jl:541-575
//   unique(vcat(old, new); dims = 1) of activate_local_refinement!        src/control_points.jl:482-494 (CPU in the reference)
// plus the exclusive scan / stream compaction they need.  Arrays are small (O(n_knots), O(control points)); every kernel
// is one thread per row / element like the reference's, launch-latency bound.  Index arrays are 1-based Int32.
#include <algorithm>

#include "sg_common.cuh"

// ---- column range of row i (1-based) -- src/refinement_matrix.jl:100-125 ------------------------------------------
__device__ __forceinline__ void sg_col_range(const int32_t *__restrict__ rp, const int32_t *__restrict__ cs, int64_t m, int64_t nnz,
                                             int64_t i, int64_t &c0, int64_t &c1)
{
    const int64_t next = (i == m) ? nnz + 1 : rp[i];   // rp[i] is row i+1 (0-based array)
    c0 = cs[i - 1];
    c1 = c0 + (next - rp[i - 1]) - 1;
}

// ---- K13 ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sg_boehm_kernel(int32_t *__restrict__ rp, int32_t *__restrict__ cs, T *__restrict__ nz, const T *__restrict__ knots_old,
                                int64_t n_rows, int64_t k, T knot_new, int p)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;   // 1-based row
    if (i > n_rows) return;
    if (i <= k - p) {
        rp[i - 1] = (int32_t)i; cs[i - 1] = (int32_t)i;
        nz[i - 1] = T(1);
    } else if (i <= k) {
        const T alpha = sg_div(sg_sub(knot_new, knots_old[i - 1]), sg_sub(knots_old[i + p - 1], knots_old[i - 1]));
        const int64_t r = 2 * i - k + p - 1;
        rp[i - 1] = (int32_t)r; cs[i - 1] = (int32_t)(i - 1);
        nz[r - 1] = sg_sub(T(1), alpha);
        nz[r] = alpha;
    } else {
        const int64_t r = i + p;
        rp[i - 1] = (int32_t)r; cs[i - 1] = (int32_t)(i - 1);
        nz[r - 1] = T(1);
    }
}

// ---- K14 ------------------------------------------------------------------------------------------------------------
__global__ void sg_refmat_validate_kernel(uint8_t *__restrict__ valid, const int32_t *__restrict__ rp, const int32_t *__restrict__ cs, int64_t m,
                                          int64_t nnz, int64_t n_columns)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
    if (i > m) return;
    int64_t c0, c1;
    sg_col_range(rp, cs, m, nnz, i, c0, c1);
    bool ok = c1 >= c0;
    if (ok) ok = c0 >= 1 && c1 <= n_columns;
    const bool first = i == 1;
    if (first) ok = c0 == 1 && rp[0] == 1;
    int64_t p0 = 0, p1 = 0;
    if (ok && !first) {
        sg_col_range(rp, cs, m, nnz, i - 1, p0, p1);
        ok = p0 <= c0 && c0 <= p1 + 1;
    }
    if (ok && !first) ok = c1 >= p1;
    valid[i - 1] = ok ? 1 : 0;
}

// ---- K16 / K15: C = A * B -------------------------------------------------------------------------------------------
__global__ void sg_refmat_mul_nonzeros_kernel(int32_t *__restrict__ nnz_C, int32_t *__restrict__ cs_C, const int32_t *__restrict__ rpA,
                                              const int32_t *__restrict__ rpB, const int32_t *__restrict__ csA, const int32_t *__restrict__ csB,
                                              int64_t mA, int64_t mB, int64_t nnzA, int64_t nnzB, int64_t n_columns_B)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
    if (i > mA) return;
    int64_t a0, a1;
    sg_col_range(rpA, csA, mA, nnzA, i, a0, a1);
    int n_non_zeros = 0, column_start = 0;
    for (int64_t j = n_columns_B; j >= 1; --j) {
        for (int64_t k = a0; k <= a1; ++k) {
            int64_t b0, b1;
            sg_col_range(rpB, csB, mB, nnzB, k, b0, b1);
            if (b0 <= j && j <= b1) {
                ++n_non_zeros;
                column_start = (int)j;
                break;
            }
        }
    }
    nnz_C[i - 1] = n_non_zeros;
    cs_C[i - 1] = column_start;
}

template <typename T>
__global__ void sg_refmat_mul_values_kernel(T *__restrict__ nzC, const int32_t *__restrict__ rpC, const int32_t *__restrict__ csC,
                                            const int32_t *__restrict__ rpA, const int32_t *__restrict__ rpB, const int32_t *__restrict__ csA,
                                            const int32_t *__restrict__ csB, const T *__restrict__ nzA, const T *__restrict__ nzB, int64_t mA,
                                            int64_t mB, int64_t nnzA, int64_t nnzB, int64_t nnzC)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
    if (i > mA) return;
    int64_t a0, a1, c0, c1;
    sg_col_range(rpA, csA, mA, nnzA, i, a0, a1);
    sg_col_range(rpC, csC, mA, nnzC, i, c0, c1);
    int64_t pA = rpA[i - 1];
    for (int64_t k = a0; k <= a1; ++k) {
        int64_t b0, b1;
        sg_col_range(rpB, csB, mB, nnzB, k, b0, b1);
        int64_t pC = rpC[i - 1];
        for (int64_t j = c0; j <= c1; ++j) {
            if (b0 <= j && j <= b1) {
                const int64_t pB = rpB[k - 1] + j - b0;
                nzC[pC - 1] = sg_add(nzC[pC - 1], sg_mul(nzA[pA - 1], nzB[pB - 1]));   // ascending k, like the reference
            }
            ++pC;
        }
        ++pA;
    }
}

// ---- K17 ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sg_refmat_collect_kernel(T *__restrict__ out, const int32_t *__restrict__ rp, const int32_t *__restrict__ cs, const T *__restrict__ nz,
                                         int64_t m, int64_t nnz)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;
    if (i > m) return;
    const int64_t d0 = rp[i - 1], d1 = (i == m) ? nnz : rp[i] - 1;
    int64_t col = cs[i - 1];
    for (int64_t d = d0; d <= d1; ++d, ++col) out[(i - 1) + m * (col - 1)] = nz[d - 1];   // out (m, n) column-major
}

// ---- K18 ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sg_refinement_values_new_kernel(T *__restrict__ values_new, const T *__restrict__ values_old, int64_t n_old,
                                                const T *__restrict__ cp_refined, const int32_t *__restrict__ idx_new, int64_t n_new, int nin,
                                                int nout, SgGridArgs<T> shape /* n_cp only */)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_new) return;
    if (i < n_old) {
        for (int o = 0; o < nout; ++o) values_new[i + n_new * o] = values_old[i + n_old * o];
        return;
    }
    int64_t lin = idx_new[i] - 1, sp = 1;
    for (int d = 1; d < nin; ++d) {
        sp *= shape.n_cp[d - 1];
        lin += (int64_t)(idx_new[i + n_new * d] - 1) * sp;
    }
    for (int o = 0; o < nout; ++o) values_new[i + n_new * o] = cp_refined[lin + shape.cp_total * o];
}

// ---- Flag variants ----------------------------------------------------------------------------------------------------
// K7 with Flag values: cp_flags[idx[i, :]] = value (src/control_points.jl:296-311 on a Flag array, one output)
__global__ void sg_scatter_flag_kernel(uint8_t *__restrict__ flags, const int32_t *__restrict__ idx, int64_t n_active, int nin,
                                       SgGridArgs<float> shape, uint8_t value)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    int64_t lin = 0, sp = 1;
    for (int d = 0; d < nin; ++d) { lin += (int64_t)(idx[i + n_active * d] - 1) * sp; sp *= shape.n_cp[d]; }
    flags[lin] = value;
}
// K8 on a Flag array: values[i] = flags[idx[i, :]]  (the reference also zeroes the entry; the array is discarded)
__global__ void sg_gather_flag_kernel(uint8_t *__restrict__ values, const uint8_t *__restrict__ flags, const int32_t *__restrict__ idx,
                                      int64_t n_active, int nin, SgGridArgs<float> shape)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    int64_t lin = 0, sp = 1;
    for (int d = 0; d < nin; ++d) { lin += (int64_t)(idx[i + n_active * d] - 1) * sp; sp *= shape.n_cp[d]; }
    values[i] = flags[lin];
}
// K6 on Flag arrays (src/adjoint.jl:93-124 with `contrib isa Flag`): B .= false, then B[J] = true for every J in the
// structural window of every Y[I] that is true (Flag * number = Flag).  All writers store the same value: no atomics.
struct SgFlagMulArgs {
    int ndims;
    int64_t sizeY[SG_MAX_DIMS], sizeB[SG_MAX_DIMS];
    int ref_of_dim[SG_MAX_DIMS];            // index into rp/cs (or -1)
    const int32_t *rp[SG_MAX_DIMS], *cs[SG_MAX_DIMS];
    int64_t nnz[SG_MAX_DIMS];
    int64_t totalY;
};
__global__ void sg_refmat_mul_adjoint_flag_kernel(uint8_t *__restrict__ B, const uint8_t *__restrict__ Y, const __grid_constant__ SgFlagMulArgs a)
{
    const int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (lin >= a.totalY || !Y[lin]) return;
    int64_t c0[SG_MAX_DIMS], nc[SG_MAX_DIMS], r = lin, total = 1;
    for (int d = 0; d < a.ndims; ++d) {
        const int64_t I = r % a.sizeY[d] + 1;
        r /= a.sizeY[d];
        const int q = a.ref_of_dim[d];
        if (q >= 0) {
            int64_t lo, hi;
            sg_col_range(a.rp[q], a.cs[q], a.sizeY[d], a.nnz[q], I, lo, hi);
            c0[d] = lo; nc[d] = hi - lo + 1;
        } else {
            c0[d] = I; nc[d] = 1;
        }
        total *= nc[d] > 0 ? nc[d] : 0;
    }
    for (int64_t t = 0; t < total; ++t) {
        int64_t rr = t, off = 0, st = 1;
        for (int d = 0; d < a.ndims; ++d) {
            const int64_t j = c0[d] + rr % nc[d];
            rr /= nc[d];
            off += (j - 1) * st;
            st *= a.sizeB[d];
        }
        B[off] = 1;
    }
}

// ---- exclusive scan of small/medium Int32 arrays + stream compaction ------------------------------------------------------
#define SG_SCAN_BLOCK 1024
__global__ void sg_scan_block_kernel(int32_t *__restrict__ out, const uint8_t *__restrict__ flags, const int32_t *__restrict__ in, int64_t n,
                                     int32_t *__restrict__ block_sums, int invert)
{
    __shared__ int32_t s[SG_SCAN_BLOCK];
    const int64_t i = blockIdx.x * (int64_t)SG_SCAN_BLOCK + threadIdx.x;
    const int32_t v = i < n ? (flags ? (((flags[i] != 0) != (invert != 0)) ? 1 : 0) : in[i]) : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < SG_SCAN_BLOCK; off <<= 1) {   // Hillis-Steele inclusive scan
        const int32_t t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = s[threadIdx.x] - v;               // exclusive
    if (threadIdx.x == SG_SCAN_BLOCK - 1 && block_sums) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void sg_scan_add_kernel(int32_t *__restrict__ out, const int32_t *__restrict__ block_offsets, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)SG_SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_offsets[blockIdx.x];
}
// out[n] (exclusive prefix sums) and total in out[n] (the array has n + 1 entries)
static int sg_scan_exclusive(int32_t *out, const uint8_t *flags, const int32_t *in, int64_t n, cudaStream_t st, int invert = 0)
{
    if (n <= 0) return cudaMemsetAsync(out, 0, sizeof(int32_t), st) == cudaSuccess ? SG_OK : SG_ERR_INVALID_ARGUMENT;
    const int64_t nb = (n + SG_SCAN_BLOCK - 1) / SG_SCAN_BLOCK;
    int32_t *sums = nullptr;
    SG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&sums), (size_t)(2 * nb + 2) * sizeof(int32_t), st));
    sg_scan_block_kernel<<<(unsigned)nb, SG_SCAN_BLOCK, 0, st>>>(out, flags, in, n, sums, invert);
    g_sg_launches.fetch_add(1);
    int rc = SG_OK;
    if (nb > 1) {
        rc = sg_scan_exclusive(sums + nb, nullptr, sums, nb, st);   // offsets of the blocks (+ total at [nb + nb])
        if (rc == SG_OK) {
            sg_scan_add_kernel<<<(unsigned)nb, SG_SCAN_BLOCK, 0, st>>>(out, sums + nb, n);
            g_sg_launches.fetch_add(1);
            SG_CUDA(cudaMemcpyAsync(out + n, sums + nb + nb, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        }
    } else {
        SG_CUDA(cudaMemcpyAsync(out + n, sums, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    cudaFreeAsync(sums, st);
    return rc;
}
__global__ void sg_compact_kernel(int32_t *__restrict__ out_idx, const uint8_t *__restrict__ flags, const int32_t *__restrict__ pos, int64_t n,
                                  int invert)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && ((flags[i] != 0) != (invert != 0))) out_idx[pos[i]] = (int32_t)i;   // ascending: findall order
}

// findall(flags) (invert = 0) or findall(.!flags) (invert = 1): out_idx[0 .. count) = 0-based positions, ascending;
// *count_host after a stream sync.  scratch: (n + 1) Int32 (device).
extern "C" int sg_compact_flags(int32_t *out_idx, int64_t *count_host, const uint8_t *flags, int64_t n, int invert, int32_t *scratch,
                                void *stream)
{
    SG_CHECK_ARG(count_host && (n == 0 || (out_idx && flags && scratch)) && n >= 0);
    *count_host = 0;
    if (n == 0) return SG_OK;
    cudaStream_t st = sg_stream(stream);
    int rc = sg_scan_exclusive(scratch, flags, nullptr, n, st, invert);
    if (rc != SG_OK) return rc;
    sg_compact_kernel<<<sg_blocks(n, 256), 256, 0, st>>>(out_idx, flags, scratch, n, invert);
    g_sg_launches.fetch_add(1);
    int32_t total = 0;
    SG_CUDA(cudaMemcpyAsync(&total, scratch + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    *count_host = total;
    return SG_OK;
}

// row_pointer = 1 + exclusive scan of the per-row counts; *total_host = sum (stream sync) -- src/refinement_matrix.jl:300-303
__global__ void sg_add_one_kernel(int32_t *__restrict__ v, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] += 1;
}
extern "C" int sg_row_pointer_from_counts(int32_t *row_pointer /* m + 1 entries */, int64_t *total_host, const int32_t *counts, int64_t m,
                                          void *stream)
{
    SG_CHECK_ARG(row_pointer && total_host && counts && m >= 1);
    cudaStream_t st = sg_stream(stream);
    int rc = sg_scan_exclusive(row_pointer, nullptr, counts, m, st);
    if (rc != SG_OK) return rc;
    int32_t total = 0;
    SG_CUDA(cudaMemcpyAsync(&total, row_pointer + m, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    sg_add_one_kernel<<<sg_blocks(m, 256), 256, 0, st>>>(row_pointer, m);
    g_sg_launches.fetch_add(1);
    SG_CUDA(cudaStreamSynchronize(st));
    *total_host = total;
    return SG_OK;
}

// gather rows of a column-major (n_in, ncols) matrix: out[r, c] = in[row_idx[r], c]
template <typename T>
__global__ void sg_gather_rows_kernel(T *__restrict__ out, const T *__restrict__ in, const int32_t *__restrict__ row_idx, int64_t n_in,
                                      int64_t n_out, int ncols)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_out) return;
    const int64_t src = row_idx[r];
    for (int c = 0; c < ncols; ++c) out[r + n_out * c] = in[src + n_in * c];
}

// ---- error_informed_local_refinement!: sum over outputs, mean, threshold, flags --------------------------------------------
template <typename T>
__global__ void sg_output_sum_kernel(T *__restrict__ grid_err, T *__restrict__ block_sums, const T *__restrict__ cp_err, int64_t cp_total, int nout)
{
    __shared__ T s[256];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    T v = T(0);
    if (i < cp_total) {
        for (int o = 0; o < nout; ++o) v += cp_err[i + cp_total * o];   // sum(...; dims = Nin + 1)
        grid_err[i] = v;
    }
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {   // fixed tree: deterministic
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s[0];
}
template <typename T>
__global__ void sg_threshold_flags_kernel(uint8_t *__restrict__ flags, T *__restrict__ threshold_out, const T *__restrict__ grid_err,
                                          const T *__restrict__ block_sums, int64_t n_blocks, int64_t cp_total, T threshold_factor)
{
    __shared__ T thr;
    if (threadIdx.x == 0) {   // every block re-derives the same threshold in the same order (n_blocks is small)
        T tot = T(0);
        for (int64_t b = 0; b < n_blocks; ++b) tot += block_sums[b];
        thr = threshold_factor * tot / (T)cp_total;
        if (blockIdx.x == 0 && threshold_out) *threshold_out = thr;
    }
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < cp_total) flags[i] = grid_err[i] > thr ? 1 : 0;
}

// linear 0-based positions -> (n, nin) 1-based Int32 index matrix, column-major (collect_indices of CartesianIndices)
__global__ void sg_indices_from_linear_kernel(int32_t *__restrict__ idx, const int32_t *__restrict__ lin, int64_t n, int nin,
                                              SgGridArgs<float> shape)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t r = lin[i];
    for (int d = 0; d < nin; ++d) {
        idx[i + n * d] = (int32_t)(r % shape.n_cp[d]) + 1;
        r /= shape.n_cp[d];
    }
}

// ---- unique(vcat(old, new); dims = 1), keeping first occurrences: which NEW rows survive ------------------------------------
// first_row[cell] = smallest row number (old rows first) that names the cell; a new row is kept iff it is that row.
__global__ void sg_first_row_kernel(int32_t *__restrict__ first_row, const int32_t *__restrict__ idx, int64_t n, int64_t row_offset, int nin,
                                    SgGridArgs<float> shape)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lin = 0, sp = 1;
    for (int d = 0; d < nin; ++d) { lin += (int64_t)(idx[i + n * d] - 1) * sp; sp *= shape.n_cp[d]; }
    atomicMin(first_row + lin, (int32_t)(row_offset + i));
}
__global__ void sg_keep_new_kernel(uint8_t *__restrict__ keep, const int32_t *__restrict__ first_row, const int32_t *__restrict__ idx, int64_t n,
                                   int64_t row_offset, int nin, SgGridArgs<float> shape)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lin = 0, sp = 1;
    for (int d = 0; d < nin; ++d) { lin += (int64_t)(idx[i + n * d] - 1) * sp; sp *= shape.n_cp[d]; }
    keep[i] = first_row[lin] == (int32_t)(row_offset + i) ? 1 : 0;
}

// =====================================================================================================================
// C ABI
// =====================================================================================================================
static int sg_shape_args(SgGridArgs<float> &s, int nin, const int64_t *n_cp)
{
    if (nin < 1 || nin > SG_MAX_DIMS || !n_cp) return SG_ERR_INVALID_ARGUMENT;
    s = SgGridArgs<float>{};
    s.nin = nin;
    s.cp_total = 1;
    for (int d = 0; d < nin; ++d) {
        if (n_cp[d] < 1) return SG_ERR_INVALID_ARGUMENT;
        s.n_cp[d] = n_cp[d];
        s.cp_total *= n_cp[d];
    }
    return SG_OK;
}

#define SG_DEFINE_SETUP_API(T, SUF)                                                                                              \
    extern "C" int sg_boehm_matrix_##SUF(int32_t *row_pointer, int32_t *column_start, T *nzval, const T *knots_all_old,          \
                                         int64_t n_basis_old, int64_t knot_span_index, T knot_new, int degree, void *stream)     \
    {                                                                                                                            \
        SG_CHECK_ARG(row_pointer && column_start && nzval && knots_all_old && n_basis_old >= 1 && degree >= 0);                  \
        SG_CHECK_ARG(knot_span_index >= degree + 1 && knot_span_index <= n_basis_old);                                           \
        const int64_t rows = n_basis_old + 1;                                                                                    \
        sg_boehm_kernel<T><<<sg_blocks(rows, 128), 128, 0, sg_stream(stream)>>>(row_pointer, column_start, nzval, knots_all_old, \
                                                                                rows, knot_span_index, knot_new, degree);       \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }                                                                                                                            \
    extern "C" int sg_refmat_mul_values_##SUF(T *nzval_C, int64_t nnz_C, const int32_t *rpC, const int32_t *csC,                 \
                                              const int32_t *rpA, const int32_t *csA, const T *nzA, int64_t mA, int64_t nnzA,    \
                                              const int32_t *rpB, const int32_t *csB, const T *nzB, int64_t mB, int64_t nnzB,    \
                                              void *stream)                                                                      \
    {                                                                                                                            \
        SG_CHECK_ARG(nzval_C && rpC && csC && rpA && csA && nzA && rpB && csB && nzB && mA >= 1 && mB >= 1 && nnz_C >= 0);       \
        cudaStream_t st = sg_stream(stream);                                                                                     \
        SG_CUDA(cudaMemsetAsync(nzval_C, 0, (size_t)nnz_C * sizeof(T), st));                                                     \
        sg_refmat_mul_values_kernel<T><<<sg_blocks(mA, 128), 128, 0, st>>>(nzval_C, rpC, csC, rpA, rpB, csA, csB, nzA, nzB, mA,  \
                                                                           mB, nnzA, nnzB, nnz_C);                               \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }                                                                                                                            \
    extern "C" int sg_refmat_collect_##SUF(T *out, const int32_t *rp, const int32_t *cs, const T *nz, int64_t m, int64_t n,      \
                                           int64_t nnz, void *stream)                                                            \
    {                                                                                                                            \
        SG_CHECK_ARG(out && rp && cs && nz && m >= 1 && n >= 1);                                                                 \
        cudaStream_t st = sg_stream(stream);                                                                                     \
        SG_CUDA(cudaMemsetAsync(out, 0, (size_t)m * n * sizeof(T), st));                                                         \
        sg_refmat_collect_kernel<T><<<sg_blocks(m, 128), 128, 0, st>>>(out, rp, cs, nz, m, nnz);                                 \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }                                                                                                                            \
    extern "C" int sg_refinement_values_new_##SUF(T *values_new, const T *values_old, int64_t n_old, const T *cp_refined,        \
                                                  int nin, const int64_t *n_cp, int nout, const int32_t *indices_new,            \
                                                  int64_t n_new, void *stream)                                                   \
    {                                                                                                                            \
        SG_CHECK_ARG(n_new >= 0 && n_old >= 0 && n_old <= n_new && nout >= 1);                                                   \
        if (n_new == 0) return SG_OK;                                                                                            \
        SG_CHECK_ARG(values_new && cp_refined && indices_new && (n_old == 0 || values_old));                                     \
        SgGridArgs<float> sh;                                                                                                    \
        int rc = sg_shape_args(sh, nin, n_cp);                                                                                   \
        if (rc != SG_OK) return rc;                                                                                              \
        SgGridArgs<T> sht{};                                                                                                     \
        sht.nin = nin; sht.cp_total = sh.cp_total;                                                                               \
        for (int d = 0; d < nin; ++d) sht.n_cp[d] = n_cp[d];                                                                     \
        sg_refinement_values_new_kernel<T><<<sg_blocks(n_new, 128), 128, 0, sg_stream(stream)>>>(                                \
            values_new, values_old, n_old, cp_refined, indices_new, n_new, nin, nout, sht);                                      \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }                                                                                                                            \
    extern "C" int sg_gather_rows_##SUF(T *out, const T *in, const int32_t *row_idx, int64_t n_in, int64_t n_out, int ncols,     \
                                        void *stream)                                                                            \
    {                                                                                                                            \
        SG_CHECK_ARG(n_out >= 0 && n_in >= 0 && ncols >= 0);                                                                     \
        if (n_out == 0 || ncols == 0) return SG_OK;                                                                              \
        SG_CHECK_ARG(out && in && row_idx);                                                                                      \
        sg_gather_rows_kernel<T><<<sg_blocks(n_out, 128), 128, 0, sg_stream(stream)>>>(out, in, row_idx, n_in, n_out, ncols);    \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }                                                                                                                            \
    extern "C" int sg_error_flags_##SUF(uint8_t *flags, T *grid_err, T *threshold_out, const T *cp_err, int64_t cp_total,        \
                                        int nout, T threshold_factor, T *scratch_block_sums, void *stream)                       \
    {                                                                                                                            \
        SG_CHECK_ARG(flags && grid_err && cp_err && scratch_block_sums && cp_total >= 1 && nout >= 1);                           \
        cudaStream_t st = sg_stream(stream);                                                                                     \
        const unsigned nb = sg_blocks(cp_total, 256);                                                                            \
        sg_output_sum_kernel<T><<<nb, 256, 0, st>>>(grid_err, scratch_block_sums, cp_err, cp_total, nout);                       \
        sg_threshold_flags_kernel<T><<<nb, 256, 0, st>>>(flags, threshold_out, grid_err, scratch_block_sums, nb, cp_total,       \
                                                         threshold_factor);                                                     \
        g_sg_launches.fetch_add(1);                                                                                              \
        SG_AFTER_LAUNCH();                                                                                                       \
        return SG_OK;                                                                                                            \
    }

SG_DEFINE_SETUP_API(float, f32)
SG_DEFINE_SETUP_API(double, f64)

extern "C" int sg_gather_rows_i32(int32_t *out, const int32_t *in, const int32_t *row_idx, int64_t n_in, int64_t n_out, int ncols, void *stream)
{
    SG_CHECK_ARG(n_out >= 0 && n_in >= 0 && ncols >= 0);
    if (n_out == 0 || ncols == 0) return SG_OK;
    SG_CHECK_ARG(out && in && row_idx);
    sg_gather_rows_kernel<int32_t><<<sg_blocks(n_out, 128), 128, 0, sg_stream(stream)>>>(out, in, row_idx, n_in, n_out, ncols);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_refmat_validate_i32(uint8_t *valid_row, const int32_t *row_pointer, const int32_t *column_start, int64_t m, int64_t nnz,
                                      int64_t n_columns, void *stream)
{
    SG_CHECK_ARG(valid_row && row_pointer && column_start && m >= 1);
    sg_refmat_validate_kernel<<<sg_blocks(m, 128), 128, 0, sg_stream(stream)>>>(valid_row, row_pointer, column_start, m, nnz, n_columns);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_refmat_mul_nonzeros_i32(int32_t *n_nonzero_C, int32_t *column_start_C, const int32_t *rpA, const int32_t *csA, int64_t mA,
                                          int64_t nnzA, const int32_t *rpB, const int32_t *csB, int64_t mB, int64_t nnzB, int64_t n_columns_B,
                                          void *stream)
{
    SG_CHECK_ARG(n_nonzero_C && column_start_C && rpA && csA && rpB && csB && mA >= 1 && mB >= 1);
    sg_refmat_mul_nonzeros_kernel<<<sg_blocks(mA, 128), 128, 0, sg_stream(stream)>>>(n_nonzero_C, column_start_C, rpA, rpB, csA, csB, mA, mB,
                                                                                    nnzA, nnzB, n_columns_B);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_scatter_active_flag(uint8_t *cp_flags, int nin, const int64_t *n_cp, const int32_t *refinement_indices, int64_t n_active,
                                      int value, void *stream)
{
    if (n_active == 0) return SG_OK;
    SG_CHECK_ARG(cp_flags && refinement_indices && n_active > 0);
    SgGridArgs<float> sh;
    int rc = sg_shape_args(sh, nin, n_cp);
    if (rc != SG_OK) return rc;
    sg_scatter_flag_kernel<<<sg_blocks(n_active, 128), 128, 0, sg_stream(stream)>>>(cp_flags, refinement_indices, n_active, nin, sh,
                                                                                   (uint8_t)(value != 0));
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_gather_active_flag(uint8_t *values, const uint8_t *cp_flags, int nin, const int64_t *n_cp, const int32_t *refinement_indices,
                                     int64_t n_active, void *stream)
{
    if (n_active == 0) return SG_OK;
    SG_CHECK_ARG(values && cp_flags && refinement_indices && n_active > 0);
    SgGridArgs<float> sh;
    int rc = sg_shape_args(sh, nin, n_cp);
    if (rc != SG_OK) return rc;
    sg_gather_flag_kernel<<<sg_blocks(n_active, 128), 128, 0, sg_stream(stream)>>>(values, cp_flags, refinement_indices, n_active, nin, sh);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_refmat_mul_adjoint_flag(uint8_t *B, const uint8_t *Y, int ndims, const int64_t *sizeY, const int64_t *sizeB, int n_ref,
                                          const int *dims, const int32_t *const *row_ptr, const int32_t *const *col_start, const int64_t *nnz,
                                          void *stream)
{
    SG_CHECK_ARG(B && Y && sizeY && sizeB && ndims >= 1 && ndims <= SG_MAX_DIMS && n_ref >= 0 && n_ref <= ndims);
    SgFlagMulArgs a{};
    a.ndims = ndims;
    a.totalY = 1;
    int64_t totalB = 1;
    for (int d = 0; d < ndims; ++d) { a.sizeY[d] = sizeY[d]; a.sizeB[d] = sizeB[d]; a.ref_of_dim[d] = -1; a.totalY *= sizeY[d]; totalB *= sizeB[d]; }
    for (int q = 0; q < n_ref; ++q) {
        SG_CHECK_ARG(dims && row_ptr && col_start && nnz && dims[q] >= 1 && dims[q] <= ndims && row_ptr[q] && col_start[q]);
        a.ref_of_dim[dims[q] - 1] = q;
        a.rp[q] = row_ptr[q]; a.cs[q] = col_start[q]; a.nnz[q] = nnz[q];
    }
    cudaStream_t st = sg_stream(stream);
    SG_CUDA(cudaMemsetAsync(B, 0, (size_t)totalB, st));              /* B .= Flag(false), src/adjoint.jl:135 */
    sg_refmat_mul_adjoint_flag_kernel<<<sg_blocks(a.totalY, 128), 128, 0, st>>>(B, Y, a);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_indices_from_linear_i32(int32_t *indices, const int32_t *linear0, int64_t n, int nin, const int64_t *n_cp, void *stream)
{
    if (n == 0) return SG_OK;
    SG_CHECK_ARG(indices && linear0 && n > 0);
    SgGridArgs<float> sh;
    int rc = sg_shape_args(sh, nin, n_cp);
    if (rc != SG_OK) return rc;
    sg_indices_from_linear_kernel<<<sg_blocks(n, 128), 128, 0, sg_stream(stream)>>>(indices, linear0, n, nin, sh);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

// keep_new[i] = 1 iff row i of `new_indices` names a control point that no row of `old_indices` and no EARLIER row of
// `new_indices` names.  first_row_scratch: prod(n_cp) Int32.
extern "C" int sg_unique_new_rows_i32(uint8_t *keep_new, const int32_t *old_indices, int64_t n_old, const int32_t *new_indices, int64_t n_new,
                                      int nin, const int64_t *n_cp, int32_t *first_row_scratch, void *stream)
{
    if (n_new == 0) return SG_OK;
    SG_CHECK_ARG(keep_new && new_indices && first_row_scratch && n_new > 0 && n_old >= 0 && (n_old == 0 || old_indices));
    SgGridArgs<float> sh;
    int rc = sg_shape_args(sh, nin, n_cp);
    if (rc != SG_OK) return rc;
    cudaStream_t st = sg_stream(stream);
    SG_CUDA(cudaMemsetAsync(first_row_scratch, 0x7f, (size_t)sh.cp_total * sizeof(int32_t), st));   // 0x7f7f7f7f: larger than any row
    if (n_old > 0) {
        sg_first_row_kernel<<<sg_blocks(n_old, 128), 128, 0, st>>>(first_row_scratch, old_indices, n_old, 0, nin, sh);
        g_sg_launches.fetch_add(1);
    }
    sg_first_row_kernel<<<sg_blocks(n_new, 128), 128, 0, st>>>(first_row_scratch, new_indices, n_new, n_old, nin, sh);
    sg_keep_new_kernel<<<sg_blocks(n_new, 128), 128, 0, st>>>(keep_new, first_row_scratch, new_indices, n_new, n_old, nin, sh);
    g_sg_launches.fetch_add(1);
    SG_AFTER_LAUNCH();
    return SG_OK;
}
