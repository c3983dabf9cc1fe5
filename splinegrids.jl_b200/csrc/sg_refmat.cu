// sg_refmat.cu -- application of hierarchical local-refinement matrices as a banded sparse kernel
// (K5 mult!, K6 mult_adjoint!) and the active-control-point scatter/gather (K7, K8).
// Reference: src/refinement_matrix.jl:365-445 (+ row helpers :103-125, src/utils.jl:204-235),
// src/adjoint.jl:85-170, src/control_points.jl:296-311.
#include "sg_common.cuh"

template <typename T>
struct SgRefmatArgs {
    int ndims;
    int64_t sizeY[SG_MAX_DIMS];
    int64_t sizeB[SG_MAX_DIMS];
    int64_t strideB[SG_MAX_DIMS];
    int64_t strideY[SG_MAX_DIMS];
    int64_t totalY, totalB;
    // per array dimension: refinement matrix data or nullptr (unrefined dimension)
    const int32_t *row_ptr[SG_MAX_DIMS];
    const int32_t *col_start[SG_MAX_DIMS];
    const T *nzval[SG_MAX_DIMS];
    int64_t nnz[SG_MAX_DIMS];
};

// Row window of matrix row I (1-based): first column (1-based) and number of stored non-zeros.
// get_column_range / get_column_end, src/refinement_matrix.jl:103-125.
__device__ __forceinline__ void sg_row_window(const int32_t *__restrict__ rp, const int32_t *__restrict__ cs, int64_t m,
                                              int64_t nnz, int64_t I, int64_t &c0, int64_t &nc, int64_t &p0)
{
    p0 = sg_ldg(rp + I - 1);
    const int64_t next = (I == m) ? nnz + 1 : (int64_t)sg_ldg(rp + I);
    c0 = sg_ldg(cs + I - 1);
    nc = next - p0;
}

// K5: one thread per element of Y; product of the per-dimension row windows.
template <typename T>
__global__ void __launch_bounds__(256) sg_refmat_mul_kernel(T *__restrict__ Y, const T *__restrict__ B,
                                                            const __grid_constant__ SgRefmatArgs<T> a)
{
    const int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (lin >= a.totalY) return;
    int64_t c0[SG_MAX_DIMS], nc[SG_MAX_DIMS], p0[SG_MAX_DIMS], Jb[SG_MAX_DIMS];
    int64_t r = lin, nterm = 1, base = 0;
    for (int d = 0; d < a.ndims; ++d) {
        const int64_t I = r % a.sizeY[d] + 1;
        r /= a.sizeY[d];
        if (a.row_ptr[d]) sg_row_window(a.row_ptr[d], a.col_start[d], a.sizeY[d], a.nnz[d], I, c0[d], nc[d], p0[d]);
        else { c0[d] = I; nc[d] = 1; p0[d] = 0; }
        nterm *= nc[d];
        base += (c0[d] - 1) * a.strideB[d];
        Jb[d] = 0;
    }
    T out = T(0);
    for (int64_t q = 0; q < nterm; ++q) {
        int64_t off = base;
        for (int d = 0; d < a.ndims; ++d) off += Jb[d] * a.strideB[d];
        T contrib = sg_ldg(B + off);
        for (int d = 0; d < a.ndims; ++d)   // multiplication order = dimension order, src/refinement_matrix.jl:392-398
            if (a.row_ptr[d]) contrib *= sg_ldg(a.nzval[d] + p0[d] + Jb[d] - 1);
        out += contrib;
        for (int d = 0; d < a.ndims; ++d) { if (++Jb[d] < nc[d]) break; Jb[d] = 0; }
    }
    Y[lin] = out;
}

// K6, atomics-free: one thread per element of B gathers over the rows whose window contains its
// column.  The reference guarantees consecutive non-zeros in every COLUMN as well
// (src/refinement_matrix.jl:4-7 and validation :134-181: row starts and ends are non-decreasing), so the
// rows touching column j form the contiguous range [first row with end >= j, last row with start <= j],
// found by binary search on the monotone row starts / ends.
template <typename T>
__device__ __forceinline__ void sg_column_rows(const int32_t *__restrict__ rp, const int32_t *__restrict__ cs, int64_t m,
                                               int64_t nnz, int64_t j, int64_t &r0, int64_t &r1)
{
    // r1 = last row (1-based) with col_start <= j
    int64_t lo = 0, hi = m;  // first row index (0-based) with cs > j
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sg_ldg(cs + mid) > j) hi = mid; else lo = mid + 1;
    }
    r1 = lo;  // 1-based last row with start <= j (0 = none)
    // r0 = first row with col_end >= j ; col_end(i) = cs[i] + (rp[i+1]-rp[i]) - 1
    lo = 0; hi = m;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        const int64_t next = (mid + 1 == m) ? nnz + 1 : (int64_t)sg_ldg(rp + mid + 1);
        const int64_t cend = sg_ldg(cs + mid) + (next - sg_ldg(rp + mid)) - 1;
        if (cend >= j) hi = mid; else lo = mid + 1;
    }
    r0 = lo + 1;
}

template <typename T>
__global__ void __launch_bounds__(256) sg_refmat_mul_adjoint_kernel(T *__restrict__ B, const T *__restrict__ Y,
                                                                    const __grid_constant__ SgRefmatArgs<T> a)
{
    const int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (lin >= a.totalB) return;
    int64_t J[SG_MAX_DIMS], r0[SG_MAX_DIMS], nr[SG_MAX_DIMS], Ib[SG_MAX_DIMS];
    int64_t r = lin, nterm = 1;
    for (int d = 0; d < a.ndims; ++d) {
        J[d] = r % a.sizeB[d] + 1;
        r /= a.sizeB[d];
        if (a.row_ptr[d]) {
            int64_t r1;
            sg_column_rows<T>(a.row_ptr[d], a.col_start[d], a.sizeY[d], a.nnz[d], J[d], r0[d], r1);
            nr[d] = r1 - r0[d] + 1;
            if (nr[d] < 0) nr[d] = 0;
        } else { r0[d] = J[d]; nr[d] = 1; }
        nterm *= nr[d];
        Ib[d] = 0;
    }
    T out = T(0);
    for (int64_t q = 0; q < nterm; ++q) {
        int64_t off = 0;
        T coef = T(1);
        bool inside = true;
        for (int d = 0; d < a.ndims; ++d) {
            const int64_t I = r0[d] + Ib[d];
            off += (I - 1) * a.strideY[d];
            if (a.row_ptr[d]) {
                int64_t c0, nc, p0;
                sg_row_window(a.row_ptr[d], a.col_start[d], a.sizeY[d], a.nnz[d], I, c0, nc, p0);
                const int64_t k = J[d] - c0;
                if (k < 0 || k >= nc) inside = false;   // defensive: never true for a valid matrix
                else coef *= sg_ldg(a.nzval[d] + p0 + k - 1);
            }
        }
        if (inside) out += sg_ldg(Y + off) * coef;
        for (int d = 0; d < a.ndims; ++d) { if (++Ib[d] < nr[d]) break; Ib[d] = 0; }
    }
    B[lin] = out;
}

// K7 / K8
template <typename T>
__global__ void sg_scatter_active_kernel(T *__restrict__ cp, int nin, const __grid_constant__ SgRefmatArgs<T> a, int nout,
                                         const int32_t *__restrict__ idx, const T *__restrict__ vals, int64_t n_active)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    int64_t off = 0;
    for (int d = 0; d < nin; ++d) off += (int64_t)(sg_ldg(idx + i + n_active * d) - 1) * a.strideY[d];
    for (int o = 0; o < nout; ++o) cp[off + a.totalY * o] = sg_ldg(vals + i + n_active * o);
}

template <typename T>
__global__ void sg_gather_zero_active_kernel(T *__restrict__ vals, T *__restrict__ cp, int nin,
                                             const __grid_constant__ SgRefmatArgs<T> a, int nout,
                                             const int32_t *__restrict__ idx, int64_t n_active)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    int64_t off = 0;
    for (int d = 0; d < nin; ++d) off += (int64_t)(sg_ldg(idx + i + n_active * d) - 1) * a.strideY[d];
    for (int o = 0; o < nout; ++o) {
        vals[i + n_active * o] = cp[off + a.totalY * o];
        cp[off + a.totalY * o] = T(0);
    }
}

template <typename T>
static int sg_fill_refmat(SgRefmatArgs<T> &a, int ndims, const int64_t *sizeY, const int64_t *sizeB, int n_ref,
                          const int *dims, const int32_t *const *row_ptr, const int32_t *const *col_start,
                          const T *const *nzval, const int64_t *nnz)
{
    SG_CHECK_ARG(sizeY && sizeB && n_ref >= 0);
    if (ndims < 1 || ndims > SG_MAX_DIMS || n_ref > ndims) return SG_ERR_UNSUPPORTED;
    SG_CHECK_ARG(n_ref == 0 || (dims && row_ptr && col_start && nzval && nnz));
    a.ndims = ndims;
    a.totalY = a.totalB = 1;
    for (int d = 0; d < SG_MAX_DIMS; ++d) {
        a.sizeY[d] = a.sizeB[d] = 1; a.strideB[d] = a.strideY[d] = 0;
        a.row_ptr[d] = nullptr; a.col_start[d] = nullptr; a.nzval[d] = nullptr; a.nnz[d] = 0;
    }
    for (int d = 0; d < ndims; ++d) {
        SG_CHECK_ARG(sizeY[d] >= 1 && sizeB[d] >= 1);
        a.sizeY[d] = sizeY[d]; a.sizeB[d] = sizeB[d];
        a.strideY[d] = a.totalY; a.strideB[d] = a.totalB;
        a.totalY *= sizeY[d]; a.totalB *= sizeB[d];
    }
    for (int r = 0; r < n_ref; ++r) {
        const int d = dims[r] - 1;
        SG_CHECK_ARG(d >= 0 && d < ndims && a.row_ptr[d] == nullptr);  // unique, in range (src/validation.jl:93)
        SG_CHECK_ARG(row_ptr[r] && col_start[r] && nzval[r] && nnz[r] >= 1);
        a.row_ptr[d] = row_ptr[r]; a.col_start[d] = col_start[r]; a.nzval[d] = nzval[r]; a.nnz[d] = nnz[r];
    }
    for (int d = 0; d < ndims; ++d)
        if (!a.row_ptr[d]) SG_CHECK_ARG(sizeY[d] == sizeB[d]);          // src/validation.jl:102-104
    return SG_OK;
}

template <typename T>
static int sg_fill_cp_strides(SgRefmatArgs<T> &a, int nin, const int64_t *n_cp)
{
    SG_CHECK_ARG(n_cp);
    if (nin < 1 || nin > SG_MAX_DIMS) return SG_ERR_UNSUPPORTED;
    a.ndims = nin;
    a.totalY = 1;
    for (int d = 0; d < SG_MAX_DIMS; ++d) a.strideY[d] = 0;
    for (int d = 0; d < nin; ++d) { SG_CHECK_ARG(n_cp[d] >= 1); a.strideY[d] = a.totalY; a.totalY *= n_cp[d]; }
    return SG_OK;
}

#define SG_DEFINE_REFMAT_API(T, SUF)                                                                                 \
    extern "C" int sg_refmat_mul_##SUF(T *Y, const T *B, int ndims, const int64_t *sizeY, const int64_t *sizeB,      \
                                       int n_ref, const int *dims, const int32_t *const *row_ptr,                    \
                                       const int32_t *const *col_start, const T *const *nzval, const int64_t *nnz,   \
                                       void *stream)                                                                 \
    {                                                                                                                \
        SG_NVTX("sg_refmat_mul");                                                                                    \
        SG_CHECK_ARG(Y && B);                                                                                        \
        SgRefmatArgs<T> a;                                                                                           \
        int rc = sg_fill_refmat<T>(a, ndims, sizeY, sizeB, n_ref, dims, row_ptr, col_start, nzval, nnz);             \
        if (rc != SG_OK) return rc;                                                                                  \
        sg_refmat_mul_kernel<T><<<sg_blocks(a.totalY, 256), 256, 0, sg_stream(stream)>>>(Y, B, a);                   \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_refmat_mul_adjoint_##SUF(T *B, const T *Y, int ndims, const int64_t *sizeY,                    \
                                               const int64_t *sizeB, int n_ref, const int *dims,                     \
                                               const int32_t *const *row_ptr, const int32_t *const *col_start,       \
                                               const T *const *nzval, const int64_t *nnz, void *stream)              \
    {                                                                                                                \
        SG_NVTX("sg_refmat_mul_adjoint");                                                                            \
        SG_CHECK_ARG(Y && B);                                                                                        \
        SgRefmatArgs<T> a;                                                                                           \
        int rc = sg_fill_refmat<T>(a, ndims, sizeY, sizeB, n_ref, dims, row_ptr, col_start, nzval, nnz);             \
        if (rc != SG_OK) return rc;                                                                                  \
        sg_refmat_mul_adjoint_kernel<T><<<sg_blocks(a.totalB, 256), 256, 0, sg_stream(stream)>>>(B, Y, a);           \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_scatter_active_##SUF(T *cp, int nin, const int64_t *n_cp, int nout, const int32_t *idx,        \
                                           const T *vals, int64_t n_active, void *stream)                            \
    {                                                                                                                \
        SG_NVTX("sg_active_points");                                                                                 \
        SG_CHECK_ARG(cp && nout >= 1 && n_active >= 0);                                                              \
        if (n_active == 0) return SG_OK;                                                                             \
        SG_CHECK_ARG(idx && vals);                                                                                   \
        SgRefmatArgs<T> a;                                                                                           \
        int rc = sg_fill_cp_strides<T>(a, nin, n_cp);                                                                \
        if (rc != SG_OK) return rc;                                                                                  \
        sg_scatter_active_kernel<T><<<sg_blocks(n_active, 256), 256, 0, sg_stream(stream)>>>(cp, nin, a, nout, idx,  \
                                                                                             vals, n_active);        \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_gather_zero_active_##SUF(T *vals, T *cp, int nin, const int64_t *n_cp, int nout,               \
                                               const int32_t *idx, int64_t n_active, void *stream)                   \
    {                                                                                                                \
        SG_NVTX("sg_active_points");                                                                                 \
        SG_CHECK_ARG(cp && nout >= 1 && n_active >= 0);                                                              \
        if (n_active == 0) return SG_OK;                                                                             \
        SG_CHECK_ARG(idx && vals);                                                                                   \
        SgRefmatArgs<T> a;                                                                                           \
        int rc = sg_fill_cp_strides<T>(a, nin, n_cp);                                                                \
        if (rc != SG_OK) return rc;                                                                                  \
        sg_gather_zero_active_kernel<T><<<sg_blocks(n_active, 256), 256, 0, sg_stream(stream)>>>(vals, cp, nin, a,   \
                                                                                                 nout, idx,          \
                                                                                                 n_active);          \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }

SG_DEFINE_REFMAT_API(float, f32)
SG_DEFINE_REFMAT_API(double, f64)
