// sg_basis.cu -- per-dimension basis tables on device (K1, K2, fused K1+K2) and the O(n) set-up
// helpers K9-K12.  Replaces the KernelAbstractions kernels in src/util_kernels.jl and
// src/spline_dimension.jl:160-215 of the reference.
#include "sg_common.cuh"

// ---------------------------------------------------------------------------------------------
// K1: knot-span lookup.  The reference scans all knots and stops at the first `t < knot`
// (src/util_kernels.jl:35-41); on a sorted knot vector that count is the partition point of
// the predicate `!(t < knot)`, found here by binary search.  NaN -> every comparison false ->
// n_knots, as in the reference.  Result clamped to [p+1, n_knots-p-1] (:43-47).  1-based.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int32_t sg_find_span(T t, const T *__restrict__ knots, int64_t n_knots, int degree)
{
    int64_t lo = 0, hi = n_knots;  // first k in [lo,hi) with t < knots[k]
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (t < sg_ldg(knots + mid)) hi = mid; else lo = mid + 1;
    }
    int64_t idx = lo;
    int64_t cl = degree + 1, ch = n_knots - degree - 1;
    idx = idx < cl ? cl : idx;
    idx = idx > ch ? ch : idx;
    return (int32_t)idx;
}

template <typename T>
__global__ void sg_span_indices_kernel(int32_t *__restrict__ out, const T *__restrict__ samples, int64_t n,
                                       const T *__restrict__ knots, int64_t n_knots, int degree)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = sg_find_span(sg_ldg(samples + i), knots, n_knots, degree);
}

// ---------------------------------------------------------------------------------------------
// K2: Cox-de Boor triangle with derivative rows, one thread per sample, thread-local storage
// (W = compile-time bound on p+1 keeps the small degrees in registers).  Operation order is the
// reference's (src/spline_dimension.jl:189-207): frac = prev/dt; cur[k_] += frac*(tmax-t);
// cur[k_+1] = frac*(t-tmin); c = (prev*k)/dt.  No FMA contraction (sg_mul/sg_add).
// FUSED: compute the span in the same thread (K1) and store it.
// ---------------------------------------------------------------------------------------------
template <typename T, int W, bool FUSED>
__global__ void sg_basis_tables_kernel(T *__restrict__ eval, int32_t *__restrict__ idx_out,
                                       const int32_t *__restrict__ idx_in, const T *__restrict__ knots,
                                       int64_t n_knots, const T *__restrict__ samples, int64_t n, int p, int mdo)
{
    int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= n) return;
    const T t = sg_ldg(samples + l);
    int64_t i;
    if (FUSED) {
        int32_t s = sg_find_span(t, knots, n_knots, p);
        idx_out[l] = s;
        i = s;
    } else {
        i = sg_ldg(idx_in + l);
    }
    T cur[W][W], prev[W][W];  // [derivative][basis]
#pragma unroll(W <= 8 ? W : 1)
    for (int d = 0; d < W; ++d)
#pragma unroll(W <= 8 ? W : 1)
        for (int j = 0; j < W; ++j) { cur[d][j] = T(0); prev[d][j] = T(0); }
    cur[0][0] = T(1);
    prev[0][0] = T(1);
#pragma unroll(W <= 8 ? W : 1)
    for (int k = 1; k < W; ++k) {
        if (k <= p) {
#pragma unroll(W <= 8 ? W : 1)
            for (int d = 0; d < W; ++d)
#pragma unroll(W <= 8 ? W : 1)
                for (int j = 0; j < W; ++j) cur[d][j] = T(0);
            const int nder = mdo + k - p;  // derivative rows produced at this degree
#pragma unroll(W <= 8 ? W : 1)
            for (int k_ = 1; k_ <= k; ++k_) {
                const T t_min = sg_ldg(knots + (i + k_ - k - 1));
                const T t_max = sg_ldg(knots + (i + k_ - 1));
                const T dt = sg_sub(t_max, t_min);
                const T frac = sg_div(prev[0][k_ - 1], dt);
                cur[0][k_ - 1] = sg_add(cur[0][k_ - 1], sg_mul(frac, sg_sub(t_max, t)));
                cur[0][k_] = sg_mul(frac, sg_sub(t, t_min));
#pragma unroll(W <= 8 ? W : 1)
                for (int d = 1; d < W; ++d) {
                    if (d <= nder) {
                        const T c = sg_div(sg_mul(prev[d - 1][k_ - 1], T(k)), dt);
                        cur[d][k_ - 1] = sg_sub(cur[d][k_ - 1], c);
                        cur[d][k_] = c;
                    }
                }
            }
            if (k != p) {
#pragma unroll(W <= 8 ? W : 1)
                for (int d = 0; d < W; ++d)
#pragma unroll(W <= 8 ? W : 1)
                    for (int j = 0; j < W; ++j) prev[d][j] = cur[d][j];
            }
        }
    }
    const int w = p + 1;
#pragma unroll(W <= 8 ? W : 1)
    for (int d = 0; d < W; ++d)
#pragma unroll(W <= 8 ? W : 1)
        for (int j = 0; j < W; ++j)
            if (d <= mdo && j < w) eval[l + n * (j + (int64_t)w * d)] = cur[d][j];
}

template <typename T, bool FUSED>
static int sg_launch_basis(T *eval, int32_t *idx_out, const int32_t *idx_in, const T *knots, int64_t n_knots,
                           const T *samples, int64_t n, int p, int mdo, cudaStream_t st)
{
    SG_NVTX(FUSED ? "sg_dimension_build" : "sg_basis_tables");
    SG_CHECK_ARG(eval && knots && samples && n >= 1);
    SG_CHECK_ARG(FUSED ? idx_out != nullptr : idx_in != nullptr);
    if (p < 0 || p > SG_MAX_DEGREE) return SG_ERR_UNSUPPORTED;
    SG_CHECK_ARG(mdo >= 0 && mdo <= p);
    SG_CHECK_ARG(n_knots >= 2 * (p + 1));
    const int threads = 128;
    const unsigned blocks = sg_blocks(n, threads);
    if (p < 4)
        sg_basis_tables_kernel<T, 4, FUSED><<<blocks, threads, 0, st>>>(eval, idx_out, idx_in, knots, n_knots, samples, n, p, mdo);
    else if (p < 8)
        sg_basis_tables_kernel<T, 8, FUSED><<<blocks, threads, 0, st>>>(eval, idx_out, idx_in, knots, n_knots, samples, n, p, mdo);
    else
        sg_basis_tables_kernel<T, 16, FUSED><<<blocks, threads, 0, st>>>(eval, idx_out, idx_in, knots, n_knots, samples, n, p, mdo);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// K9-K12: O(n) set-up helpers (src/util_kernels.jl:1-20, 51-88)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void sg_expand_knots_kernel(T *__restrict__ knots_all, const T *__restrict__ values,
                                       const int32_t *__restrict__ mult, int64_t n_values)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_values) return;
    int64_t start = 0;
    for (int64_t j = 0; j < i; ++j) start += mult[j];
    const T v = values[i];
    for (int64_t k = start; k < start + mult[i]; ++k) knots_all[k] = v;
}

template <typename T>
__global__ void sg_decompress_kernel(T *__restrict__ out, const T *__restrict__ eval,
                                     const int32_t *__restrict__ idx, int64_t n, int p, int der)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t l = idx[i];  // 1-based span; columns l-p .. l (1-based)
    for (int j = 0; j <= p; ++j) out[i + n * (l - p - 1 + j)] = eval[i + n * (j + (int64_t)(p + 1) * der)];
}

template <typename T>
__global__ void sg_insert_kernel(T *__restrict__ out, const T *__restrict__ v, int64_t len, int64_t i_insert, T x)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1;  // 1-based
    if (i > len + 1) return;
    out[i - 1] = (i < i_insert) ? v[i - 1] : (i > i_insert) ? v[i - 2] : x;
}

__global__ void sg_collect_indices_kernel(int32_t *__restrict__ indices, const int64_t *__restrict__ cart, int64_t n, int nin)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int d = 0; d < nin; ++d) indices[i + n * d] = (int32_t)cart[i * nin + d];
}

template <typename T>
static int sg_insert_impl(T *out, const T *v, int64_t len, int64_t i_insert, T x, void *stream)
{
    SG_CHECK_ARG(out && (v || len == 0) && len >= 0 && i_insert >= 1 && i_insert <= len + 1);
    sg_insert_kernel<T><<<sg_blocks(len + 1, 256), 256, 0, sg_stream(stream)>>>(out, v, len, i_insert, x);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
#define SG_DEFINE_BASIS_API(T, SUF)                                                                                  \
    extern "C" int sg_expand_knot_vector_##SUF(T *knots_all, const T *values, const int32_t *mult, int64_t n_values, \
                                               void *stream)                                                         \
    {                                                                                                                \
        SG_CHECK_ARG(knots_all && values && mult && n_values >= 1);                                                  \
        sg_expand_knots_kernel<T><<<sg_blocks(n_values, 128), 128, 0, sg_stream(stream)>>>(knots_all, values, mult,  \
                                                                                           n_values);                \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_span_indices_##SUF(int32_t *out, const T *samples, int64_t n, const T *knots, int64_t n_knots, \
                                         int degree, void *stream)                                                   \
    {                                                                                                                \
        SG_CHECK_ARG(out && samples && knots && n >= 1 && degree >= 0 && n_knots >= 2 * (degree + 1));               \
        sg_span_indices_kernel<T><<<sg_blocks(n, 256), 256, 0, sg_stream(stream)>>>(out, samples, n, knots, n_knots, \
                                                                                    degree);                         \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_basis_tables_##SUF(T *eval, const T *knots, int64_t n_knots, const T *samples,                 \
                                         const int32_t *idx, int64_t n, int degree, int mdo, void *stream)           \
    {                                                                                                                \
        return sg_launch_basis<T, false>(eval, nullptr, idx, knots, n_knots, samples, n, degree, mdo,                \
                                         sg_stream(stream));                                                         \
    }                                                                                                                \
    extern "C" int sg_dimension_build_##SUF(int32_t *idx, T *eval, const T *knots, int64_t n_knots,                  \
                                            const T *samples, int64_t n, int degree, int mdo, void *stream)          \
    {                                                                                                                \
        return sg_launch_basis<T, true>(eval, idx, nullptr, knots, n_knots, samples, n, degree, mdo,                 \
                                        sg_stream(stream));                                                          \
    }                                                                                                                \
    extern "C" int sg_decompress_##SUF(T *out, const T *eval, const int32_t *idx, int64_t n, int64_t n_basis,        \
                                       int degree, int der, void *stream)                                            \
    {                                                                                                                \
        SG_CHECK_ARG(out && eval && idx && n >= 1 && n_basis >= degree + 1 && degree >= 0 && der >= 0);              \
        sg_decompress_kernel<T><<<sg_blocks(n, 256), 256, 0, sg_stream(stream)>>>(out, eval, idx, n, degree, der);   \
        SG_AFTER_LAUNCH();                                                                                           \
        return SG_OK;                                                                                                \
    }                                                                                                                \
    extern "C" int sg_insert_##SUF(T *out, const T *v, int64_t len, int64_t i_insert, T x, void *stream)             \
    {                                                                                                                \
        return sg_insert_impl<T>(out, v, len, i_insert, x, stream);                                                  \
    }

SG_DEFINE_BASIS_API(float, f32)
SG_DEFINE_BASIS_API(double, f64)

extern "C" int sg_insert_i32(int32_t *out, const int32_t *v, int64_t len, int64_t i_insert, int32_t x, void *stream)
{
    return sg_insert_impl<int32_t>(out, v, len, i_insert, x, stream);
}

extern "C" int sg_collect_indices_i32(int32_t *indices, const int64_t *cart, int64_t n, int nin, void *stream)
{
    SG_CHECK_ARG(indices && cart && n >= 1 && nin >= 1 && nin <= SG_MAX_DIMS);
    sg_collect_indices_kernel<<<sg_blocks(n, 256), 256, 0, sg_stream(stream)>>>(indices, cart, n, nin);
    SG_AFTER_LAUNCH();
    return SG_OK;
}
