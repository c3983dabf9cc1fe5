// sg_evaluate_generic.cuh -- generic (any Nin, any degree, any Nout, rational or not) forward and
// adjoint kernels: the correctness back-stop for every shape the reference's tests use.
// One thread per sample, the reference's window loop (src/spline_grid.jl:146-175,
// src/adjoint.jl:21-39) with register accumulation instead of global read-modify-writes.
#pragma once
#include "sg_common.cuh"


// Decode the linear sample index (dim 1 fastest) and compute the control-point base offset.
template <typename T>
__device__ __forceinline__ void sg_decode_sample(const SgGridArgs<T> &a, int64_t lin, int64_t *J, int64_t &base)
{
    base = 0;
    int64_t r = lin;
#pragma unroll
    for (int d = 0; d < SG_MAX_DIMS; ++d) {
        if (d < a.nin) {
            J[d] = r % a.n_samples[d];
            r /= a.n_samples[d];
            base += (int64_t)(sg_ldg(a.index[d] + J[d]) - a.degree[d] - 1) * a.cp_stride[d];
        }
    }
}

// Basis product and control-point offset of window term I (accumulation order = reference:
// ((1*B1)*B2)..., src/spline_grid.jl:151-156).
template <typename T>
__device__ __forceinline__ T sg_window_term(const SgGridArgs<T> &a, const int64_t *J, const int *I, int64_t base,
                                            int64_t &off)
{
    T prod = T(1);
    off = base;
#pragma unroll
    for (int d = 0; d < SG_MAX_DIMS; ++d) {
        if (d < a.nin) {
            prod *= sg_ldg(a.table[d] + J[d] + a.n_samples[d] * I[d]);
            off += I[d] * a.cp_stride[d];
        }
    }
    return prod;
}

__device__ __forceinline__ void sg_next_offset(int nin, const int *degree, int *I)
{
#pragma unroll
    for (int d = 0; d < SG_MAX_DIMS; ++d) {
        if (d < nin) {
            if (++I[d] <= degree[d]) return;
            I[d] = 0;
        }
    }
}

template <typename T, bool NURBS>
__global__ void __launch_bounds__(256) sg_evaluate_generic_kernel(T *__restrict__ eval, const __grid_constant__ SgGridArgs<T> a,
                                                                  const T *__restrict__ cp, const T *__restrict__ weights)
{
    const int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (lin >= a.n_total) return;
    int64_t J[SG_MAX_DIMS], base;
    sg_decode_sample(a, lin, J, base);
    T denom = T(0);
    for (int o0 = 0; o0 < a.nout; o0 += SG_GEN_OCHUNK) {
        T acc[SG_GEN_OCHUNK];
#pragma unroll
        for (int q = 0; q < SG_GEN_OCHUNK; ++q) acc[q] = T(0);
        int I[SG_MAX_DIMS] = {0};
        T den = T(0);
        for (int64_t w = 0; w < a.n_window; ++w) {
            int64_t off;
            T b = sg_window_term(a, J, I, base, off);
            if (NURBS) {
                b *= sg_ldg(weights + off);
                den += b;
            }
#pragma unroll
            for (int q = 0; q < SG_GEN_OCHUNK; ++q)
                if (o0 + q < a.nout) acc[q] += b * sg_ldg(cp + off + a.cp_total * (o0 + q));
            sg_next_offset(a.nin, a.degree, I);
        }
        if (NURBS) denom = den;
#pragma unroll
        for (int q = 0; q < SG_GEN_OCHUNK; ++q)
            if (o0 + q < a.nout) eval[lin + a.n_total * (o0 + q)] = NURBS ? acc[q] / denom : acc[q];
    }
}

// Atomic scatter adjoint: the reference's algorithm (src/adjoint.jl:21-39).  Used only when the
// span indices are not monotone (user-built SplineDimension) or the grid is tiny.
template <typename T, bool NURBS>
__global__ void __launch_bounds__(256) sg_adjoint_scatter_kernel(T *__restrict__ cp, const __grid_constant__ SgGridArgs<T> a,
                                                                 const SgAdjointHeader *hdr_or_null,
                                                                 const T *__restrict__ eval, const T *__restrict__ weights)
{
    // With a header: run only if the prep kernel found non-monotone span indices.
    if (hdr_or_null != nullptr && !hdr_or_null->nonmonotone) return;
    // grid-stride loop: as a (normally idle) device-side fallback it is launched with a small, fixed grid
    for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < a.n_total;
         lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t J[SG_MAX_DIMS], base;
    sg_decode_sample(a, lin, J, base);
    T inv_denom = T(1);
    if (NURBS) {
        T den = T(0);
        int I[SG_MAX_DIMS] = {0};
        for (int64_t w = 0; w < a.n_window; ++w) {
            int64_t off;
            T b = sg_window_term(a, J, I, base, off);
            den += b * sg_ldg(weights + off);
            sg_next_offset(a.nin, a.degree, I);
        }
        inv_denom = T(1) / den;
    }
    for (int o0 = 0; o0 < a.nout; o0 += SG_GEN_OCHUNK) {
        T e[SG_GEN_OCHUNK];
#pragma unroll
        for (int q = 0; q < SG_GEN_OCHUNK; ++q) e[q] = (o0 + q < a.nout) ? sg_ldg(eval + lin + a.n_total * (o0 + q)) : T(0);
        int I[SG_MAX_DIMS] = {0};
        for (int64_t w = 0; w < a.n_window; ++w) {
            int64_t off;
            T b = sg_window_term(a, J, I, base, off);
            if (NURBS) b = b * sg_ldg(weights + off) * inv_denom;
#pragma unroll
            for (int q = 0; q < SG_GEN_OCHUNK; ++q)
                if (o0 + q < a.nout) atomicAdd(cp + off + a.cp_total * (o0 + q), b * e[q]);
            sg_next_offset(a.nin, a.degree, I);
        }
    }
    }
}
