// sg_adjoint_generic.cuh -- atomics-free generic adjoint: gather form over per-control-point
// sample ranges.  Requires non-decreasing span indices per dimension (checked ON DEVICE by the
// prep kernel, no host synchronisation): with monotone spans the samples whose window contains
// control index i form one contiguous range per dimension.
#pragma once
#include <climits>
#include "sg_common.cuh"

// span_start[d][s] (s = 0..n_cp+1, 1-based span s) = first sample j with index[j] >= s
// (lower bound), so the samples of span s are [span_start[s], span_start[s+1]).
template <typename T>
struct SgSpanStarts {
    int32_t *start[SG_MAX_DIMS];
    // gather table of dimension 1 (used by the double march's post kernel): for every control index i the first sample
    // of its support, the number of samples, and the first SG_GATHER_RMAX basis weights B1[lo + r, i - span + p]
    int32_t *g_lo;     // [c_1][2] = (lo, len)
    T *g_w;            // [SG_GATHER_RMAX][c_1]
    // per-column-block tables of dimension 1 for the fused double march (sg_adjoint_march2g.cuh); bt_hdr == nullptr: none
    SgM2gBlockHdr *bt_hdr;   // [nb1]
    int32_t *bt_lol;         // [nb1][icap]  (first sample of the support relative to the block | number of samples << 16)
    T *bt_w;                 // [nb1][rmcap][icap] gather weights B1[lo + r, i - span + p], zero beyond the support
    int icap, rmcap, nb1;
    int bw, rfast;           // samples per column block; weights layout: 1 = [block][li][r], 0 = [block][r][li]
};
#define SG_GATHER_RMAX 20

template <typename T>
__global__ void sg_adjoint_prep_kernel(const __grid_constant__ SgGridArgs<T> a, const __grid_constant__ SgSpanStarts<T> ss,
                                       SgAdjointHeader *hdr)
{
    if ((int)blockIdx.y == a.nin + 1) {
        // second extra row of blocks: the column-block tables of the fused double march.  One CUDA block per column
        // block of 128 samples (grid-stride); a thread owns one local control index.
        if (ss.bt_hdr == nullptr) return;
        __shared__ int s_ilo, s_ni, s_rm;
        const int64_t n0 = a.n_samples[0], c0 = a.n_cp[0];
        const int p0 = a.degree[0];
        const int32_t *__restrict__ idx0 = a.index[0];
        for (int jb = blockIdx.x; jb < ss.nb1; jb += gridDim.x) {
            const int64_t j_lo = (int64_t)jb * ss.bw, j_hi = min(j_lo + ss.bw, n0);
            __syncthreads();
            if (threadIdx.x == 0) {
                s_ilo = idx0[j_lo] - p0;
                s_ni = idx0[j_hi - 1] - s_ilo + 1;
                s_rm = 0;
            }
            __syncthreads();
            const int ilo = s_ilo, ni = s_ni;
            const bool fits = ni >= 1 && ni <= ss.icap;
            for (int li = threadIdx.x; li < min(ni, ss.icap); li += blockDim.x) {
                const int64_t i = (int64_t)ilo + li;                    // 1-based control index
                const int64_t s0 = i > p0 + 1 ? i : p0 + 1, s1 = (i + p0 < c0 ? i + p0 : c0) + 1;
                int64_t lo = j_lo, hi = j_hi;                           // first sample of the block with span >= s0
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (idx0[mid] >= s0) hi = mid; else lo = mid + 1; }
                const int64_t first = lo;
                hi = j_hi;                                              // first sample of the block with span >= s1
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (idx0[mid] >= s1) hi = mid; else lo = mid + 1; }
                const int len = (int)(lo - first);
                atomicMax(&s_rm, len);
                const int len_w = min(len, ss.rmcap);
                ss.bt_lol[(int64_t)jb * ss.icap + li] = (int32_t)(first - j_lo) | (len_w << 16);
                T *__restrict__ wp = ss.bt_w + (int64_t)jb * ss.rmcap * ss.icap + (ss.rfast ? (int64_t)li * ss.rmcap : (int64_t)li);
                const int64_t wstride = ss.rfast ? 1 : ss.icap;
                for (int r = 0; r < ss.rmcap; ++r) {
                    T w = T(0);
                    if (r < len_w) {
                        const int k = min(max((int)(i - idx0[first + r] + p0), 0), p0);   // clamp: garbage-safe
                        w = a.table[0][first + r + n0 * k];
                    }
                    wp[(int64_t)r * wstride] = w;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                SgM2gBlockHdr bh;
                bh.i1_lo = ilo; bh.ni = fits ? ni : 1; bh.rm = s_rm; bh.pad = 0;
                ss.bt_hdr[jb] = bh;
                if (!fits || s_rm > ss.rmcap) hdr->m2g_bad = 1;
            }
        }
        return;
    }
    if ((int)blockIdx.y == a.nin) {
        // extra row of blocks: the gather table of dimension 1 (binary searches of its own: no dependence on start[])
        if (ss.g_lo == nullptr) return;
        const int64_t n0 = a.n_samples[0], c0 = a.n_cp[0];
        const int p0 = a.degree[0];
        const int32_t *__restrict__ idx0 = a.index[0];
        for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < c0; i0 += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i = i0 + 1;
            const int64_t s0 = i > p0 + 1 ? i : p0 + 1, s1 = (i + p0 < c0 ? i + p0 : c0) + 1;
            int64_t lo = 0, hi = n0;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (idx0[mid] >= s0) hi = mid; else lo = mid + 1; }
            const int64_t first = lo;
            // the spans of the next SG_GATHER_RMAX samples (independent loads) give the weights' columns and, for
            // monotone spans, the length of the support; only longer supports need the second binary search
            int sp[SG_GATHER_RMAX];
#pragma unroll
            for (int r = 0; r < SG_GATHER_RMAX; ++r) sp[r] = first + r < n0 ? idx0[first + r] : INT_MAX;
            int len = 0;
#pragma unroll
            for (int r = 0; r < SG_GATHER_RMAX; ++r) len += sp[r] < s1 ? 1 : 0;
            if (len == SG_GATHER_RMAX) {
                hi = n0;
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (idx0[mid] >= s1) hi = mid; else lo = mid + 1; }
                len = (int)(lo - first);
            }
            ss.g_lo[2 * i0] = (int32_t)first;
            ss.g_lo[2 * i0 + 1] = len;
            T w[SG_GATHER_RMAX];
#pragma unroll
            for (int r = 0; r < SG_GATHER_RMAX; ++r) {
                const int k = min(max((int)(i - sp[r] + p0), 0), p0);       // clamp: garbage-safe for non-monotone spans
                w[r] = r < len ? a.table[0][first + r + n0 * k] : T(0);
            }
#pragma unroll
            for (int r = 0; r < SG_GATHER_RMAX; ++r) ss.g_w[(int64_t)r * c0 + i0] = w[r];
        }
        return;
    }
    const int d = blockIdx.y;
    const int64_t n = a.n_samples[d];
    const int32_t *__restrict__ idx = a.index[d];
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = tid; j + 1 < n; j += stride)
        if (idx[j] > idx[j + 1]) hdr->nonmonotone = 1;
    if (tid == 0) {
        hdr->span_first[d] = idx[0];
        hdr->span_last[d] = idx[n - 1];
    }
    int rows_max = 0;
    for (int64_t s = tid; s <= a.n_cp[d] + 1; s += stride) {
        int64_t lo = 0, hi = n;  // first j with idx[j] >= s
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (idx[mid] >= s) hi = mid; else lo = mid + 1;
        }
        ss.start[d][s] = (int32_t)lo;
        if (d == 1) {                                                   // samples in span s (3-D double march: ring row slots)
            int64_t l2 = lo;
            hi = n;
            while (l2 < hi) { const int64_t mid = (l2 + hi) >> 1; if (idx[mid] >= s + 1) hi = mid; else l2 = mid + 1; }
            rows_max = max(rows_max, (int)(l2 - lo));
        }
    }
    if (d == 1 && rows_max > 0) atomicMax(&hdr->rows2_max, rows_max);
}

// One thread per control point; gathers prod_d B_d[J_d, i_d - span(J_d) + p_d] * eval[J, o]
// over J in prod_d [lo_d, hi_d).  RATIONAL: eval is pre-divided by denom[J] and the result
// multiplied by w[i] (transpose of the fixed-weights rational map).
template <typename T, bool RATIONAL>
__global__ void __launch_bounds__(128) sg_adjoint_gather_kernel(T *__restrict__ cp, const __grid_constant__ SgGridArgs<T> a,
                                                                const __grid_constant__ SgSpanStarts<T> ss,
                                                                const SgAdjointHeader *hdr, const T *__restrict__ eval,
                                                                const T *__restrict__ weights, const T *__restrict__ denom)
{
    if (hdr->nonmonotone) return;
    const int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (lin >= a.cp_total) return;
    int64_t i[SG_MAX_DIMS], lo[SG_MAX_DIMS], hi[SG_MAX_DIMS], J[SG_MAX_DIMS];
    int64_t r = lin, n_terms = 1;
    for (int d = 0; d < a.nin; ++d) {
        i[d] = r % a.n_cp[d] + 1;  // 1-based control index
        r /= a.n_cp[d];
        const int p = a.degree[d];
        int64_t s0 = i[d] > p + 1 ? i[d] : p + 1;
        int64_t s1 = i[d] + p < a.n_cp[d] ? i[d] + p : a.n_cp[d];
        lo[d] = ss.start[d][s0];
        hi[d] = ss.start[d][s1 + 1];
        J[d] = lo[d];
        n_terms *= (hi[d] - lo[d]);
    }
    for (int o0 = 0; o0 < a.nout; o0 += SG_GEN_OCHUNK) {
        T acc[SG_GEN_OCHUNK];
#pragma unroll
        for (int q = 0; q < SG_GEN_OCHUNK; ++q) acc[q] = T(0);
        for (int d = 0; d < a.nin; ++d) J[d] = lo[d];
        for (int64_t t = 0; t < n_terms; ++t) {
            T b = T(1);
            int64_t off = 0, st = 1;
            for (int d = 0; d < a.nin; ++d) {
                const int I = (int)(i[d] - sg_ldg(a.index[d] + J[d]) + a.degree[d]);
                b *= sg_ldg(a.table[d] + J[d] + a.n_samples[d] * I);
                off += J[d] * st;
                st *= a.n_samples[d];
            }
            if (RATIONAL) b /= sg_ldg(denom + off);
#pragma unroll
            for (int q = 0; q < SG_GEN_OCHUNK; ++q)
                if (o0 + q < a.nout) acc[q] += b * sg_ldg(eval + off + a.n_total * (o0 + q));
            for (int d = 0; d < a.nin; ++d) {
                if (++J[d] < hi[d]) break;
                J[d] = lo[d];
            }
        }
        const T w = RATIONAL ? sg_ldg(weights + lin) : T(1);
#pragma unroll
        for (int q = 0; q < SG_GEN_OCHUNK; ++q)
            if (o0 + q < a.nout) cp[lin + a.cp_total * (o0 + q)] = RATIONAL ? acc[q] * w : acc[q];
    }
}
