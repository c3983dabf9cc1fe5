// sg_fast_eval.cuh -- tiled "march" kernels for evaluate! (K3), 2-D and 3-D, uniform degree P.
//
// evaluate! is recast as separable mode-n contractions.  A thread owns a small block of sample COLUMNS
// (V1 consecutive samples along dimension 1 [x V2 along dimension 2 in 3-D]) and MARCHES along the slowest
// sample axis.  For its columns it keeps, in registers, the partial contraction over all but the marching
// dimension for the P+1 control planes/rows the current knot span needs:
//     3-D:  T2[v1][v2][k] = sum_{a,b} B1[j1,a] B2[j2,b] cp[i1+a, i2+b, i3(k)]
//     2-D:  T1[v1][k][ch] = sum_a     B1[j1,a]          cp[i1+a, i2(k), ch]
// Each output then costs only P+1 FMAs:  out = sum_k B_L[j_L,k] * T[..][k].  When the marching index
// crosses a knot span the window shifts by one and ONE new plane/row is contracted (the control points
// come from L1/L2: a CTA touches a few hundred bytes of them per plane).  Output stores are 128-bit and
// each warp writes 512 contiguous bytes.  The marching axis' table rows and span indices are staged once
// per CTA in shared memory (broadcast reads).
//
// Generality: the columns of one thread may straddle knot spans; their basis weights are expanded to a
// common, zero-padded window of width P+1+E (E = 1).  If a thread's columns straddle more than E spans
// (fewer samples than spans, unsorted samples) that thread alone takes a slow, direct path, so the
// kernels are correct for ANY span indices.  Reference semantics: src/spline_grid.jl:119-183.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "sg_common.cuh"

// ---- TMA / mbarrier primitives (PTX; sm_90+ async proxy, used here on sm_100a) -----------------------
#define SG_TMA_B1 24   // control-tile box: dimension 1 (elements)
#define SG_TMA_B2 12   //                   dimension 2
#define SG_TMA_B3 32   //                   planes of the marching dimension
__device__ __forceinline__ uint32_t sg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sg_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void sg_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sg_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONE_%=;\n"
        "bra SG_WAIT_%=;\n"
        "SG_DONE_%=:\n"
        "}\n" ::"r"(sg_smem_u32(bar)), "r"(parity) : "memory");
}
// 4-D tiled bulk tensor load global -> shared, completion signalled on the mbarrier (cp.async.bulk.tensor = TMA)
__device__ __forceinline__ void sg_tma_load_4d(void *dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(sg_smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(sg_smem_u32(bar))
        : "memory");
}

template <typename T, int N> struct SgVecT;
template <> struct SgVecT<float, 4> { using type = float4; };
template <> struct SgVecT<float, 2> { using type = float2; };
template <> struct SgVecT<double, 2> { using type = double2; };
template <> struct SgVecT<float, 1> { using type = float; };
template <> struct SgVecT<double, 1> { using type = double; };

// store V consecutive values; vectorised when allowed (alignment checked by the dispatcher)
template <typename T, int V>
__device__ __forceinline__ void sg_store_vec(T *p, const T (&v)[V], bool vec_ok, int n_valid)
{
    if (vec_ok && n_valid == V) {
        typename SgVecT<T, V>::type pk;
        T *q = reinterpret_cast<T *>(&pk);
#pragma unroll
        for (int i = 0; i < V; ++i) q[i] = v[i];
        __stcs(reinterpret_cast<typename SgVecT<T, V>::type *>(p), pk);   // streaming: written once, never re-read
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (i < n_valid) __stcs(p + i, v[i]);
    }
}

// Direct evaluation of one sample (all window terms) -- slow path for irregular threads.
template <typename T, bool NURBS>
__device__ __noinline__ T sg_eval_point_slow(const SgGridArgs<T> &a, const int64_t *J, const T *__restrict__ cp,
                                             const T *__restrict__ weights, int o)
{
    int64_t base = 0;
    for (int d = 0; d < a.nin; ++d) base += (int64_t)(sg_ldg(a.index[d] + J[d]) - a.degree[d] - 1) * a.cp_stride[d];
    int I[SG_MAX_DIMS] = {0};
    T acc = T(0), den = T(0);
    for (int64_t w = 0; w < a.n_window; ++w) {
        T b = T(1);
        int64_t off = base;
        for (int d = 0; d < a.nin; ++d) {
            b *= sg_ldg(a.table[d] + J[d] + a.n_samples[d] * I[d]);
            off += I[d] * a.cp_stride[d];
        }
        if (NURBS && b != T(0)) { b *= sg_ldg(weights + off); den += b; }
        if (b != T(0)) acc += b * sg_ldg(cp + off + a.cp_total * o);   // (zero weight: padded entry of a degree-padded table, maybe out of range)
        for (int d = 0; d < a.nin; ++d) { if (++I[d] <= a.degree[d]) break; I[d] = 0; }
    }
    return NURBS ? acc / den : acc;
}

// Degree padding (mixed degrees on the uniform-degree march kernels): dst (n, pmax+1) = src (n, p+1) with pmax - p leading
// zero columns.  The window of span idx then starts at idx - pmax - 1 for every dimension (it may reach below control index 0,
// where the weight is zero and the kernels clamp the address).
template <typename T>
__global__ void sg_pad_table_kernel(T *__restrict__ dst, const T *__restrict__ src, int64_t n, int p, int pmax)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int lead = pmax - p;
    for (int k = 0; k <= pmax; ++k) dst[j + n * k] = k >= lead ? src[j + n * (k - lead)] : T(0);
}

// Expanded, zero-padded weights of V consecutive samples of one dimension.
// W[v][a'] multiplies control index (minbase + a'), a' in [0, P+E].  Returns false if irregular.
template <typename T, int P, int V, int E>
__device__ __forceinline__ bool sg_expand_weights(const T *__restrict__ table, const int32_t *__restrict__ index,
                                                  int64_t n, int64_t j0, T (&W)[V][P + 1 + E], int &minbase)
{
    int base[V];
    minbase = 0x7fffffff;
    int maxbase = -1;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int64_t j = (j0 + v < n) ? j0 + v : n - 1;
        base[v] = sg_ldg(index + j) - P - 1;
        minbase = min(minbase, base[v]);
        maxbase = max(maxbase, base[v]);
    }
    const bool regular = (maxbase - minbase) <= E;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int64_t j = (j0 + v < n) ? j0 + v : n - 1;
        const int off = regular ? base[v] - minbase : 0;
#pragma unroll
        for (int ap = 0; ap < P + 1 + E; ++ap) {
            const int k = ap - off;
            W[v][ap] = (k >= 0 && k <= P) ? sg_ldg(table + j + n * (int64_t)k) : T(0);
        }
    }
    return regular;
}

// =============================================================================================
// 3-D
// =============================================================================================
// TMA = true: the CTA's whole control-point window (SG_TMA_B1 x SG_TMA_B2 control points x SG_TMA_B3 planes, a 4-D
// tensor-map box, zero-filled out of bounds) is fetched into shared memory by ONE bulk tensor copy per output
// channel, signalled on an mbarrier; plane contractions then read shared memory instead of L2.  Threads (or
// planes) whose window does not fit the box transparently use the global-load path.
template <typename T, int P, int V1, int V2, int TY, bool TMA>
__global__ void __launch_bounds__(32 * TY) sg_eval3d_march_kernel(T *__restrict__ eval, const __grid_constant__ SgGridArgs<T> a,
                                                                  const T *__restrict__ cp, int chunk, int o_count, bool vec_ok,
                                                                  const __grid_constant__ CUtensorMap tmap)
{
    constexpr int E = 1;
    constexpr int WD = P + 1 + E;
    extern __shared__ __align__(128) unsigned char sg_smem[];
    constexpr size_t TILE_BYTES = TMA ? sizeof(T) * SG_TMA_B1 * SG_TMA_B2 * SG_TMA_B3 : 0;
    T *tile = reinterpret_cast<T *>(sg_smem);                         // [B3][B2][B1] (TMA only)
    T *b3s = reinterpret_cast<T *>(sg_smem + TILE_BYTES);             // [chunk][P+1]
    int *s3s = reinterpret_cast<int *>(b3s + (size_t)chunk * (P + 1));   // [chunk]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sg_smem + TILE_BYTES + (((size_t)chunk * ((P + 1) * sizeof(T) + sizeof(int)) + 15) & ~(size_t)15));

    const int64_t n1 = a.n_samples[0], n2 = a.n_samples[1], n3 = a.n_samples[2];
    const int64_t c1 = a.n_cp[0], c2 = a.n_cp[1];
    const int64_t j1_0 = ((int64_t)blockIdx.x * 32 + threadIdx.x) * V1;
    const int64_t j2_0 = ((int64_t)blockIdx.y * TY + threadIdx.y) * V2;
    const int64_t j3_lo = (int64_t)blockIdx.z * chunk;
    const int nstep = (int)min((int64_t)chunk, n3 - j3_lo);

    // stage the marching axis' table rows and span indices (coalesced reads, broadcast LDS later)
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int s = tid; s < nstep; s += 32 * TY) {
        s3s[s] = sg_ldg(a.index[2] + j3_lo + s);
#pragma unroll
        for (int k = 0; k <= P; ++k) b3s[s * (P + 1) + k] = sg_ldg(a.table[2] + j3_lo + s + n3 * k);
    }
    // origin of the CTA's control window: the first sample of the tile in each dimension
    int o1 = 0, o2 = 0, o3 = 0;
    if (TMA) {
        // TMA needs a 16-byte aligned start address: round the innermost coordinate down
        o1 = (sg_ldg(a.index[0] + min((int64_t)blockIdx.x * 32 * V1, n1 - 1)) - P - 1) & ~(int)(16 / sizeof(T) - 1);
        o2 = sg_ldg(a.index[1] + min((int64_t)blockIdx.y * TY * V2, n2 - 1)) - P - 1;
        o3 = sg_ldg(a.index[2] + j3_lo) - P - 1;
        if (tid == 0) {
            sg_mbar_init(mbar, 1);
            sg_mbar_expect_tx(mbar, (unsigned)TILE_BYTES);
            sg_tma_load_4d(tile, &tmap, o1, o2, o3, 0, mbar);
        }
    }
    __syncthreads();
    // NOTE: no thread may exit before the last barrier / mbarrier use when TMA is on
    const bool in_range = !(j1_0 >= n1 || j2_0 >= n2);
    if (!TMA && !in_range) return;

    T W1[V1][WD], W2[V2][WD];
    int min1, min2;
    const bool reg1 = sg_expand_weights<T, P, V1, E>(a.table[0], a.index[0], n1, j1_0, W1, min1);
    const bool reg2 = sg_expand_weights<T, P, V2, E>(a.table[1], a.index[1], n2, j2_0, W2, min2);
    const int nv1 = (int)min((int64_t)V1, n1 - j1_0);
    const int nv2 = (int)min((int64_t)V2, n2 - j2_0);

    // does this thread's padded window fit the staged tile?
    const bool fits12 = TMA && (min1 - o1 >= 0) && (min1 - o1 + WD <= SG_TMA_B1) && (min2 - o2 >= 0) && (min2 - o2 + WD <= SG_TMA_B2);
    const T *__restrict__ tbase = tile + (min1 - o1) + SG_TMA_B1 * (min2 - o2);

    for (int o = 0; o < o_count; ++o) {
        const T *__restrict__ cpo = cp + a.cp_total * o;
        T *__restrict__ evo = eval + a.n_total * o;
        if (TMA) {
            if (o > 0) {   // next output channel: refill the tile (everybody is done reading the previous one)
                __syncthreads();
                if (tid == 0) {
                    sg_mbar_expect_tx(mbar, (unsigned)TILE_BYTES);
                    sg_tma_load_4d(tile, &tmap, o1, o2, o3, o, mbar);
                }
            }
            sg_mbar_wait(mbar, (unsigned)(o & 1));
            if (!in_range) continue;
        }
        if (!(reg1 && reg2)) {   // irregular thread: direct evaluation of every sample it owns
            for (int s = 0; s < nstep; ++s)
                for (int v2 = 0; v2 < nv2; ++v2)
                    for (int v1 = 0; v1 < nv1; ++v1) {
                        int64_t J[SG_MAX_DIMS] = {j1_0 + v1, j2_0 + v2, j3_lo + s};
                        evo[J[0] + n1 * (J[1] + n2 * J[2])] = sg_eval_point_slow<T, false>(a, J, cp, nullptr, o);
                    }
            continue;
        }
        // clamped control offsets of the padded windows
        int64_t col1[WD], row2[WD];
#pragma unroll
        for (int q = 0; q < WD; ++q) {
            col1[q] = min(max((int64_t)min1 + q, (int64_t)0), c1 - 1);      // (below 0 only for degree-padded tables: zero weight)
            row2[q] = min(max((int64_t)min2 + q, (int64_t)0), c2 - 1) * c1;
        }
        T T2[V1][V2][P + 1] = {};
        int cur = -0x40000000;

        auto contract_plane = [&](int64_t i3, T (&out)[V1][V2]) {
            const T *__restrict__ pl = cpo + max(i3, (int64_t)0) * c1 * c2;
            const int p3 = (int)i3 - o3;
            const bool from_tile = TMA && fits12 && p3 >= 0 && p3 < SG_TMA_B3;
            const T *__restrict__ tp = tbase + SG_TMA_B1 * SG_TMA_B2 * (from_tile ? p3 : 0);
#pragma unroll
            for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
                for (int v2 = 0; v2 < V2; ++v2) out[v1][v2] = T(0);
#pragma unroll
            for (int bq = 0; bq < WD; ++bq) {
                T c[WD];
                if (from_tile) {
#pragma unroll
                    for (int aq = 0; aq < WD; ++aq) c[aq] = tp[aq + SG_TMA_B1 * bq];
                } else {
#pragma unroll
                    for (int aq = 0; aq < WD; ++aq) c[aq] = sg_ldg(pl + row2[bq] + col1[aq]);
                }
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1) {
                    T t1 = T(0);
#pragma unroll
                    for (int aq = 0; aq < WD; ++aq) t1 = fma(W1[v1][aq], c[aq], t1);
#pragma unroll
                    for (int v2 = 0; v2 < V2; ++v2) out[v1][v2] = fma(W2[v2][bq], t1, out[v1][v2]);
                }
            }
        };

        // running pointers: one 64-bit add per step instead of re-deriving addresses
        T *__restrict__ dst = evo + j1_0 + n1 * (j2_0 + n2 * j3_lo);
        const int64_t plane_stride = n1 * n2;
        const bool full_tile = vec_ok && nv1 == V1 && nv2 == V2;
        const T *__restrict__ brow = b3s;
        for (int s = 0; s < nstep; ++s, dst += plane_stride, brow += (P + 1)) {
            const int s3 = s3s[s];   // 1-based span of this step (CTA-uniform)
            if (s3 != cur) {
                // the window slides by one plane (span advanced by one) or is rebuilt by P+1 slides
                const int nslide = (s3 == cur + 1) ? 1 : P + 1;
#pragma unroll 1
                for (int q = nslide - 1; q >= 0; --q) {
#pragma unroll
                    for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
                        for (int v2 = 0; v2 < V2; ++v2)
#pragma unroll
                            for (int k = 0; k < P; ++k) T2[v1][v2][k] = T2[v1][v2][k + 1];
                    T np_[V1][V2];
                    contract_plane((int64_t)s3 - 1 - q, np_);
#pragma unroll
                    for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
                        for (int v2 = 0; v2 < V2; ++v2) T2[v1][v2][P] = np_[v1][v2];
                }
                cur = s3;
            }
            T b3[P + 1];
#pragma unroll
            for (int k = 0; k <= P; ++k) b3[k] = brow[k];
            T outv[V2][V1];
#pragma unroll
            for (int v2 = 0; v2 < V2; ++v2)
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1) outv[v2][v1] = b3[0] * T2[v1][v2][0];
#pragma unroll
            for (int k = 1; k <= P; ++k)
#pragma unroll
                for (int v2 = 0; v2 < V2; ++v2)
#pragma unroll
                    for (int v1 = 0; v1 < V1; ++v1) outv[v2][v1] = fma(b3[k], T2[v1][v2][k], outv[v2][v1]);
            if (full_tile) {
#pragma unroll
                for (int v2 = 0; v2 < V2; ++v2) sg_store_vec<T, V1>(dst + n1 * v2, outv[v2], true, V1);
            } else {
#pragma unroll
                for (int v2 = 0; v2 < V2; ++v2)
                    if (v2 < nv2) sg_store_vec<T, V1>(dst + n1 * v2, outv[v2], false, nv1);
            }
        }
    }
}

// =============================================================================================
// 2-D (optionally rational).  NT = outputs handled per pass (<= 4); grid.z walks output tiles.
// =============================================================================================
template <typename T, int P, int V1, int NT, bool NURBS>
__global__ void __launch_bounds__(128) sg_eval2d_march_kernel(T *__restrict__ eval, const __grid_constant__ SgGridArgs<T> a,
                                                              const T *__restrict__ cp, const T *__restrict__ weights,
                                                              int chunk, bool vec_ok)
{
    constexpr int E = 1;
    constexpr int WD = P + 1 + E;
    constexpr int NCH = NT + (NURBS ? 1 : 0);
    extern __shared__ __align__(16) unsigned char sg_smem[];
    T *b2s = reinterpret_cast<T *>(sg_smem);                          // [chunk][P+1]
    int *s2s = reinterpret_cast<int *>(b2s + (size_t)chunk * (P + 1));    // [chunk]

    const int64_t n1 = a.n_samples[0], n2 = a.n_samples[1];
    const int64_t c1 = a.n_cp[0];
    const int64_t j1_0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V1;
    const int64_t j2_lo = (int64_t)blockIdx.y * chunk;
    const int nstep = (int)min((int64_t)chunk, n2 - j2_lo);
    const int o0 = blockIdx.z * NT;
    const int no = min(NT, a.nout - o0);

    for (int s = threadIdx.x; s < nstep; s += blockDim.x) {
        s2s[s] = sg_ldg(a.index[1] + j2_lo + s);
#pragma unroll
        for (int k = 0; k <= P; ++k) b2s[s * (P + 1) + k] = sg_ldg(a.table[1] + j2_lo + s + n2 * k);
    }
    __syncthreads();
    if (j1_0 >= n1) return;

    T W1[V1][WD];
    int min1;
    const bool reg1 = sg_expand_weights<T, P, V1, E>(a.table[0], a.index[0], n1, j1_0, W1, min1);
    const int nv1 = (int)min((int64_t)V1, n1 - j1_0);
    if (!reg1) {
        for (int s = 0; s < nstep; ++s)
            for (int q = 0; q < no; ++q)
                for (int v1 = 0; v1 < nv1; ++v1) {
                    int64_t J[SG_MAX_DIMS] = {j1_0 + v1, j2_lo + s};
                    eval[J[0] + n1 * J[1] + a.n_total * (o0 + q)] = sg_eval_point_slow<T, NURBS>(a, J, cp, weights, o0 + q);
                }
        return;
    }
    int64_t col1[WD];
#pragma unroll
    for (int q = 0; q < WD; ++q) col1[q] = min(max((int64_t)min1 + q, (int64_t)0), c1 - 1);

    T T1[V1][P + 1][NCH] = {};
    int cur = -0x40000000;

    auto contract_row = [&](int64_t i2, T (&out)[V1][NCH]) {
        i2 = max(i2, (int64_t)0);                                      // (degree-padded tables: zero weight)
#pragma unroll
        for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) out[v1][ch] = T(0);
        T wv[WD];
        if (NURBS) {
#pragma unroll
            for (int aq = 0; aq < WD; ++aq) wv[aq] = sg_ldg(weights + i2 * c1 + col1[aq]);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const bool is_w = NURBS && ch == NCH - 1;
            const int o = min(o0 + ch, a.nout - 1);
#pragma unroll
            for (int aq = 0; aq < WD; ++aq) {
                T c;
                if (is_w) c = wv[aq];
                else {
                    c = sg_ldg(cp + a.cp_total * o + i2 * c1 + col1[aq]);
                    if (NURBS) c *= wv[aq];
                }
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1) out[v1][ch] = fma(W1[v1][aq], c, out[v1][ch]);
            }
        }
    };

    T *__restrict__ dst = eval + j1_0 + n1 * j2_lo + a.n_total * o0;
    const bool full_tile = vec_ok && nv1 == V1 && no == NT;
    const T *__restrict__ brow = b2s;
    for (int s = 0; s < nstep; ++s, dst += n1, brow += (P + 1)) {
        const int s2 = s2s[s];
        if (s2 != cur) {
            const int nslide = (s2 == cur + 1) ? 1 : P + 1;
#pragma unroll 1
            for (int q = nslide - 1; q >= 0; --q) {
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
                    for (int k = 0; k < P; ++k)
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) T1[v1][k][ch] = T1[v1][k + 1][ch];
                T nr[V1][NCH];
                contract_row((int64_t)s2 - 1 - q, nr);
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1)
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) T1[v1][P][ch] = nr[v1][ch];
            }
            cur = s2;
        }
        T b2[P + 1];
#pragma unroll
        for (int k = 0; k <= P; ++k) b2[k] = brow[k];
        T acc[NCH][V1];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int v1 = 0; v1 < V1; ++v1) acc[ch][v1] = b2[0] * T1[v1][0][ch];
#pragma unroll
        for (int k = 1; k <= P; ++k)
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1) acc[ch][v1] = fma(b2[k], T1[v1][k][ch], acc[ch][v1]);
        if (NURBS) {
#pragma unroll
            for (int v1 = 0; v1 < V1; ++v1) {
                const T inv = T(1) / acc[NCH - 1][v1];
#pragma unroll
                for (int q = 0; q < NT; ++q) acc[q][v1] *= inv;
            }
        }
        if (full_tile) {
#pragma unroll
            for (int q = 0; q < NT; ++q) sg_store_vec<T, V1>(dst + a.n_total * q, acc[q], true, V1);
        } else {
#pragma unroll
            for (int q = 0; q < NT; ++q)
                if (q < no) sg_store_vec<T, V1>(dst + a.n_total * q, acc[q], false, nv1);
        }
    }
}
