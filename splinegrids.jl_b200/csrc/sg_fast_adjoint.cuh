// sg_fast_adjoint.cuh -- atomics-free, deterministic evaluate_adjoint! (K4) as a sequence of
// per-dimension transposed contractions ("passes"), from the slowest sample axis to the first:
//
//   pass A (dims D..2):  Y[q, i, r] = sum_{j in range(i)} B_d[j, i-base(j)] * X[q, j, r]
//       q = flattened faster dims (contiguous -> coalesced, vector loads), r = slower dims incl. Nout.
//       A thread owns V consecutive q and MARCHES along j keeping the P+1 live output rows in registers;
//       when the knot span advances the oldest row is complete and is written out.  No cross-thread
//       reduction, no atomics.  For parallelism the marching axis can be cut into chunks of G whole knot
//       spans; a chunk writes its G+P (partial) rows to scratch and a tiny combine kernel adds the <= few
//       chunks that share a row.
//   pass B (dim 1):      cp[i, r] = sum_{j in range(i)} B_1[j, i-base(j)] * X[j, r]   (gather form)
//
// Every pass shrinks the data by n_d / c_d, so the first pass (reading the full eval array once, coalesced)
// dominates.  range(i) comes from the span_start arrays built on device by the prep kernel; all kernels
// exit immediately if that kernel flagged non-monotone span indices (the atomic scatter kernel then runs).
// Rational (NURBS) 2-D grids: the first pass also forward-marches the weights to get each sample's
// denominator (eval is scaled by 1/denom on the fly) and pass B multiplies by w: R' e = w .* B'(e ./ (B w)).
// Reference semantics: src/adjoint.jl:1-83.
#pragma once
#include <type_traits>
#include "sg_adjoint_generic.cuh"
#include "sg_common.cuh"
#include "sg_fast_eval.cuh"

#define SG_ADJ_PIECE 64   // marching steps staged in shared memory at a time

// reciprocal of the rational denominator: correctly rounded for both types; the Float32 one is MUFU.RCP + one refinement
// instead of the full IEEE division sequence
__device__ __forceinline__ float sg_recip(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double sg_recip(double x) { return 1.0 / x; }

// All fast kernels return at their first instruction when the prep kernel flagged non-monotone span indices (the atomic
// scatter kernel then does the work) -- no host synchronisation.
enum { SG_PATH_MULTIPASS = 0 };
__device__ __forceinline__ bool sg_adj_path_active(const SgAdjointHeader *h, int path)
{
    (void)path;
    return h->nonmonotone == 0;
}
// Is control index i_D (1-based) of the last dimension inside the support [first_span - p, last_span]?
__device__ __forceinline__ bool sg_adj_row_in_support(const SgAdjointHeader *h, int last_dim, int last_P, int64_t last_div,
                                                      int64_t last_c, int64_t r)
{
    if (last_dim < 0) return true;
    const int64_t iD = (r / last_div) % last_c + 1;
    return iD >= h->span_first[last_dim] - last_P && iD <= h->span_last[last_dim];
}

template <typename T>
struct SgAdjPassArgs {
    const T *X;                 // input  [inner][n_d][outer]
    T *Y;                       // output [inner][rows][chunks][outer]  (rows = G+P; == [inner][c_d][outer] if chunks==1)
    const T *table;             // B_d, (n_d, P+1) column-major (derivative slice already selected)
    const int32_t *index;       // span per sample (1-based)
    const int32_t *span_start;  // [0..c_d+1]
    const SgAdjointHeader *hdr;
    int64_t inner, n_d, c_d;
    int G, nchunks;             // spans per chunk, number of chunks
    // rational 2-D first pass only
    const T *weights;           // (c_1, c_2)
    const T *table1;            // B_1 (n_1, P+1)
    const int32_t *index1;
    int64_t c1;
    int path;                   // SG_PATH_*
    int dim;                    // 0-based dimension contracted by this pass (selects hdr->span_first/last)
    int restrict_spans;         // 1 (last dimension only): visit just the spans that hold samples; rows outside stay unwritten
    // passes over dims below the last one: the outer index r contains the LAST dimension's control index,
    // i_D = (r / last_div) % last_c; rows outside the support of this (slab of the) grid are skipped.
    int last_dim;               // -1: no skipping
    int last_P;
    int64_t last_div, last_c;
    // fused 2-D march (F1): column-block tables of dimension 1 (prep kernel), weights [block][li][r]
    const SgM2gBlockHdr *bt_hdr;
    const int32_t *bt_lol;
    const T *bt_w;
    int icap, rmcap, nb1;
};

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
// NT = independent outer channels (e.g. the Nout planes) processed by one thread: they share the table
// rows, the span bookkeeping and -- for rational grids -- the denominators.
// V = samples of the contiguous axis per thread; V > 1 requires inner % V == 0 and 16-byte aligned
// arrays (checked by the dispatcher, which otherwise instantiates V = 1), so every active thread owns
// exactly V columns and all loads/stores are full vectors.
template <typename T, int V>
__device__ __forceinline__ void sg_load_vec(const T *__restrict__ p, T (&x)[V])
{
    typename SgVecT<T, V>::type pk = __ldcs(reinterpret_cast<const typename SgVecT<T, V>::type *>(p));
    const T *pq = reinterpret_cast<const T *>(&pk);
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = pq[v];
}

// F1 (2-D grids, planned calls): the kernel ALSO contracts dimension 1.  Whenever a control row of dimension 2 is finished
// (once per knot span, i.e. every n2 / spans rows -- rare), the CTA parks the row's 128 * V values per channel in shared
// memory and gathers the NI control indices its column block touches: one warp per output (li, channel), the lanes walk
// the support (coalesced weights from the per-block table, conflict-free shared loads), a shuffle tree finishes the sum.
// The partials then hold NI values per column block instead of 128 * V: the (n1, c2) intermediate, the chunk-combine
// kernel and the first-dimension kernel of the multi-pass pipeline disappear (sg_adj_combine_f2d_kernel sums the halos).
template <typename T, int P, int V, int NT, bool RAT2D, bool F1 = false>
__global__ void __launch_bounds__(128) sg_adj_march_kernel(const __grid_constant__ SgAdjPassArgs<T> a)
{
    if (!sg_adj_path_active(a.hdr, a.path)) return;
    __shared__ __align__(16) T bs[SG_ADJ_PIECE * (P + 1)];
    __shared__ int ss[SG_ADJ_PIECE];
    constexpr int BW = 128 * V;                                         // columns per CTA (F1: one column block)
    __shared__ __align__(16) T Es[F1 ? NT * BW : 1];
    __shared__ int lol_s[F1 ? 136 : 1];
    constexpr int E = 1;
    constexpr int WD = P + 1 + E;
    constexpr int U = NT >= 2 ? 4 : 8;                                 // steps whose loads are issued together

    int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    // F1: every thread takes part in the epilogue's barriers; a thread past the end re-reads the last columns (its
    // values are parked at positions no control index of the block reads)
    if (F1) q0 = min(q0, a.inner - V);
    const int c = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.z * NT;                       // first of this thread's NT outer channels
    if (!sg_adj_row_in_support(a.hdr, a.last_dim, a.last_P, a.last_div, a.last_c, r)) return;   // block-uniform
    const int s_lo0 = P + 1 + c * a.G;                                 // first span of this chunk (1-based)
    const int s_hi0 = (int)min((int64_t)s_lo0 + a.G, a.c_d + 1);       // one past the last span
    // only the spans that actually hold samples (a slab of a sharded grid covers a sub-range)
    const int s_lo = a.restrict_spans ? max(s_lo0, a.hdr->span_first[a.dim]) : s_lo0;
    const int s_hi = a.restrict_spans ? min(s_hi0, a.hdr->span_last[a.dim] + 1) : s_hi0;
    if (s_lo >= s_hi) return;                                          // block-uniform
    const int64_t j_lo = a.span_start[s_lo], j_hi = a.span_start[s_hi];
    bool active = F1 || q0 < a.inner;
    const int rows = a.G + P;
    const int64_t inner = a.inner;

    const T *__restrict__ xp = a.X + q0 + inner * (j_lo + a.n_d * r);
    T *__restrict__ yp = a.Y + q0 + inner * ((int64_t)rows * (c + (int64_t)a.nchunks * r) + (s_lo - s_lo0));   // oldest live row
    // F1: partials [li][row][chunk][block][channel]
    int f_ni = 0;
    int64_t f_row = s_lo - s_lo0;
    if (F1) {
        f_ni = a.bt_hdr[blockIdx.x].ni;
        for (int q = threadIdx.x; q < f_ni; q += blockDim.x) lol_s[q] = sg_ldg(a.bt_lol + (int64_t)blockIdx.x * a.icap + q);
    }
    const int64_t x_ch = inner * a.n_d;                                // channel strides
    const int64_t y_ch = inner * (int64_t)rows * a.nchunks;

    T acc[NT][V][P + 1];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int k = 0; k <= P; ++k) acc[t][v][k] = T(0);
    int cur = s_lo;

    // rational 2-D: forward march of the weights for this thread's V columns (cf. sg_eval2d_march_kernel)
    T W1[V][WD];
    T Tw[V][P + 1];
    int64_t col1[WD];
    bool reg1 = true;
    if (RAT2D && active) {
        int min1;
        reg1 = sg_expand_weights<T, P, V, E>(a.table1, a.index1, inner, q0, W1, min1);
#pragma unroll
        for (int q = 0; q < WD; ++q) col1[q] = min((int64_t)min1 + q, a.c1 - 1);
    }
    auto wrow = [&](int64_t i2, T (&out)[V]) {   // out[v] = sum_a B1[j1_v, a] w[i1+a, i2]
#pragma unroll
        for (int v = 0; v < V; ++v) out[v] = T(0);
        if (reg1) {
#pragma unroll
            for (int aq = 0; aq < WD; ++aq) {
                const T w = sg_ldg(a.weights + i2 * a.c1 + col1[aq]);
#pragma unroll
                for (int v = 0; v < V; ++v) out[v] = fma(W1[v][aq], w, out[v]);
            }
        } else {   // irregular thread: direct per-column sums
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int64_t j1 = q0 + v;
                const int64_t b1 = sg_ldg(a.index1 + j1) - P - 1;
                T sacc = T(0);
                for (int k = 0; k <= P; ++k) sacc = fma(sg_ldg(a.table1 + j1 + inner * k), sg_ldg(a.weights + i2 * a.c1 + b1 + k), sacc);
                out[v] = sacc;
            }
        }
    };
    if (RAT2D && active) {   // window for span `cur`: control rows cur-P-1 .. cur-1 (0-based)
#pragma unroll
        for (int k = 0; k <= P; ++k) {
            T o[V];
            wrow((int64_t)cur - P - 1 + k, o);
#pragma unroll
            for (int v = 0; v < V; ++v) Tw[v][k] = o[v];
        }
    }

    // write the oldest live row (complete, or chunk-partial) and slide the window by one span
    auto emit_oldest = [&]() {
        if (F1) __syncthreads();                                        // the previous row's gather is over
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            T o[V];
#pragma unroll
            for (int v = 0; v < V; ++v) o[v] = acc[t][v][0];
            if (F1) {
                typename SgVecT<T, V>::type pk;
                T *pq = reinterpret_cast<T *>(&pk);
#pragma unroll
                for (int v = 0; v < V; ++v) pq[v] = o[v];
                *reinterpret_cast<typename SgVecT<T, V>::type *>(Es + t * BW + threadIdx.x * V) = pk;
            } else {
                sg_store_vec<T, V>(yp + y_ch * t, o, true, V);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
#pragma unroll
                for (int k = 0; k < P; ++k) acc[t][v][k] = acc[t][v][k + 1];
                acc[t][v][P] = T(0);
            }
        }
        yp += inner;
        ++cur;
        if (F1) {
            __syncthreads();
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const T *__restrict__ wblk = a.bt_w + (int64_t)blockIdx.x * a.rmcap * a.icap;
            for (int out = warp; out < f_ni * NT; out += 4) {           // one warp per output (li, channel)
                const int t = out / f_ni, li = out - t * f_ni;
                const int ll = lol_s[li];
                const int lo = ll & 0xffff, len = ll >> 16;
                const T *__restrict__ wp = wblk + (int64_t)li * a.rmcap;
                const T *__restrict__ ep = Es + t * BW + lo;
                T sacc = T(0);
                for (int rr = lane; rr < len; rr += 32) sacc = fma(sg_ldg(wp + rr), ep[rr], sacc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
                if (lane == 0)
                    a.Y[li + (int64_t)a.icap * (f_row + (int64_t)rows * (c + (int64_t)a.nchunks * ((int64_t)blockIdx.x + (int64_t)a.nb1 * (r + t))))] = sacc;
            }
            ++f_row;
        }
        if (RAT2D && cur <= (int)a.c_d) {
            T o[V];
            wrow((int64_t)cur - 1, o);
#pragma unroll
            for (int v = 0; v < V; ++v) {
#pragma unroll
                for (int k = 0; k < P; ++k) Tw[v][k] = Tw[v][k + 1];
                Tw[v][P] = o[v];
            }
        }
    };

    // one marching step: sample s of the staged piece, values x[NT][V]
    // the accumulation of one sample row of span `cur` (no span bookkeeping)
    auto accumulate = [&](int s, T (&x)[NT][V]) {
        T b[P + 1];
#pragma unroll
        for (int k = 0; k <= P; ++k) b[k] = bs[s * (P + 1) + k];
        if (RAT2D) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                T den = b[0] * Tw[v][0];
#pragma unroll
                for (int k = 1; k <= P; ++k) den = fma(b[k], Tw[v][k], den);
                const T inv = sg_recip(den);
#pragma unroll
                for (int t = 0; t < NT; ++t) x[t][v] *= inv;
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int k = 0; k <= P; ++k)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[t][v][k] = fma(b[k], x[t][v], acc[t][v][k]);
    };
    auto step = [&](int s, T (&x)[NT][V]) {
        const int sp = ss[s];
        if (cur < sp) {
            do emit_oldest(); while (cur < sp);
        }
        accumulate(s, x);
    };

    for (int64_t jp = j_lo; jp < j_hi; jp += SG_ADJ_PIECE) {
        const int np = (int)min((int64_t)SG_ADJ_PIECE, j_hi - jp);
        __syncthreads();
        for (int s = threadIdx.x; s < np; s += blockDim.x) {
            ss[s] = sg_ldg(a.index + jp + s);
#pragma unroll
            for (int k = 0; k <= P; ++k) bs[s * (P + 1) + k] = sg_ldg(a.table + jp + s + a.n_d * k);
        }
        __syncthreads();
        if (!active) continue;
        int s = 0;
        // full groups of U steps: all loads first (memory-level parallelism), no bounds checks
        for (; s + U <= np; s += U) {
            T xs[U][NT][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int t = 0; t < NT; ++t) sg_load_vec<T, V>(xp + x_ch * t, xs[u][t]);
                xp += inner;
            }
            if (ss[s + U - 1] == cur) {   // block-uniform: the whole batch lies in the current span (indices are monotone)
#pragma unroll
                for (int u = 0; u < U; ++u) accumulate(s + u, xs[u]);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) step(s + u, xs[u]);
            }
        }
        for (; s < np; ++s) {   // tail
            T x1[NT][V];
#pragma unroll
            for (int t = 0; t < NT; ++t) sg_load_vec<T, V>(xp + x_ch * t, x1[t]);
            xp += inner;
            step(s, x1);
        }
    }
    if (!active) return;
    // flush: finish the chunk's spans, then the P still-live rows
    if (F1) {
        while (cur < s_hi + P) emit_oldest();
        return;
    }
    while (cur < s_hi) emit_oldest();
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int k = 0; k < P; ++k) {
            T o[V];
#pragma unroll
            for (int v = 0; v < V; ++v) o[v] = acc[t][v][k];
            sg_store_vec<T, V>(yp + y_ch * t + inner * k, o, true, V);
        }
}

// combine chunk partials: Y[q, i, r] = sum_c P[q, i - (c*G + 1), c, r]   (i 1-based control index).
// grid = (inner / (128*V), ceil(c_d / SG_COMBINE_ROWS), outer): no integer division, vector loads/stores.
#define SG_COMBINE_ROWS 8
template <typename T, int V>
__global__ void __launch_bounds__(128) sg_adj_combine_kernel(T *__restrict__ Y, const T *__restrict__ Pp, const SgAdjointHeader *hdr,
                                                             int64_t inner, int64_t c_d, int G, int nchunks, int P, bool vec_ok, int path,
                                                             int dim, int restrict_spans, int last_dim, int last_P, int64_t last_div, int64_t last_c)
{
    if (!sg_adj_path_active(hdr, path)) return;
    const int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (q0 >= inner) return;
    const int64_t r = blockIdx.z;
    if (!sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c, r)) return;
    const int sf = restrict_spans ? hdr->span_first[dim] : P + 1;
    const int sl = restrict_spans ? hdr->span_last[dim] : (int)c_d;
    const int nv = (int)min((int64_t)V, inner - q0);
    const bool vec = vec_ok && nv == V;
    const int rows = G + P;
    const int64_t i_end = min((int64_t)(blockIdx.y + 1) * SG_COMBINE_ROWS, c_d);
    for (int64_t i = (int64_t)blockIdx.y * SG_COMBINE_ROWS + 1; i <= i_end; ++i) {
        if (i < sf - P || i > sl) continue;                             // row not touched by any sample: stays unread
        // chunk c holds rows i in [c*G + 1, c*G + G + P]
        int64_t c_hi = (i - 1) / G;
        if (c_hi > nchunks - 1) c_hi = nchunks - 1;
        int64_t c_lo = (i - P - 1 >= 0) ? (i - P - 1) / G : 0;
        if (c_lo > 0 && (c_lo - 1) * G + G + P >= i) --c_lo;
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = T(0);
        for (int64_t c = c_lo; c <= c_hi; ++c) {
            const int64_t local = i - (c * G + 1);
            if (local < 0 || local >= rows) continue;
            // rows a chunk really wrote: spans [max(s_lo0, sf), min(s_hi0, sl+1)) -> control rows [lo-P, hi-1]
            const int64_t cs_lo = max((int64_t)(P + 1 + c * G), (int64_t)sf);
            const int64_t cs_hi = min(min((int64_t)(P + 1 + c * G + G), c_d + 1), (int64_t)sl + 1);
            if (cs_lo >= cs_hi || i < cs_lo - P || i > cs_hi - 1) continue;
            const T *__restrict__ src = Pp + q0 + inner * (local + (int64_t)rows * (c + (int64_t)nchunks * r));
            if (vec) {
                T x[V];
                sg_load_vec<T, V>(src, x);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += x[v];
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) if (v < nv) acc[v] += __ldcs(src + v);
            }
        }
        sg_store_vec<T, V>(Y + q0 + inner * ((i - 1) + c_d * r), acc, vec, nv);
    }
}

// Halo sum of the fused 2-D march: cp[i1, i2, o] = (w[i1, i2]) * sum over the chunks of dimension 2 that hold row i2 and the
// column blocks that hold i1 of the partials; writes every control point (zeros outside the support).
template <typename T, bool RATIONAL>
__global__ void __launch_bounds__(128) sg_adj_combine_f2d_kernel(T *__restrict__ cp, const T *__restrict__ Pp, const int32_t *__restrict__ g_lo,
                                                                 const SgM2gBlockHdr *__restrict__ bt_hdr, const SgAdjointHeader *hdr,
                                                                 const T *__restrict__ weights, int64_t c1, int64_t c2, int G, int nchunks, int P,
                                                                 int icap, int nb1, int bw)
{
    const int64_t i1 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 0-based
    if (i1 >= c1) return;
    const int64_t i = (int64_t)blockIdx.y + 1;                          // 1-based control row of dimension 2
    const int64_t o = blockIdx.z;
    const int2 gl = *reinterpret_cast<const int2 *>(g_lo + 2 * i1);
    const int sf = hdr->span_first[1], sl = hdr->span_last[1];
    const int rows = G + P;
    T acc = T(0);
    if (hdr->nonmonotone == 0 && gl.y > 0 && i >= sf - P && i <= sl) {
        int64_t c_hi = (i - 1) / G;
        if (c_hi > nchunks - 1) c_hi = nchunks - 1;
        int64_t c_lo = (i - P - 1 >= 0) ? (i - P - 1) / G : 0;
        if (c_lo > 0 && (c_lo - 1) * G + G + P >= i) --c_lo;
        const int jb_a = gl.x / bw, jb_b = min((gl.x + gl.y - 1) / bw, nb1 - 1);
        for (int64_t c = c_lo; c <= c_hi; ++c) {
            const int64_t local = i - (c * G + 1);
            if (local < 0 || local >= rows) continue;
            // rows a chunk really wrote: spans [max(s_lo0, sf), min(s_hi0, sl+1)) -> control rows [lo-P, hi-1]
            const int64_t cs_lo = max((int64_t)(P + 1 + c * G), (int64_t)sf);
            const int64_t cs_hi = min(min((int64_t)(P + 1 + c * G + G), c2 + 1), (int64_t)sl + 1);
            if (cs_lo >= cs_hi || i < cs_lo - P || i > cs_hi - 1) continue;
            for (int jb = jb_a; jb <= jb_b; ++jb) {
                const int li = (int)(i1 + 1) - bt_hdr[jb].i1_lo;
                if (li < 0 || li >= icap) continue;
                acc += __ldcs(Pp + li + (int64_t)icap * (local + (int64_t)rows * (c + (int64_t)nchunks * (jb + (int64_t)nb1 * o))));
            }
        }
    }
    const int64_t lin = i1 + c1 * ((i - 1) + c2 * o);
    cp[lin] = RATIONAL ? acc * sg_ldg(weights + i1 + c1 * (i - 1)) : acc;
}

// ---------------------------------------------------------------------------------------------
// pass B: first (contiguous) dimension, gather form.  A group of L lanes cooperates on one output
// (i, r): the lanes read consecutive samples j (coalesced) and the partial sums are combined with warp
// shuffles.  L = 8 for short ranges (few samples per knot span), 32 for long ones.
// ---------------------------------------------------------------------------------------------
template <typename T, int L, bool RATIONAL>
__global__ void __launch_bounds__(256) sg_adj_first_dim_kernel(T *__restrict__ cp, const T *__restrict__ X, const T *__restrict__ table,
                                                               const int32_t *__restrict__ index, const int32_t *__restrict__ span_start,
                                                               const SgAdjointHeader *hdr, int64_t n1, int64_t c1, int64_t outer, int P,
                                                               const T *__restrict__ weights, int64_t cp_total, int path,
                                                               int last_dim, int last_P, int64_t last_div, int64_t last_c)
{
    if (!sg_adj_path_active(hdr, path)) return;
    if (!sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c,
                               (int64_t)blockIdx.y + (int64_t)gridDim.y * blockIdx.z)) return;   // cp stays zero (memset)
    // grid = (ceil(c1 / (256/L)), outer split over y and z): no integer division
    const int64_t i0 = (int64_t)blockIdx.x * (256 / L) + threadIdx.x / L;   // 0-based control index
    const int64_t r = (int64_t)blockIdx.y + (int64_t)gridDim.y * blockIdx.z;
    const int lane = threadIdx.x % L;
    const bool valid = i0 < c1 && r < outer;
    const int64_t lin = i0 + c1 * r;
    T acc = T(0);
    if (valid) {
        const int64_t i = i0 + 1;   // 1-based control index
        const int64_t s0 = i > P + 1 ? i : P + 1;
        const int64_t s1 = i + P < c1 ? i + P : c1;
        const int64_t lo = span_start[s0], hi = span_start[s1 + 1];
        const T *__restrict__ xr = X + n1 * r;
        for (int64_t j = lo + lane; j < hi; j += L) {
            const int k = (int)(i - sg_ldg(index + j) + P);
            acc = fma(sg_ldg(table + j + n1 * k), sg_ldg(xr + j), acc);
        }
    }
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off, L);
    if (valid && lane == 0) {
        if (RATIONAL) acc *= sg_ldg(weights + lin % cp_total);
        cp[lin] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// pass B, short-range regime (few samples per knot span): one thread per control index i keeps its
// <= RMAX gather weights in registers and loops over many rows r (the weights do not depend on r).
// Lanes own consecutive i, so a warp reads one contiguous stretch of the row (L1-resident after the first
// touch).  A thread whose range is longer than RMAX falls back to table look-ups (still correct).
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX, bool RATIONAL>
__global__ void __launch_bounds__(128) sg_adj_first_dim_rows_kernel(T *__restrict__ cp, const T *__restrict__ X, const T *__restrict__ table,
                                                                    const int32_t *__restrict__ index, const int32_t *__restrict__ span_start,
                                                                    const SgAdjointHeader *hdr, int64_t n1, int64_t c1, int64_t outer, int P,
                                                                    int rows_per_block, const T *__restrict__ weights, int64_t cp_total, int path,
                                                                    int last_dim, int last_P, int64_t last_div, int64_t last_c)
{
    if (!sg_adj_path_active(hdr, path)) return;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 0-based control index
    if (i0 >= c1) return;
    const int64_t i = i0 + 1;
    const int64_t s0 = i > P + 1 ? i : P + 1;
    const int64_t s1 = i + P < c1 ? i + P : c1;
    const int64_t lo = span_start[s0], hi = span_start[s1 + 1];
    const int len = (int)(hi - lo);
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r_hi = min(r_lo + rows_per_block, outer);
    if (len > RMAX) {   // long range: direct look-ups
        for (int64_t r = r_lo; r < r_hi; ++r) {
            if (!sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c, r)) continue;
            T acc = T(0);
            for (int64_t j = lo; j < hi; ++j) {
                const int k = (int)(i - sg_ldg(index + j) + P);
                acc = fma(sg_ldg(table + j + n1 * k), sg_ldg(X + j + n1 * r), acc);
            }
            const int64_t lin = i0 + c1 * r;
            cp[lin] = RATIONAL ? acc * sg_ldg(weights + lin % cp_total) : acc;
        }
        return;
    }
    T w[RMAX];
#pragma unroll
    for (int t = 0; t < RMAX; ++t) {
        const int64_t j = lo + t;
        if (t < len) {
            const int k = (int)(i - sg_ldg(index + j) + P);
            w[t] = sg_ldg(table + j + n1 * k);
        } else {
            w[t] = T(0);
        }
    }
    const int64_t jmax = n1 - 1;
    int64_t r = r_lo;
    for (; r + 1 < r_hi; r += 2) {   // two rows per iteration for ILP
        const bool in0 = sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c, r);
        const bool in1 = sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c, r + 1);
        if (!in0 && !in1) continue;                                     // outside the slab's support: cp stays zero
        const T *__restrict__ x0 = X + n1 * (in0 ? r : r + 1), *__restrict__ x1 = X + n1 * (in1 ? r + 1 : r);
        T a0 = T(0), a1 = T(0);
#pragma unroll
        for (int t = 0; t < RMAX; ++t) {
            if (t < len) {   // len is warp-nearly-uniform; padded tail skipped
                const int64_t j = min(lo + t, jmax);
                a0 = fma(w[t], sg_ldg(x0 + j), a0);
                a1 = fma(w[t], sg_ldg(x1 + j), a1);
            }
        }
        const int64_t lin = i0 + c1 * r;
        if (RATIONAL) { a0 *= sg_ldg(weights + lin % cp_total); a1 *= sg_ldg(weights + (lin + c1) % cp_total); }
        if (in0) cp[lin] = a0;
        if (in1) cp[lin + c1] = a1;
    }
    if (r < r_hi && sg_adj_row_in_support(hdr, last_dim, last_P, last_div, last_c, r)) {
        const T *__restrict__ x0 = X + n1 * r;
        T a0 = T(0);
#pragma unroll
        for (int t = 0; t < RMAX; ++t)
            if (t < len) a0 = fma(w[t], sg_ldg(x0 + min(lo + t, jmax)), a0);
        const int64_t lin = i0 + c1 * r;
        cp[lin] = RATIONAL ? a0 * sg_ldg(weights + lin % cp_total) : a0;
    }
}

// =============================================================================================
// 3-D "double march": contract dimensions 3 AND 2 in one pass, entirely in registers.
//
// A thread owns ONE sample column j1 (lanes = consecutive j1: coalesced 8/4-byte loads, full sectors per warp),
// a tile of G2 whole knot spans of dimension 2 (S = G2+P control slots) and a chunk of G3 spans of dimension 3.
// For every sample plane j3 it contracts the tile's rows over dimension 2 into T[S] (the slot of a row is
// (its span - tile start) + k: COMPILE-TIME indices, because the loops run over spans g = 0..G2-1 and, inside
// a span, over its rows), then marches dimension 3 with the usual P+1 live planes per slot: acc3[S][P+1].
// When span 3 advances, S completed values leave per thread.  No shared-memory exchange, no atomics; the
// only block-level sync is the staging of the (CTA-uniform) table rows.
// Output: partial[j1][slot S][tile2][row3 G3+P][chunk3][o]; the tile/chunk halos are summed -- and dimension 1 is
// contracted -- by sg_adj_post2_kernel (sg_adjoint_post2.cuh).  For C3 the pass writes 135 MB instead of the
// 268 + 67 MB of two separate passes.  The chunks of dimension 3 adapt on device to the spans that hold samples
// (sg_m2_chunk_len); a.G3 is only the row stride of a chunk in the partials.
// Rows of a span are processed RS at a time (any number of samples per span works).
// =============================================================================================
// Chunks of dimension 3 adapt to the spans that really hold samples (a slab of a sharded grid touches few of them):
// the host fixes the NUMBER of chunks (a CTA-count target) and the worst-case chunk length a.G3 (row stride of the
// partials); on device the active spans [span_first, span_last] are cut evenly, G3e = max(P, ceil(nact / chunks3)) <=
// a.G3 spans per chunk, chunk c = spans [sf + c*G3e, sf + (c+1)*G3e).  G3e >= P: only neighbouring chunks overlap.
__device__ __forceinline__ int sg_m2_chunk_len(const SgAdjointHeader *h, int P, int chunks3)
{
    const int nact = h->span_last[2] - h->span_first[2] + 1;
    return max(max(P, 1), (nact + chunks3 - 1) / chunks3);
}

#define SG_M2_FAST_ROWS 5      // row slots per knot span in the TMA-fed kernel's straight-line contraction
template <typename T>
struct SgAdj2Args {
    const T *X;                 // eval (n1, n2, n3, nout)
    T *Y;                       // partials
    const T *table2, *table3;   // (n2, P+1), (n3, P+1) selected derivative slices
    const int32_t *index3;
    const int32_t *start2, *start3;   // span_start arrays
    SgAdjointHeader *hdr;
    int64_t n1, n2, n3, c2, c3;
    int tiles2, G3, chunks3;
    int path;
};

template <typename T, int P, int G2, int RS>
__device__ __forceinline__ void sg_adj_march2_body(const SgAdj2Args<T> &a, int only_tiles_above_rows, const unsigned bx, const unsigned by,
                                                   const unsigned bz)
{
    constexpr int S = G2 + P;
    constexpr int PIECE = 32;
    __shared__ __align__(16) T b3s[PIECE * (P + 1)];
    __shared__ int s3s[PIECE];
    __shared__ int row0[G2 + 1];
    constexpr int B2ROWS = G2 * RS * 2;                                 // table rows of dimension 2 staged per tile
    __shared__ __align__(16) T b2s[B2ROWS * (P + 1)];

    const int64_t j1 = (int64_t)bx * blockDim.x + threadIdx.x;
    const int tile2 = by;
    const int c3k = bz % a.chunks3;
    const int64_t o = bz / a.chunks3;
    const bool active = j1 < a.n1;

    // spans of dimension 2 in this tile: [s2_lo, s2_lo + G2) clipped; rows of span g: [row0[g], row0[g+1])
    const int s2_lo = P + 1 + tile2 * G2;
    __syncthreads();                                                   // (persistent caller) the previous tile is done with row0
    if (threadIdx.x <= G2) {
        const int sidx = (int)min((int64_t)s2_lo + threadIdx.x, a.c2 + 1);
        row0[threadIdx.x] = a.start2[sidx];
    }
    __syncthreads();
    // complement of the TMA-fed kernel: only the tiles it skipped (more rows than its ring holds)
    if (only_tiles_above_rows > 0) {                                   // block-uniform: exactly the tiles the TMA kernel skipped
        bool tma_did_it = row0[G2] - row0[0] <= only_tiles_above_rows;
#pragma unroll
        for (int g = 0; g < G2; ++g) tma_did_it = tma_did_it && row0[g + 1] - row0[g] <= SG_M2_FAST_ROWS;
        if (tma_did_it) return;
    }
    {   // table rows of the tile (rows are contiguous: [row0[0], row0[G2])); rows beyond B2ROWS use global look-ups
        const int r_first = row0[0], n_rows = min(row0[G2] - row0[0], B2ROWS);
        for (int q = threadIdx.x; q < n_rows * (P + 1); q += blockDim.x) {
            const int rr = q / (P + 1), k = q % (P + 1);
            b2s[q] = sg_ldg(a.table2 + (r_first + rr) + a.n2 * k);
        }
    }
    // spans of dimension 3 in this chunk, restricted to those that hold samples (slabs of a sharded grid)
    const int G3e = sg_m2_chunk_len(a.hdr, P, a.chunks3);
    const int s3_lo = a.hdr->span_first[2] + c3k * G3e;
    const int s3_hi = min(s3_lo + G3e, a.hdr->span_last[2] + 1);
    if (s3_lo >= s3_hi) return;                                        // block-uniform
    const int64_t j3_lo = a.start3[s3_lo], j3_hi = a.start3[s3_hi];
    const int rows3 = a.G3 + P;

    T acc3[S][P + 1];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k <= P; ++k) acc3[s][k] = T(0);
    int cur = s3_lo;

    const int64_t plane = a.n1 * a.n2;
    const T *__restrict__ xcol = a.X + j1 + plane * (a.n3 * o);        // + n1*row + plane*j3
    // Y index = j1 + n1*(slot + S*(tile2 + tiles2*(row3 + rows3*(chunk3 + chunks3*o))))
    const int64_t y_slot = a.n1;
    const int64_t y_row3 = a.n1 * (int64_t)S * a.tiles2;
    T *__restrict__ yp = a.Y + j1 + a.n1 * ((int64_t)S * tile2) + y_row3 * ((int64_t)rows3 * (c3k + (int64_t)a.chunks3 * o));

    auto emit_oldest = [&]() {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            __stcs(yp + y_slot * s, acc3[s][0]);
#pragma unroll
            for (int k = 0; k < P; ++k) acc3[s][k] = acc3[s][k + 1];
            acc3[s][P] = T(0);
        }
        yp += y_row3;
        ++cur;
    };

    for (int64_t jp = j3_lo; jp < j3_hi; jp += PIECE) {
        const int np = (int)min((int64_t)PIECE, j3_hi - jp);
        __syncthreads();
        for (int s = threadIdx.x; s < np; s += blockDim.x) {
            s3s[s] = sg_ldg(a.index3 + jp + s);
#pragma unroll
            for (int k = 0; k <= P; ++k) b3s[s * (P + 1) + k] = sg_ldg(a.table3 + jp + s + a.n3 * k);
        }
        __syncthreads();
        if (!active) continue;
        for (int s = 0; s < np; ++s) {
            const T *__restrict__ xpl = xcol + plane * (jp + s);
            // ---- contract the tile's rows over dimension 2: T2[g + k] += B2[row, k] * x[row]
            // all loads of the plane-tile are issued before the first FMA (memory-level parallelism)
            T x[G2][RS];
#pragma unroll
            for (int g = 0; g < G2; ++g)
#pragma unroll
                for (int q = 0; q < RS; ++q) x[g][q] = (row0[g] + q < row0[g + 1]) ? __ldcs(xpl + a.n1 * (int64_t)(row0[g] + q)) : T(0);
            T T2[S];
#pragma unroll
            for (int q = 0; q < S; ++q) T2[q] = T(0);
#pragma unroll
            for (int g = 0; g < G2; ++g) {
                const int r_lo = row0[g], r_hi = row0[g + 1];
#pragma unroll
                for (int q = 0; q < RS; ++q) {
                    if (r_lo + q < r_hi) {                              // CTA-uniform
#pragma unroll
                        for (int k = 0; k <= P; ++k) {
                            const int rr = r_lo + q - row0[0];
                            const T bw = rr < B2ROWS ? b2s[rr * (P + 1) + k] : sg_ldg(a.table2 + (r_lo + q) + a.n2 * k);
                            T2[g + k] = fma(bw, x[g][q], T2[g + k]);
                        }
                    }
                }
                for (int rb = r_lo + RS; rb < r_hi; rb += RS) {         // spans with more than RS rows: further batches
                    T xe[RS];
#pragma unroll
                    for (int q = 0; q < RS; ++q) xe[q] = (rb + q < r_hi) ? __ldcs(xpl + a.n1 * (int64_t)(rb + q)) : T(0);
#pragma unroll
                    for (int q = 0; q < RS; ++q) {
                        if (rb + q < r_hi) {
#pragma unroll
                            for (int k = 0; k <= P; ++k) T2[g + k] = fma(sg_ldg(a.table2 + (rb + q) + a.n2 * k), xe[q], T2[g + k]);
                        }
                    }
                }
            }
            // ---- march dimension 3
            const int sp = s3s[s];
            if (cur < sp) {
                do emit_oldest(); while (cur < sp);
            }
            T b[P + 1];
#pragma unroll
            for (int k = 0; k <= P; ++k) b[k] = b3s[s * (P + 1) + k];
#pragma unroll
            for (int q = 0; q < S; ++q)
#pragma unroll
                for (int k = 0; k <= P; ++k) acc3[q][k] = fma(b[k], T2[q], acc3[q][k]);
        }
    }
    if (!active) return;
    while (cur < s3_hi) emit_oldest();
#pragma unroll
    for (int k = 0; k < P; ++k) {
#pragma unroll
        for (int s = 0; s < S; ++s) __stcs(yp + y_slot * s + y_row3 * k, acc3[s][k]);
    }
}

template <typename T, int P, int G2, int RS>
__global__ void __launch_bounds__(128) sg_adj_march2_kernel(const __grid_constant__ SgAdj2Args<T> a, int only_tiles_above_rows)
{
    if (!sg_adj_path_active(a.hdr, a.path)) return;
    sg_adj_march2_body<T, P, G2, RS>(a, only_tiles_above_rows, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Complement of the TMA-fed kernel: a small persistent grid that walks the (column block, tile, chunk) space and does
// the tiles the TMA kernel skipped (more rows than its ring holds).  Normally there are none: the TMA kernel raises
// hdr->m2_skipped only when it skips a tile, so this kernel costs one launch and exits.
template <typename T, int P, int G2, int RS>
__global__ void __launch_bounds__(128) sg_adj_march2_complement_kernel(const __grid_constant__ SgAdj2Args<T> a, int only_tiles_above_rows,
                                                                       unsigned nbx, unsigned nby, unsigned nbz)
{
    if (!sg_adj_path_active(a.hdr, a.path) || a.hdr->m2_skipped == 0) return;
    const uint64_t total = (uint64_t)nbx * nby * nbz;
    for (uint64_t vb = blockIdx.x; vb < total; vb += gridDim.x)
        sg_adj_march2_body<T, P, G2, RS>(a, only_tiles_above_rows, (unsigned)(vb % nbx), (unsigned)((vb / nbx) % nby), (unsigned)(vb / ((uint64_t)nbx * nby)));
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variant of the double march: the tile's rows of each sample plane (n_rows x 128 columns) are streamed
// into an NS-stage shared-memory ring with 1-D bulk async copies (cp.async.bulk, one per row, issued by one
// elected thread, completion counted on an mbarrier per stage), NS-1 planes ahead of the consumers.  Loads no
// longer occupy registers or stall the math: the kernel becomes bandwidth-bound.  Tiles with more than RTMAX rows
// (or a ragged / misaligned n1) are handled by sg_adj_march2_kernel instead (host + device checks).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sg_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sg_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sg_smem_u32(bar))
                 : "memory");
}

// 3-D tiled bulk tensor load (TMA) issued by one elected lane of a converged warp; completion on the mbarrier
__device__ __forceinline__ void sg_m2_tma_load_3d_elect(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        "}\n" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// L2 policies: the sample array is read exactly once (evict first), the partials are read back by the post kernel right
// after this kernel (evict last): with a 126 MB L2 a good part of the 135 MB of partials never waits for HBM.
__device__ __forceinline__ uint64_t sg_l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t sg_l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void sg_m2_tma_load_3d_elect_hint(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar, uint64_t pol)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n"
        "}\n" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol)
        : "memory");
}
template <typename T>
__device__ __forceinline__ void sg_st_hint(T *p, T v, uint64_t pol);
template <>
__device__ __forceinline__ void sg_st_hint<double>(double *p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
template <>
__device__ __forceinline__ void sg_st_hint<float>(float *p, float v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
// Tensor maps over eval viewed as (n1, n2, n3*nout) with boxes of 128 columns x (1 .. SG_M2_FAST_ROWS) rows x 1 plane: the rows of
// one knot span of a tile are ONE tensor copy (4 copy instructions per plane instead of ~17 row copies).
struct SgM2Maps {
    CUtensorMap m[SG_M2_FAST_ROWS];
};

// Producer-warp primitives: executed by the WHOLE converged warp with warp-uniform operands (which then live in uniform
// registers); one elected lane issues.  A lane-divergent caller (if (lane == 0) ...) costs ~25 instructions per copy.
__device__ __forceinline__ void sg_m2_bulk_g2s_elect(uint32_t dst, const void *src, unsigned bytes, uint32_t bar)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
        "}\n" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void sg_m2_expect_tx_elect(uint32_t bar, unsigned bytes)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "}\n" ::"r"(bar),
        "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void sg_m2_mbar_wait_u(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_M2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_M2_DONE_%=;\n"
        "bra SG_M2_WAIT_%=;\n"
        "SG_M2_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void sg_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sg_smem_u32(bar)) : "memory");
}

// Weights of dimension 2 and the first sample of each of its knot spans, passed BY VALUE as a kernel parameter (a plan
// keeps the host copy): they are CTA-uniform, so the compiler reads them through the constant bank into uniform
// registers (LDCU) and feeds them to DFMA/FFMA as uniform operands -- none of the 80 weight loads per sample plane
// touches the shared-memory pipe any more (they were half of the kernel's LSU wavefronts), and absent row slots are
// skipped with uniform predicates.
#define SG_M2U_B2_BYTES 24576
#define SG_M2U_STARTS 512
template <typename T>
struct SgM2Uni {
    T b2[SG_M2U_B2_BYTES / sizeof(T)];   // [row of dimension 2][k], k = 0..P
    int start2[SG_M2U_STARTS];           // span_start of dimension 2, [0 .. c2 + 1]
    int start3[SG_M2U_STARTS];           // span_start of dimension 3, [0 .. c3 + 1]
    int span_first3, span_last3;         // first / last span of dimension 3 that holds samples (header values)
};
template <typename T>
struct SgM2UniNone {};

// Warp-specialised: warps 0-3 (128 threads) consume, warp 4 produces (one elected lane issues the bulk copies).
// full[st]  : producer -> consumers, completes when the plane's bytes have landed (expect_tx)
// empty[st] : consumers -> producer, 128 arrivals once every consumer has read the stage
// No block-wide barrier inside the plane loop; the (CTA-uniform) dimension-3 table rows are staged per piece of
// MAXPL planes.
template <typename T, int P, int G2, int RTMAX, int NS>
__global__ void __launch_bounds__(160, 3) sg_adj_march2_tma_kernel(const __grid_constant__ SgAdj2Args<T> a, const __grid_constant__ SgM2Maps maps,
                                                                   int use_maps)
{
    if (!sg_adj_path_active(a.hdr, a.path)) return;
    constexpr int S = G2 + P;
    constexpr int MAXPL = 128;                                          // planes per staged piece of dim-3 tables
    constexpr int CW = 128;                                             // columns per CTA == consumer threads
    // The ring is aligned to 128 bytes by hand (the launcher adds the slack).  Claiming __align__(128) on the extern array
    // would let the compiler fold the adjustment away, while the run-time base is only 16-byte aligned.
    extern __shared__ __align__(16) unsigned char sg_smem2[];
    T *xs = reinterpret_cast<T *>(sg_smem2 + ((128u - (sg_smem_u32(sg_smem2) & 127u)) & 127u));   // [NS][RTMAX][CW]
    __shared__ __align__(16) T b3s[MAXPL * (P + 1)];
    __shared__ int s3s[MAXPL];
    __shared__ int row0[G2 + 1];
    constexpr int RS5 = SG_M2_FAST_ROWS;                                // row slots per span of the straight-line contraction
    __shared__ __align__(16) T b2pad[G2 * RS5 * (P + 1)];               // [span g][row q][k], zero for absent rows
    __shared__ __align__(8) uint64_t full[NS];
    __shared__ __align__(8) uint64_t empty[NS];

    const int tid = threadIdx.x;
    // the warp index through a shuffle: the role branch is then warp-uniform for the compiler (uniform datapath inside)
    const bool is_producer = __shfl_sync(0xffffffffu, tid >> 5, 0) >= CW / 32;
    const int64_t j1_0 = (int64_t)blockIdx.x * CW;
    const int64_t j1 = j1_0 + tid;
    const int tile2 = blockIdx.y;
    const int c3k = blockIdx.z % a.chunks3;
    const int64_t o = blockIdx.z / a.chunks3;
    const bool active = !is_producer && j1 < a.n1;
    const int ncols = (int)min((int64_t)CW, a.n1 - j1_0);

    const int s2_lo = P + 1 + tile2 * G2;
    // rows of the tile's spans: [ur0[g], ur0[g + 1]) (register copies of row0[])
    int ur0[G2 + 1];
    if (tid <= G2) row0[tid] = a.start2[(int)min((int64_t)s2_lo + tid, a.c2 + 1)];
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) { sg_mbar_init(&full[q], 1); sg_mbar_init(&empty[q], CW / 32); }   // one arrival per consumer warp
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g <= G2; ++g) ur0[g] = row0[g];
    const int r_first = ur0[0], n_rows = ur0[G2] - ur0[0];
    // Straight-line contraction of dimension 2 (no data-dependent loops, no predicates): a ring stage holds G2 x RS5
    // fixed row slots, slot (g, q) = q-th row of span g of the tile.  The producer copies every present row into its
    // slot; absent slots are zeroed once here and never written again, and their weights are zero.  Tiles with a span
    // of more than RS5 rows are left to the register kernel (complement pass).  All CTA-uniform.
    static_assert(G2 * SG_M2_FAST_ROWS <= RTMAX, "ring stage holds G2 x RS5 row slots");
    bool fastrows = true;
#pragma unroll
    for (int g = 0; g < G2; ++g) fastrows = fastrows && ur0[g + 1] - ur0[g] <= RS5;
    if (!fastrows) {
        if (tid == 0) a.hdr->m2_skipped = 1;
        return;
    }
    if (!is_producer) {
        for (int q = tid; q < G2 * RS5 * (P + 1); q += CW) {
            const int k = q % (P + 1), gq = q / (P + 1), g = gq / RS5, qq = gq % RS5;
            b2pad[q] = qq < row0[g + 1] - row0[g] ? sg_ldg(a.table2 + (row0[g] + qq) + a.n2 * k) : T(0);
        }
        for (int sl = 0; sl < NS * G2 * RS5; ++sl) {                    // absent slots of every stage (never touched by the copies)
            const int gq = sl % (G2 * RS5), g = gq / RS5, qq = gq % RS5;
            if (qq >= row0[g + 1] - row0[g]) xs[(size_t)(sl / (G2 * RS5)) * RTMAX * CW + (size_t)gq * CW + tid] = T(0);
        }
    }

    const int sf3 = a.hdr->span_first[2], sl3 = a.hdr->span_last[2];
    const int G3e = max(max(P, 1), (sl3 - sf3 + 1 + a.chunks3 - 1) / a.chunks3);   // == sg_m2_chunk_len
    const int s3_lo = sf3 + c3k * G3e;
    const int s3_hi = min(s3_lo + G3e, sl3 + 1);
    if (s3_lo >= s3_hi) return;                                        // block-uniform
    const int64_t j3_lo = a.start3[s3_lo], j3_hi = a.start3[s3_hi];
    const int np_total = (int)(j3_hi - j3_lo);
    const int rows3 = a.G3 + P;

    T acc3[S][P + 1];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k <= P; ++k) acc3[s][k] = T(0);
    int cur = s3_lo;

    const int64_t plane = a.n1 * a.n2;
    const T *__restrict__ xtile = a.X + j1_0 + a.n1 * (int64_t)r_first + plane * (a.n3 * o + j3_lo);
    // 32-bit element offsets into the partials (the host checks that they fit): fewer live registers in the plane loop
    const unsigned y_slot = (unsigned)a.n1;
    const unsigned y_row3 = (unsigned)(a.n1 * (int64_t)S * a.tiles2);
    unsigned yoff = (unsigned)(j1 + a.n1 * ((int64_t)S * tile2) + (int64_t)y_row3 * ((int64_t)rows3 * (c3k + (int64_t)a.chunks3 * o)));
    T *__restrict__ const ybase = a.Y;
    const unsigned row_bytes = (unsigned)(ncols * sizeof(T));

    const uint64_t pol_last = sg_l2_policy_evict_last();
    auto emit_oldest = [&]() {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (active) sg_st_hint<T>(ybase + (yoff + y_slot * s), acc3[s][0], pol_last);
#pragma unroll
            for (int k = 0; k < P; ++k) acc3[s][k] = acc3[s][k + 1];
            acc3[s][P] = T(0);
        }
        yoff += y_row3;
        ++cur;
    };

    int st = 0;                                                         // ring stage / phase of the next plane (producer and
    unsigned ph = 0;                                                    // consumers each keep their own running copy)
    if (is_producer) {
        // The producer starts feeding the ring at once and runs through the whole chunk on its own; it never meets the
        // consumers at a block-wide barrier (they stage their tables behind a named barrier of their own).
        if (n_rows > 0) {                                               // whole warp, converged; one elected lane issues
            const uint32_t xs_u = sg_smem_u32(xs), full_u = sg_smem_u32(full), empty_u = sg_smem_u32(empty);
            int ra[G2], rn[G2];
#pragma unroll
            for (int g = 0; g < G2; ++g) { ra[g] = ur0[g] - r_first; rn[g] = ur0[g + 1] - ur0[g]; }
            if (use_maps) {
                // one tensor copy per knot span: box = 128 columns x rn[g] rows (columns past n1 are zero-filled by the TMA unit)
                const unsigned stage_tx = (unsigned)(CW * sizeof(T)) * (unsigned)n_rows;
                const int pl0 = (int)(a.n3 * o + j3_lo);                // first plane of the chunk in the (n1, n2, n3*nout) view
                const uint64_t pol_first = sg_l2_policy_evict_first();
                for (int p = 0; p < np_total; ++p) {
                    if (p >= NS) sg_m2_mbar_wait_u(empty_u + (uint32_t)st * 8u, ph ^ 1u);
                    const uint32_t fb = full_u + (uint32_t)st * 8u;
                    sg_m2_expect_tx_elect(fb, stage_tx);
                    const uint32_t dst = xs_u + (uint32_t)st * (uint32_t)(RTMAX * CW * sizeof(T));
#pragma unroll
                    for (int g = 0; g < G2; ++g)
                        if (rn[g] > 0)                                  // warp-uniform
                            sg_m2_tma_load_3d_elect_hint(dst + (uint32_t)(g * RS5 * CW * sizeof(T)), &maps.m[rn[g] - 1], (int)j1_0, r_first + ra[g], pl0 + p, fb, pol_first);
                    if (++st == NS) { st = 0; ph ^= 1u; }
                }
                return;
            }
            const unsigned stage_tx = row_bytes * (unsigned)n_rows;
            for (int p = 0; p < np_total; ++p) {
                if (p >= NS) sg_m2_mbar_wait_u(empty_u + (uint32_t)st * 8u, ph ^ 1u);   // every consumer warp has read this stage's previous plane
                const uint32_t fb = full_u + (uint32_t)st * 8u;
                sg_m2_expect_tx_elect(fb, stage_tx);
                const T *src = xtile + plane * (int64_t)p;
                const uint32_t dst = xs_u + (uint32_t)st * (uint32_t)(RTMAX * CW * sizeof(T));
#pragma unroll
                for (int g = 0; g < G2; ++g) {
#pragma unroll
                    for (int q = 0; q < RS5; ++q)
                        if (q < rn[g])                                  // warp-uniform
                            sg_m2_bulk_g2s_elect(dst + (uint32_t)((g * RS5 + q) * CW * sizeof(T)), src + a.n1 * (int64_t)(ra[g] + q), row_bytes, fb);
                }
                if (++st == NS) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }
    for (int p0 = 0; p0 < np_total; p0 += MAXPL) {
        const int p1 = min(p0 + MAXPL, np_total);
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");           // consumers are done with the previous piece's tables
        for (int s = tid; s < p1 - p0; s += CW) {
            s3s[s] = sg_ldg(a.index3 + j3_lo + p0 + s);
#pragma unroll
            for (int k = 0; k <= P; ++k) b3s[s * (P + 1) + k] = sg_ldg(a.table3 + j3_lo + p0 + s + a.n3 * k);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");           // (also covers the dimension-2 weights staged above)
        for (int p = p0; p < p1; ++p) {
            T T2[S];
#pragma unroll
            for (int q = 0; q < S; ++q) T2[q] = T(0);
            if (n_rows > 0) {
                sg_mbar_wait(&full[st], ph);
                const T *__restrict__ xst = xs + (size_t)st * RTMAX * CW + tid;
                // ---- contract the tile's rows over dimension 2 (values come from the shared-memory ring)
#pragma unroll
                for (int g = 0; g < G2; ++g) {
#pragma unroll
                    for (int q = 0; q < RS5; ++q) {
                        const T x = xst[(g * RS5 + q) * CW];
#pragma unroll
                        for (int k = 0; k <= P; ++k) T2[g + k] = fma(b2pad[(g * RS5 + q) * (P + 1) + k], x, T2[g + k]);
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) sg_mbar_arrive(&empty[st]);        // this warp no longer needs the stage
                if (++st == NS) { st = 0; ph ^= 1u; }
            }
            // ---- march dimension 3 (columns past n1 carry zeros / garbage that is never stored)
            const int sl = p - p0;
            const int sp = s3s[sl];
            if (cur < sp) {
                do emit_oldest(); while (cur < sp);
            }
            T b[P + 1];
#pragma unroll
            for (int k = 0; k <= P; ++k) b[k] = b3s[sl * (P + 1) + k];
#pragma unroll
            for (int q = 0; q < S; ++q)
#pragma unroll
                for (int k = 0; k <= P; ++k) acc3[q][k] = fma(b[k], T2[q], acc3[q][k]);
        }
    }
    while (cur < s3_hi) emit_oldest();
    if (!active) return;
#pragma unroll
    for (int k = 0; k < P; ++k) {
#pragma unroll
        for (int s = 0; s < S; ++s) sg_st_hint<T>(ybase + (yoff + y_slot * s + y_row3 * k), acc3[s][k], pol_last);
    }
}
