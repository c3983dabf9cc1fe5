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
#include "sg_adjoint_generic.cuh"
#include "sg_common.cuh"
#include "sg_fast_eval.cuh"

#define SG_ADJ_PIECE 64   // marching steps staged in shared memory at a time

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
};

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
// NT = independent outer channels (e.g. the Nout planes) processed by one thread: they share the table
// rows, the span bookkeeping and -- for rational grids -- the denominators.
template <typename T, int P, int V, int NT, bool RAT2D>
__global__ void __launch_bounds__(128) sg_adj_march_kernel(const __grid_constant__ SgAdjPassArgs<T> a, bool vec_ok)
{
    if (a.hdr->nonmonotone) return;
    __shared__ __align__(16) T bs[SG_ADJ_PIECE * (P + 1)];
    __shared__ int ss[SG_ADJ_PIECE];
    constexpr int E = 1;
    constexpr int WD = P + 1 + E;

    const int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int c = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.z * NT;                       // first of this thread's NT outer channels
    const int s_lo = P + 1 + c * a.G;                                  // first span of this chunk (1-based)
    const int s_hi = (int)min((int64_t)s_lo + a.G, a.c_d + 1);         // one past the last span
    const int64_t j_lo = a.span_start[s_lo], j_hi = a.span_start[s_hi];
    const bool active = q0 < a.inner;
    const int nv = active ? (int)min((int64_t)V, a.inner - q0) : 0;
    const bool vec = vec_ok && nv == V;
    const int rows = a.G + P;

    const T *__restrict__ xp = a.X + q0 + a.inner * (j_lo + a.n_d * r);
    T *__restrict__ yp = a.Y + q0 + a.inner * ((int64_t)rows * (c + (int64_t)a.nchunks * r));
    const int64_t x_ch = a.inner * a.n_d;                              // channel strides
    const int64_t y_ch = a.inner * (int64_t)rows * a.nchunks;

    T acc[NT][V][P + 1];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int k = 0; k <= P; ++k) acc[t][v][k] = T(0);
    int cur = s_lo;
    int row = 0;   // local index of the oldest live row (1-based control index cur-P  <->  row)

    // rational 2-D: forward march of the weights for this thread's V columns (cf. sg_eval2d_march_kernel)
    T W1[V][WD];
    T Tw[V][P + 1];
    int64_t col1[WD];
    bool reg1 = true;
    if (RAT2D && active) {
        int min1;
        reg1 = sg_expand_weights<T, P, V, E>(a.table1, a.index1, a.inner, q0, W1, min1);
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
            for (int v = 0; v < nv; ++v) {
                const int64_t j1 = q0 + v;
                const int64_t b1 = sg_ldg(a.index1 + j1) - P - 1;
                T sacc = T(0);
                for (int k = 0; k <= P; ++k) sacc = fma(sg_ldg(a.table1 + j1 + a.inner * k), sg_ldg(a.weights + i2 * a.c1 + b1 + k), sacc);
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

    auto emit_oldest = [&]() {   // write row `row` (complete or chunk-partial) and slide the window
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (active) {
                T o[V];
#pragma unroll
                for (int v = 0; v < V; ++v) o[v] = acc[t][v][0];
                sg_store_vec<T, V>(yp + y_ch * t + a.inner * row, o, vec, nv);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
#pragma unroll
                for (int k = 0; k < P; ++k) acc[t][v][k] = acc[t][v][k + 1];
                acc[t][v][P] = T(0);
            }
        }
        ++row;
        ++cur;
        if (RAT2D && active && cur <= (int)a.c_d) {
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
        // software pipeline: issue the loads of U steps back to back (memory-level parallelism), then consume
        constexpr int U = NT > 2 ? 4 : (NT == 2 ? 4 : 8);
        for (int s0 = 0; s0 < np; s0 += U) {
            T xs[U][NT][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (s0 + u < np) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const T *__restrict__ xq = xp + a.inner * u + x_ch * t;
                        if (vec) {
                            typename SgVecT<T, V>::type pk = __ldcs(reinterpret_cast<const typename SgVecT<T, V>::type *>(xq));
                            const T *pq = reinterpret_cast<const T *>(&pk);
#pragma unroll
                            for (int v = 0; v < V; ++v) xs[u][t][v] = pq[v];
                        } else {
#pragma unroll
                            for (int v = 0; v < V; ++v) xs[u][t][v] = v < nv ? __ldcs(xq + v) : T(0);
                        }
                    }
                }
            }
            xp += a.inner * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s = s0 + u;
                if (s < np) {
                    const int sp = ss[s];
                    while (cur < sp) emit_oldest();
                    T b[P + 1];
#pragma unroll
                    for (int k = 0; k <= P; ++k) b[k] = bs[s * (P + 1) + k];
                    if (RAT2D) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            T den = b[0] * Tw[v][0];
#pragma unroll
                            for (int k = 1; k <= P; ++k) den = fma(b[k], Tw[v][k], den);
                            const T inv = T(1) / den;
#pragma unroll
                            for (int t = 0; t < NT; ++t) xs[u][t][v] *= inv;
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int k = 0; k <= P; ++k)
#pragma unroll
                            for (int v = 0; v < V; ++v) acc[t][v][k] = fma(b[k], xs[u][t][v], acc[t][v][k]);
                }
            }
        }
    }
    // flush: finish the chunk's spans, then the P still-live rows
    while (cur < s_hi) emit_oldest();
    if (active) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int k = 0; k < P; ++k) {
                T o[V];
#pragma unroll
                for (int v = 0; v < V; ++v) o[v] = acc[t][v][k];
                sg_store_vec<T, V>(yp + y_ch * t + a.inner * (row + k), o, vec, nv);
            }
    }
}

// combine chunk partials: Y[q, i, r] = sum_c P[q, i - (c*G + 1), c, r]   (i 1-based control index).
// grid = (inner / (128*V), c_d, outer): no integer division in the kernel, vector loads/stores.
template <typename T, int V>
__global__ void __launch_bounds__(128) sg_adj_combine_kernel(T *__restrict__ Y, const T *__restrict__ Pp, const SgAdjointHeader *hdr,
                                                             int64_t inner, int64_t c_d, int G, int nchunks, int P, bool vec_ok)
{
    if (hdr->nonmonotone) return;
    const int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (q0 >= inner) return;
    const int64_t i = (int64_t)blockIdx.y + 1;
    const int64_t r = blockIdx.z;
    const int nv = (int)min((int64_t)V, inner - q0);
    const bool vec = vec_ok && nv == V;
    const int rows = G + P;
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
        const T *__restrict__ src = Pp + q0 + inner * (local + (int64_t)rows * (c + (int64_t)nchunks * r));
        if (vec) {
            typename SgVecT<T, V>::type pk = __ldcs(reinterpret_cast<const typename SgVecT<T, V>::type *>(src));
            const T *pq = reinterpret_cast<const T *>(&pk);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] += pq[v];
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) if (v < nv) acc[v] += __ldcs(src + v);
        }
    }
    sg_store_vec<T, V>(Y + q0 + inner * ((i - 1) + c_d * r), acc, vec, nv);
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
                                                               const T *__restrict__ weights, int64_t cp_total)
{
    if (hdr->nonmonotone) return;
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
