// sg_eval_multi.cuh -- evaluate! for SEVERAL derivative orders in one launch (2-D grids).
// Every real caller of the reference evaluates the value and two or three partial derivatives back to back on the same
// control points (docs/src/examples_optics.md:189-191: u, d1 u, d2 u; docs/src/examples_pde.md:69-72); each call
// re-reads the tables and re-contracts the same control-point window.  Here one thread block marches dimension 2 once:
// the per-dimension-1 contraction of each control row is done once per derivative tuple (their basis weights differ),
// the control points are loaded ONCE per row and shared by all tuples, the span bookkeeping, the column weights' set-up
// and the launch are shared, and each tuple's values leave through its own output array.
// Semantics per tuple: src/spline_grid.jl:130-182 with derivative_order = that tuple.
#pragma once
#include "sg_fast_eval.cuh"

#define SG_MULTI_MAX 4

template <typename T>
struct SgMultiArgs {
    T *eval[SG_MULTI_MAX];            // output array of every derivative tuple, each (n1, n2, nout)
    const T *table1[SG_MULTI_MAX];    // dimension-1 table slice selected by the tuple (n1, P+1)
    const T *table2[SG_MULTI_MAX];    // dimension-2 table slice selected by the tuple (n2, P+1)
};

template <typename T, int P, int V1, int ND>
__global__ void __launch_bounds__(128, 4) sg_eval2d_multi_kernel(const __grid_constant__ SgMultiArgs<T> m, const __grid_constant__ SgGridArgs<T> a,
                                                              const T *__restrict__ cp, int chunk, bool vec_ok)
{
    constexpr int E = 1;
    constexpr int WD = P + 1 + E;
    extern __shared__ __align__(16) unsigned char sg_smem_multi[];
    T *b2s = reinterpret_cast<T *>(sg_smem_multi);                              // [ND][chunk][P+1]
    int *s2s = reinterpret_cast<int *>(b2s + (size_t)ND * chunk * (P + 1));     // [chunk]

    const int64_t n1 = a.n_samples[0], n2 = a.n_samples[1];
    const int64_t c1 = a.n_cp[0];
    const int64_t j1_0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V1;
    const int64_t j2_lo = (int64_t)blockIdx.y * chunk;
    const int nstep = (int)min((int64_t)chunk, n2 - j2_lo);
    const int o = blockIdx.z;

    for (int s = threadIdx.x; s < nstep; s += blockDim.x) {
        s2s[s] = sg_ldg(a.index[1] + j2_lo + s);
#pragma unroll
        for (int q = 0; q < ND; ++q)
#pragma unroll
            for (int k = 0; k <= P; ++k) b2s[((size_t)q * chunk + s) * (P + 1) + k] = sg_ldg(m.table2[q] + j2_lo + s + n2 * k);
    }
    __syncthreads();
    if (j1_0 >= n1) return;

    // Column weights (same spans for every tuple, different derivative slices).  They are needed only when the span of
    // dimension 2 advances (once per n2 / spans rows), so they are NOT kept in registers (ND x V1 x WD values would cost a
    // third of the occupancy): contract_row re-expands them from the L1-resident tables.
    int min1 = 0;
    bool reg1;
    {
        T W0[V1][WD];
        reg1 = sg_expand_weights<T, P, V1, E>(m.table1[0], a.index[0], n1, j1_0, W0, min1);
    }
    const int nv1 = (int)min((int64_t)V1, n1 - j1_0);
    if (!reg1) {   // columns that straddle more than one extra span: direct evaluation of every sample
        for (int s = 0; s < nstep; ++s)
            for (int v1 = 0; v1 < nv1; ++v1) {
                const int64_t j1 = j1_0 + v1, j2 = j2_lo + s;
                const int64_t b1 = sg_ldg(a.index[0] + j1) - P - 1, b2 = sg_ldg(a.index[1] + j2) - P - 1;
#pragma unroll
                for (int q = 0; q < ND; ++q) {
                    T acc = T(0);
                    for (int k2 = 0; k2 <= P; ++k2)
                        for (int k1 = 0; k1 <= P; ++k1)
                            acc += sg_ldg(m.table1[q] + j1 + n1 * k1) * sg_ldg(m.table2[q] + j2 + n2 * k2) *
                                   sg_ldg(cp + a.cp_total * o + (b2 + k2) * c1 + b1 + k1);
                    m.eval[q][j1 + n1 * j2 + a.n_total * o] = acc;
                }
            }
        return;
    }
    int64_t col1[WD];
#pragma unroll
    for (int q = 0; q < WD; ++q) col1[q] = min(max((int64_t)min1 + q, (int64_t)0), c1 - 1);

    T T1[ND][V1][P + 1] = {};
    int cur = -0x40000000;
    const T *__restrict__ cpo = cp + a.cp_total * o;

    auto contract_row = [&](int64_t i2) {   // slide every tuple's window and append control row i2
        T c[WD];
#pragma unroll
        for (int aq = 0; aq < WD; ++aq) c[aq] = sg_ldg(cpo + i2 * c1 + col1[aq]);   // loaded once, used by all tuples
#pragma unroll
        for (int q = 0; q < ND; ++q) {
            T W1[V1][WD];
            int mb;
            sg_expand_weights<T, P, V1, E>(m.table1[q], a.index[0], n1, j1_0, W1, mb);
#pragma unroll
            for (int v1 = 0; v1 < V1; ++v1) {
#pragma unroll
                for (int k = 0; k < P; ++k) T1[q][v1][k] = T1[q][v1][k + 1];
                T r = T(0);
#pragma unroll
                for (int aq = 0; aq < WD; ++aq) r = fma(W1[v1][aq], c[aq], r);
                T1[q][v1][P] = r;
            }
        }
    };

    const bool full_tile = vec_ok && nv1 == V1;
    for (int s = 0; s < nstep; ++s) {
        const int s2 = s2s[s];
        if (s2 != cur) {
            const int nslide = (s2 == cur + 1) ? 1 : P + 1;
#pragma unroll 1
            for (int q = nslide - 1; q >= 0; --q) contract_row((int64_t)s2 - 1 - q);
            cur = s2;
        }
#pragma unroll
        for (int q = 0; q < ND; ++q) {
            const T *__restrict__ brow = b2s + ((size_t)q * chunk + s) * (P + 1);
            T acc[V1];
#pragma unroll
            for (int v1 = 0; v1 < V1; ++v1) acc[v1] = brow[0] * T1[q][v1][0];
#pragma unroll
            for (int k = 1; k <= P; ++k)
#pragma unroll
                for (int v1 = 0; v1 < V1; ++v1) acc[v1] = fma(brow[k], T1[q][v1][k], acc[v1]);
            T *__restrict__ dst = m.eval[q] + j1_0 + n1 * (j2_lo + s) + a.n_total * o;
            if (full_tile) sg_store_vec<T, V1>(dst, acc, true, V1);
            else sg_store_vec<T, V1>(dst, acc, false, nv1);
        }
    }
}
