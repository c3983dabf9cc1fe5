// sg_adjoint_post2.cuh -- second half of the 3-D double-march adjoint (K4, src/adjoint.jl:1-83) as ONE kernel:
// halo combine of the tile/chunk partials of dimensions 2 and 3 AND the contraction of dimension 1.
//
//   cp[i1, i2, i3, o] = sum_{j1 in range(i1)} B1[j1, i1 - base(j1)] * R[j1, i2, i3, o]
//   R[j1, i2, i3, o]  = sum over the (<= 2) tiles of dimension 2 and the chunks of dimension 3 that hold (i2, i3)
//                       of the partials written by sg_adj_march2(_tma)_kernel
//
// A CTA owns a block of 128 control indices i1, the G2 control rows i2 of one tile and one control plane i3.  It sums
// the partial rows that cover its G2 rows (slots 0..G2-1 of its own tile, the P halo slots of the previous tile, each
// from every chunk holding i3) with coalesced streaming loads into a shared-memory row buffer -- R never goes to
// HBM -- and then every thread gathers its control index over the samples of its support with the basis weights of
// dimension 1.  The order of every sum is fixed: deterministic, no atomics.
// The sample range of the i1 block is walked in pieces of SG_POST2_JMAX samples, so any n1 works.
// Replaces sg_adj_combine2_scan_kernel + sg_adj_first_dim_*_kernel (102 us -> one pass over the partials on C3).
#pragma once
#include <type_traits>
#include "sg_fast_adjoint.cuh"

#define SG_POST2_JMAX 512
#define SG_POST2_RMAX 20          // gather weights kept in registers per control index
#define SG_POST2_PITCH (SG_POST2_JMAX + SG_POST2_JMAX / 4 + 4)   // skewed row: sample jj lives at jj + (jj >> 2)

__device__ __forceinline__ void sg_discard_l2(const void *p)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
// grid = ((tiles2 + 1) * ceil(c1 / 128), c3, nout); requires G2 >= P (only neighbouring tiles overlap)
template <typename T, int P, int G2>
__global__ void __launch_bounds__(128, 4) sg_adj_post2_kernel(T *__restrict__ cp, const T *__restrict__ Pp, const T *__restrict__ table1,
                                                           const int32_t *__restrict__ index1, const int32_t *__restrict__ g_lo,
                                                           const T *__restrict__ g_w, const SgAdjointHeader *hdr, int64_t n1, int64_t c1, int64_t c2, int64_t c3, int P1,
                                                           int tiles2, int G3, int chunks3, int path,
                                                           const __grid_constant__ SgPushSpec push, int discard, int sf_known, int sl_known, int i3_top)
{
    constexpr int S = G2 + P;
    constexpr int JMAX = SG_POST2_JMAX;
    __shared__ T rows[G2][SG_POST2_PITCH];
    const int tid = threadIdx.x;
    const int t = (int)(blockIdx.x % (unsigned)(tiles2 + 1));          // tile of dimension 2 (tiles2 = the last tile's halo rows)
    const int64_t ib = blockIdx.x / (unsigned)(tiles2 + 1);            // block of control indices of dimension 1
    // planes in DESCENDING order: the march kernel wrote the high planes last, they are the ones still in L2
    // (i3_top = c3 with gridDim.y = c3: every plane; a push-only call launches the planes of the slab's support only)
    const int64_t i3 = (int64_t)i3_top - blockIdx.y;                    // 1-based control index of dimension 3
    const int64_t o = blockIdx.z;
    const int64_t i2_0 = (int64_t)t * G2;                               // 0-based control row of slot 0
    if (i2_0 >= c2) return;
    // Round 1 of loads, all independent of each other and of the header: this thread's entry of the gather table of
    // dimension 1 (prep kernel) -- first sample, number of samples, basis weights -- and the block's sample range.
    constexpr int RMAX = SG_POST2_RMAX;
    static_assert(SG_POST2_RMAX == SG_GATHER_RMAX, "gather table width");
    const int64_t i1 = ib * 128 + tid;                                   // 0-based control index of dimension 1
    const bool valid = i1 < c1;
    const int64_t i1c = valid ? i1 : c1 - 1;
    const int2 gl = *reinterpret_cast<const int2 *>(g_lo + 2 * i1c);
    const int lo_first = sg_ldg(g_lo + 2 * (ib * 128));
    const int64_t i_last = min(ib * 128 + 127, c1 - 1);
    const int2 gz = *reinterpret_cast<const int2 *>(g_lo + 2 * i_last);
    T w[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) w[r] = sg_ldg(g_w + (int64_t)r * c1 + i1c);
    // planned calls pass the slab's first / last span of dimension 3 as arguments (one dependent load chain less per CTA)
    const int sf = sf_known > 0 ? sf_known : hdr->span_first[2], sl = sf_known > 0 ? sl_known : hdr->span_last[2];
    // This kernel writes EVERY control point (no memset before the pipeline): zeros outside the support of this (slab
    // of the) grid, and zeros everywhere if the prep kernel flagged non-monotone spans (the scatter kernel accumulates).
    const bool write_local = push.world == 0 || push.keep_local != 0;
    if ((sf_known <= 0 && !sg_adj_path_active(hdr, path)) || i3 < sf - P || i3 > sl) {   // (a plan has seen monotone spans)
        const int64_t iz = ib * 128 + tid;
        if (iz < c1 && write_local) {
            T *__restrict__ out = cp + iz + c1 * (i2_0 + c2 * ((i3 - 1) + c3 * o));
#pragma unroll
            for (int q = 0; q < G2; ++q)
                if (i2_0 + q < c2) out[c1 * q] = T(0);
        }
        return;
    }
    const int rows3 = G3 + P;                                           // row stride of a chunk in the partials (host worst case)
    // the (<= 2, because G3e >= P) chunks of dimension 3 whose rows contain i3: chunk c holds the control indices
    // [sf + c*G3e - P, min(sf + (c+1)*G3e, sl + 1) - 1]  (sg_m2_chunk_len)
    int64_t off0 = 0, off1 = 0;
    int nc = 0;
    {
        const int G3e = max(max(P, 1), (sl - sf + 1 + chunks3 - 1) / chunks3);   // == sg_m2_chunk_len
        const int nch = (sl - sf + 1 + G3e - 1) / G3e;                 // chunks that hold samples
        const int c_lo = i3 >= sf ? (int)((i3 - sf) / G3e) : 0;
        const int c_hi = min((int)((i3 - sf + P) / G3e), nch - 1);
        for (int c = c_lo; c <= c_hi; ++c) {
            const int64_t l3 = i3 - ((int64_t)sf + (int64_t)c * G3e - P);
            // slot 0 of tile t in row l3 of chunk c
            const int64_t off = n1 * (int64_t)S * (t + (int64_t)tiles2 * (l3 + (int64_t)rows3 * (c + (int64_t)chunks3 * o)));
            if (nc == 0) off0 = off; else if (nc == 1) off1 = off;
            ++nc;
        }
        nc = min(nc, 2);
    }
    const bool own = t < tiles2, prev = t >= 1;                          // CTA-uniform sources

    // this thread's sample range [lo, hi) and the block's [jb_lo, jb_hi) (spans are monotone: the last control index
    // of the block ends last)
    const int64_t lo = gl.x, hi = (int64_t)gl.x + gl.y;
    const int len_i = valid ? gl.y : 0;
    const int64_t jb_lo = lo_first, jb_hi = (int64_t)gz.x + gz.y;
    // weights in registers (loaded above); ranges longer than RMAX, or a block range longer than one piece, take the
    // look-up loop below
    const bool regw = __syncthreads_and(len_i <= RMAX ? 1 : 0) && (jb_hi - jb_lo) <= JMAX;   // CTA-uniform
    T acc[G2];
#pragma unroll
    for (int q = 0; q < G2; ++q) acc[q] = T(0);

    for (int64_t jc = jb_lo; jc < jb_hi; jc += JMAX) {
        const int len = (int)min((int64_t)JMAX, jb_hi - jc);
        __syncthreads();                                                // the previous piece has been consumed
        // Every load of a sample column is issued before the first use (predicates are CTA-uniform compile-time
        // constants inside each instantiation): 2 columns x up to 2 chunks x (G2 + P) rows in flight per thread.
        auto load_piece = [&](auto OWN, auto PREV, auto TWO) {
            constexpr bool own_c = decltype(OWN)::value, prev_c = decltype(PREV)::value, two_c = decltype(TWO)::value;
#pragma unroll 4
            for (int jj = tid; jj < len; jj += 128) {
                const T *__restrict__ p0 = Pp + jc + jj + off0;
                const T *__restrict__ p1 = Pp + jc + jj + off1;
                T ra[G2], rb[P], sa[G2], sb[P];
#pragma unroll
                for (int q = 0; q < G2; ++q) { ra[q] = own_c ? __ldcs(p0 + n1 * q) : T(0); sa[q] = (own_c && two_c) ? __ldcs(p1 + n1 * q) : T(0); }
#pragma unroll
                for (int q = 0; q < P; ++q) {                           // slot G2 + q of tile t - 1
                    rb[q] = prev_c ? __ldcs(p0 - n1 * (S - G2 - q)) : T(0);
                    sb[q] = (prev_c && two_c) ? __ldcs(p1 - n1 * (S - G2 - q)) : T(0);
                }
                const int js = jj + (jj >> 2);
#pragma unroll
                for (int q = 0; q < G2; ++q) {
                    T v = ra[q] + sa[q];
                    if (q < P) v += rb[q] + sb[q];
                    rows[q][js] = v;
                }
                if (discard) {
                    // Every partial is read exactly once, here.  Its cache line is still dirty in L2 (the march kernel stored
                    // it with evict-last): discard it instead of letting the next march's stores push 135 MB of dead data
                    // out to HBM.  All lanes of the warp have consumed their loads once their shared-memory stores above
                    // are issued, so one lane per 128-byte line (16 / 32 consecutive columns) drops it.
                    __syncwarp();
                    constexpr int LINE = 128 / (int)sizeof(T);
                    if (((jc + jj) % LINE) == 0 && jj + LINE <= len) {
#pragma unroll
                        for (int q = 0; q < G2; ++q) {
                            if (own_c) {
                                sg_discard_l2(p0 + n1 * q);
                                if (two_c) sg_discard_l2(p1 + n1 * q);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            if (prev_c) {
                                sg_discard_l2(p0 - n1 * (S - G2 - q));
                                if (two_c) sg_discard_l2(p1 - n1 * (S - G2 - q));
                            }
                        }
                    }
                }
            }
        };
        using TT = std::true_type;
        using FF = std::false_type;
        if (nc == 2) {
            if (own && prev) load_piece(TT{}, TT{}, TT{}); else if (own) load_piece(TT{}, FF{}, TT{}); else load_piece(FF{}, TT{}, TT{});
        } else if (nc == 1) {
            if (own && prev) load_piece(TT{}, TT{}, FF{}); else if (own) load_piece(TT{}, FF{}, FF{}); else load_piece(FF{}, TT{}, FF{});
        } else {
            load_piece(FF{}, FF{}, FF{});                               // no chunk wrote this control plane: zeros
        }
        __syncthreads();
        if (regw) {
            if (valid) {
                const int j0 = (int)(lo - jc);                          // one piece: jc == jb_lo <= lo
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    if (r < len_i) {
                        const int jj = j0 + r, js = jj + (jj >> 2);
#pragma unroll
                        for (int q = 0; q < G2; ++q) acc[q] = fma(w[r], rows[q][js], acc[q]);
                    }
                }
            }
        } else if (valid) {
            const int64_t ja = max(lo, jc), jz = min(hi, jc + len);
            for (int64_t j = ja; j < jz; ++j) {
                const int k = (int)(i1 + 1 - sg_ldg(index1 + j) + P1);
                const T wj = sg_ldg(table1 + j + n1 * k);
                const int jj = (int)(j - jc), js = jj + (jj >> 2);
#pragma unroll
                for (int q = 0; q < G2; ++q) acc[q] = fma(wj, rows[q][js], acc[q]);
            }
        }
    }
    if (valid) {
        if (write_local) {
            T *__restrict__ out = cp + i1 + c1 * (i2_0 + c2 * ((i3 - 1) + c3 * o));
#pragma unroll
            for (int q = 0; q < G2; ++q)
                if (i2_0 + q < c2) out[c1 * q] = acc[q];
        }
        // Fused gradient push (slab-sharded grids): control plane i3 is plane l = i3 - (sf - P) of this rank's support;
        // it goes into this rank's slot of every rank's staging buffer with peer-to-peer stores (256 contiguous bytes per
        // warp and row), overlapping the rest of the kernel.  sg_exchange_reduce sums the slots after the barrier.
        if (push.world > 0) {
            const int64_t l = i3 - (sf - P);
            if (l >= 0 && l < push.max_planes) {
                const int64_t plane_elems = c1 * c2;
                const int64_t off = push.max_planes * plane_elems * (o + (int64_t)gridDim.z * push.my_rank) + plane_elems * l + i1 + c1 * i2_0;
#pragma unroll 1
                for (int r = 0; r < push.n_dst; ++r) {
                    if (l < push.dst_lo[r] || l >= push.dst_hi[r]) continue;
                    T *__restrict__ st = static_cast<T *>(push.stage[r]) + off;
#pragma unroll
                    for (int q = 0; q < G2; ++q)
                        if (i2_0 + q < c2) st[c1 * q] = acc[q];
                }
            }
        }
    }
}
