// sg_adjoint_march2g.cuh -- 3-D evaluate_adjoint! (K4, src/adjoint.jl:1-83) with ALL THREE contractions in the
// TMA-fed double march: the kernel of sg_fast_adjoint.cuh (dimensions 2 and 3 in registers) plus an epilogue that
// contracts dimension 1 across the CTA's 128 sample columns every time a control plane of dimension 3 is finished.
//
//   The unfused pipeline writes one partial value per SAMPLE column (n1 per row: 135 MB on C3) and the post kernel
//   reads them back; here a column block leaves only the NI ~ 128 (c1 - p1) / n1 + p1 control indices it touches
//   (36 instead of 128 on C3: 41 MB), and the second kernel is a plain halo sum.
//
// Epilogue (per finished plane, executed by the 128 consumer threads themselves, ~15 % extra instructions):
//   1. every thread parks its S = G2 + P finished values in a shared-memory row buffer E[S][.], column c at position
//      c + c/4 (lanes that gather consecutive control indices read ~4 columns apart: stride 5 -> conflict-free);
//   2. thread (li, sg) gathers control index li over the columns of its support INSIDE the block,
//      out[li][s] = sum_r w[r][li] * E[s][lo_li + r], for up to three slot rows s per pass; the weights (gather form
//      of B1, clipped to the block) come from a per-block table built once by the prep kernel (L1-resident, 6 KB);
//   3. out[li][s] goes to the partials with li fastest (coalesced).
// The sum order is fixed: deterministic, no atomics.  Any number of samples per knot span of dimension 1 works as long
// as the block table fits (NI <= icap, support length <= rmcap; checked on device by the prep kernel, which raises
// hdr->m2g_bad otherwise -- planned calls then use the unfused pipeline).
// sg_adj_combine2g_kernel sums the <= 2 x 2 x (few) partial rows that hold a control point (tile halo of dimension 2,
// chunk halo of dimension 3, column blocks of dimension 1) and writes EVERY control point (zeros outside a slab's
// support), with the fused peer push of sg_adjoint_post2.cuh.
#pragma once
#include <type_traits>
#include "sg_fast_adjoint.cuh"

#define SG_M2G_EP (128 + 32)      // row pitch of the epilogue buffer: column c lives at c + (c >> 2)
#define SG_M2G_MAXPL 64           // planes of dimension-3 tables staged at a time

template <typename T>
struct SgAdj2gArgs {
    const T *X;                 // eval (n1, n2, n3, nout)
    T *Y;                       // partials [icap][S][tiles2][rows3][chunks3][nb1][nout]
    const T *table2, *table3;   // (n2, P+1), (n3, P+1) selected derivative slices
    const int32_t *index3;
    const int32_t *start2, *start3;
    const SgAdjointHeader *hdr;
    const SgM2gBlockHdr *bt_hdr;   // [nb1]
    const int32_t *bt_lol;         // [nb1][icap]   (lo_rel | len << 16)
    const T *bt_w;                 // [nb1][rmcap][icap]
    int64_t n1, n2, n3, c2, c3;
    int tiles2, G3, chunks3, icap, rmcap, nb1;
};

template <typename T, int P, int G2, int RTMAX, int NS, bool UW>
__global__ void __launch_bounds__(160, 3) sg_adj_march2g_kernel(const __grid_constant__ SgAdj2gArgs<T> a, const __grid_constant__ SgM2Maps maps,
                                                                const __grid_constant__ typename std::conditional<UW, SgM2Uni<T>, SgM2UniNone<T>>::type uni)
{
    constexpr int S = G2 + P;
    constexpr int MAXPL = SG_M2G_MAXPL;
    constexpr int CW = 128;
    constexpr int EP = SG_M2G_EP;
    constexpr int RS5 = SG_M2_FAST_ROWS;
    extern __shared__ __align__(16) unsigned char sg_smem2g[];
    T *xs = reinterpret_cast<T *>(sg_smem2g + ((128u - (sg_smem_u32(sg_smem2g) & 127u)) & 127u));   // [NS][RTMAX][CW], 128-byte aligned
    T *Ebuf = xs + (size_t)NS * RTMAX * CW;                             // [S][EP]
    __shared__ __align__(16) T b3s[MAXPL * (P + 1)];
    __shared__ int s3s[MAXPL];
    __shared__ int row0[G2 + 1];
    __shared__ __align__(16) T b2pad[UW ? 1 : G2 * RS5 * (P + 1)];
    __shared__ int lol_s[CW + 8];                                       // (lo_rel | len << 16) per local control index
    __shared__ __align__(8) uint64_t full[NS];
    __shared__ __align__(8) uint64_t empty[NS];

    const int tid = threadIdx.x;
    // the warp index through a shuffle: the compiler then knows that the role branch and everything the loops below are
    // bounded by is warp-uniform, and may use the uniform datapath inside them
    const bool is_producer = __shfl_sync(0xffffffffu, tid >> 5, 0) >= CW / 32;
    const int jb = blockIdx.x;
    const int64_t j1_0 = (int64_t)jb * CW;
    const int tile2 = blockIdx.y;
    const int c3k = blockIdx.z % a.chunks3;
    const int64_t o = blockIdx.z / a.chunks3;

    const int s2_lo = P + 1 + tile2 * G2;
    // rows of the tile's spans: [ur0[g], ur0[g + 1]).  UW: from the kernel parameter (uniform registers), else from memory.
    int ur0[G2 + 1];
    if constexpr (UW) {
#pragma unroll
        for (int g = 0; g <= G2; ++g) ur0[g] = uni.start2[min(s2_lo + g, (int)a.c2 + 1)];
    } else {
        if (tid <= G2) row0[tid] = a.start2[(int)min((int64_t)s2_lo + tid, a.c2 + 1)];
    }
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) { sg_mbar_init(&full[q], 1); sg_mbar_init(&empty[q], CW / 32); }
    }
    __syncthreads();
    if constexpr (!UW) {
#pragma unroll
        for (int g = 0; g <= G2; ++g) ur0[g] = row0[g];
    }
    const int r_first = ur0[0], n_rows = ur0[G2] - ur0[0];
    const SgM2gBlockHdr bh = a.bt_hdr[jb];
    const int ni = bh.ni;
    if (!is_producer) {
        if constexpr (!UW) {
            for (int q = tid; q < G2 * RS5 * (P + 1); q += CW) {
                const int k = q % (P + 1), gq = q / (P + 1), g = gq / RS5, qq = gq % RS5;
                b2pad[q] = qq < row0[g + 1] - row0[g] ? sg_ldg(a.table2 + (row0[g] + qq) + a.n2 * k) : T(0);
            }
            for (int sl = 0; sl < NS * G2 * RS5; ++sl) {                // absent row slots of every stage: zero once
                const int gq = sl % (G2 * RS5), g = gq / RS5, qq = gq % RS5;
                if (qq >= row0[g + 1] - row0[g]) xs[(size_t)(sl / (G2 * RS5)) * RTMAX * CW + (size_t)gq * CW + tid] = T(0);
            }
        }
        for (int q = tid; q < ni; q += CW) lol_s[q] = sg_ldg(a.bt_lol + (int64_t)jb * a.icap + q);
    }

    int sf3, sl3;
    if constexpr (UW) { sf3 = uni.span_first3; sl3 = uni.span_last3; }
    else { sf3 = a.hdr->span_first[2]; sl3 = a.hdr->span_last[2]; }
    const int G3e = max(max(P, 1), (sl3 - sf3 + 1 + a.chunks3 - 1) / a.chunks3);   // == sg_m2_chunk_len
    const int s3_lo = sf3 + c3k * G3e;
    const int s3_hi = min(s3_lo + G3e, sl3 + 1);
    if (s3_lo >= s3_hi) return;                                        // block-uniform
    int64_t j3_lo, j3_hi;
    if constexpr (UW) { j3_lo = uni.start3[s3_lo]; j3_hi = uni.start3[s3_hi]; }
    else { j3_lo = a.start3[s3_lo]; j3_hi = a.start3[s3_hi]; }
    const int np_total = (int)(j3_hi - j3_lo);
    const int rows3 = a.G3 + P;

    int st = 0;
    unsigned ph = 0;
    if (is_producer) {
        if (n_rows > 0) {                                               // whole warp, converged; one elected lane issues
            const uint32_t xs_u = sg_smem_u32(xs), full_u = sg_smem_u32(full), empty_u = sg_smem_u32(empty);
            int ra[G2], rn[G2];
#pragma unroll
            for (int g = 0; g < G2; ++g) { ra[g] = ur0[g] - r_first; rn[g] = ur0[g + 1] - ur0[g]; }
            const unsigned stage_tx = (unsigned)(CW * sizeof(T)) * (unsigned)n_rows;
            const int pl0 = (int)(a.n3 * o + j3_lo);
            for (int p = 0; p < np_total; ++p) {
                if (p >= NS) sg_m2_mbar_wait_u(empty_u + (uint32_t)st * 8u, ph ^ 1u);
                const uint32_t fb = full_u + (uint32_t)st * 8u;
                sg_m2_expect_tx_elect(fb, stage_tx);
                const uint32_t dst = xs_u + (uint32_t)st * (uint32_t)(RTMAX * CW * sizeof(T));
#pragma unroll
                for (int g = 0; g < G2; ++g)
                    if (rn[g] > 0)                                      // warp-uniform
                        sg_m2_tma_load_3d_elect(dst + (uint32_t)(g * RS5 * CW * sizeof(T)), &maps.m[rn[g] - 1], (int)j1_0, r_first + ra[g], pl0 + p, fb);
                if (++st == NS) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------------------
    T acc3[S][P + 1];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int k = 0; k <= P; ++k) acc3[s][k] = T(0);
    int cur = s3_lo;

    // epilogue roles: thread (li0, sg) gathers control index li0 (+ 128, ...) for the slot rows sg, sg + tpl, ...
    const int tpl = max(1, min(S, CW / max(ni, 1)));                    // threads per control index (CTA-uniform)
    const int sg = ni <= CW ? tid / max(ni, 1) : 0;
    const int li0 = ni <= CW ? tid - sg * ni : tid;
    const bool epi = sg < tpl;
    const int epos = tid + (tid >> 2);                                  // this column's position in a row of E
    const T *__restrict__ wblk = a.bt_w + (int64_t)jb * a.rmcap * a.icap;
    const unsigned icap = (unsigned)a.icap;
    const unsigned y_row3 = icap * (unsigned)(S * a.tiles2);
    unsigned yoff = icap * (unsigned)S * (unsigned)tile2 +
                    y_row3 * (unsigned)rows3 * ((unsigned)c3k + (unsigned)a.chunks3 * ((unsigned)jb + (unsigned)a.nb1 * (unsigned)o));
    T *__restrict__ const ybase = a.Y;

    auto emit_oldest = [&]() {
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");           // the previous plane's gather is over
#pragma unroll
        for (int s = 0; s < S; ++s) {
            Ebuf[s * EP + epos] = acc3[s][0];
#pragma unroll
            for (int k = 0; k < P; ++k) acc3[s][k] = acc3[s][k + 1];
            acc3[s][P] = T(0);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");
        if (epi) {
            for (int li = li0; li < ni; li += CW) {
                const int ll = lol_s[li];
                const int lo = ll & 0xffff, len = ll >> 16;
                const T *__restrict__ wp = wblk + li;
                for (int sb = sg; sb < S; sb += 3 * tpl) {
                    const int s1 = sb + tpl, s2 = sb + 2 * tpl;
                    const T *__restrict__ e0 = Ebuf + sb * EP;
                    const T *__restrict__ e1 = Ebuf + min(s1, S - 1) * EP;
                    const T *__restrict__ e2 = Ebuf + min(s2, S - 1) * EP;
                    T a0 = T(0), a1 = T(0), a2 = T(0);
#pragma unroll 4
                    for (int r = 0; r < len; ++r) {
                        const int c = lo + r, pos = c + (c >> 2);
                        const T wr = sg_ldg(wp + (unsigned)r * icap);
                        a0 = fma(wr, e0[pos], a0);
                        a1 = fma(wr, e1[pos], a1);
                        a2 = fma(wr, e2[pos], a2);
                    }
                    T *__restrict__ yp = ybase + (yoff + (unsigned)li);
                    __stcs(yp + icap * (unsigned)sb, a0);
                    if (s1 < S) __stcs(yp + icap * (unsigned)s1, a1);
                    if (s2 < S) __stcs(yp + icap * (unsigned)s2, a2);
                }
            }
        }
        yoff += y_row3;
        ++cur;
    };

    for (int p0 = 0; p0 < np_total; p0 += MAXPL) {
        const int p1 = min(p0 + MAXPL, np_total);
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");
        for (int s = tid; s < p1 - p0; s += CW) {
            s3s[s] = sg_ldg(a.index3 + j3_lo + p0 + s);
#pragma unroll
            for (int k = 0; k <= P; ++k) b3s[s * (P + 1) + k] = sg_ldg(a.table3 + j3_lo + p0 + s + a.n3 * k);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CW) : "memory");
        for (int p = p0; p < p1; ++p) {
            T T2[S];
#pragma unroll
            for (int q = 0; q < S; ++q) T2[q] = T(0);
            if (n_rows > 0) {
                sg_mbar_wait(&full[st], ph);
                const T *__restrict__ xst = xs + (size_t)st * RTMAX * CW + tid;
#pragma unroll
                for (int g = 0; g < G2; ++g) {
#pragma unroll
                    for (int q = 0; q < RS5; ++q) {
                        if constexpr (UW) {
                            if (q < ur0[g + 1] - ur0[g]) {              // uniform predicate: absent slots cost nothing
                                const T x = xst[(g * RS5 + q) * CW];
#pragma unroll
                                for (int k = 0; k <= P; ++k) T2[g + k] = fma(uni.b2[(ur0[g] + q) * (P + 1) + k], x, T2[g + k]);
                            }
                        } else {
                            const T x = xst[(g * RS5 + q) * CW];
#pragma unroll
                            for (int k = 0; k <= P; ++k) T2[g + k] = fma(b2pad[(g * RS5 + q) * (P + 1) + k], x, T2[g + k]);
                        }
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) sg_mbar_arrive(&empty[st]);
                if (++st == NS) { st = 0; ph ^= 1u; }
            }
            const int sl = p - p0;
            const int sp = s3s[sl];
            if (cur < sp) {                                             // CTA-uniform: every consumer takes part in the epilogue
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
    while (cur < s3_hi + P) emit_oldest();                              // the chunk's last span and its P still-live planes
}

// ---- halo sum ------------------------------------------------------------------------------------------------
// grid = ((tiles2 + 1) * ceil(c1 / 128), c3, nout), 128 threads: thread = control index i1, CTA = the G2 control rows of
// one tile of dimension 2 and one control plane i3.
template <typename T, int P, int G2>
__global__ void __launch_bounds__(128) sg_adj_combine2g_kernel(T *__restrict__ cp, const T *__restrict__ Pp, const int32_t *__restrict__ g_lo,
                                                               const SgM2gBlockHdr *__restrict__ bt_hdr, const SgAdjointHeader *hdr,
                                                               int64_t c1, int64_t c2, int64_t c3, int tiles2, int G3, int chunks3, int icap, int nb1,
                                                               const __grid_constant__ SgPushSpec push)
{
    constexpr int S = G2 + P;
    const int tid = threadIdx.x;
    const int t = (int)(blockIdx.x % (unsigned)(tiles2 + 1));
    const int64_t ib = blockIdx.x / (unsigned)(tiles2 + 1);
    const int64_t i3 = (int64_t)blockIdx.y + 1;                         // 1-based control index of dimension 3
    const int64_t o = blockIdx.z;
    const int64_t i2_0 = (int64_t)t * G2;
    if (i2_0 >= c2) return;
    const int64_t i1 = ib * 128 + tid;                                   // 0-based control index of dimension 1
    if (i1 >= c1) return;
    const int2 gl = *reinterpret_cast<const int2 *>(g_lo + 2 * i1);     // support of i1: first sample, number of samples
    const int sf = hdr->span_first[2], sl = hdr->span_last[2];
    const bool write_local = push.world == 0 || push.keep_local != 0;
    T acc[G2];
#pragma unroll
    for (int q = 0; q < G2; ++q) acc[q] = T(0);
    const bool in_support = i3 >= sf - P && i3 <= sl;
    if (in_support && gl.y > 0) {
        const int rows3 = G3 + P;
        const int G3e = sg_m2_chunk_len(hdr, P, chunks3);
        const int nch = (sl - sf + 1 + G3e - 1) / G3e;
        const int c_lo = i3 >= sf ? (int)((i3 - sf) / G3e) : 0;
        const int c_hi = min((int)((i3 - sf + P) / G3e), nch - 1);
        const int jb_a = gl.x >> 7, jb_b = min((gl.x + gl.y - 1) >> 7, nb1 - 1);
        const bool own = t < tiles2, prev = t >= 1;
        for (int c = c_lo; c <= c_hi; ++c) {
            const int64_t l3 = i3 - ((int64_t)sf + (int64_t)c * G3e - P);
            for (int jbk = jb_a; jbk <= jb_b; ++jbk) {
                const int li = (int)(i1 + 1) - bt_hdr[jbk].i1_lo;
                if (li < 0 || li >= icap) continue;                     // (cannot happen for monotone spans)
                const T *__restrict__ p0 = Pp + li + (int64_t)icap * S * (t + (int64_t)tiles2 * (l3 + (int64_t)rows3 * (c + (int64_t)chunks3 * (jbk + (int64_t)nb1 * o))));
                if (own) {
#pragma unroll
                    for (int q = 0; q < G2; ++q) acc[q] += __ldcs(p0 + icap * q);
                }
                if (prev) {
#pragma unroll
                    for (int q = 0; q < P; ++q) acc[q] += __ldcs(p0 - icap * (S - G2 - q));
                }
            }
        }
    }
    if (write_local) {
        T *__restrict__ out = cp + i1 + c1 * (i2_0 + c2 * ((i3 - 1) + c3 * o));
#pragma unroll
        for (int q = 0; q < G2; ++q)
            if (i2_0 + q < c2) out[c1 * q] = acc[q];
    }
    if (push.world > 0 && in_support) {
        const int64_t l = i3 - (sf - P);
        if (l >= 0 && l < push.max_planes) {
            const int64_t plane_elems = c1 * c2;
            const int64_t off = push.max_planes * plane_elems * (o + (int64_t)gridDim.z * push.my_rank) + plane_elems * l + i1 + c1 * i2_0;
#pragma unroll 1
            for (int r = 0; r < push.n_dst; ++r) {
                if (l < push.dst_lo[r] || l >= push.dst_hi[r]) continue;
                T *__restrict__ stg = static_cast<T *>(push.stage[r]) + off;
#pragma unroll
                for (int q = 0; q < G2; ++q)
                    if (i2_0 + q < c2) stg[c1 * q] = acc[q];
            }
        }
    }
}
