// sg_adjoint_march3.cuh -- evaluate_adjoint! (K4) for 3-D grids as ONE streaming pass over the sample array.
//
//   cp[i1,i2,i3,o] = sum_{j1,j2,j3} B1[j1,i1] B2[j2,i2] B3[j3,i3] eval[j1,j2,j3,o]       (src/adjoint.jl:1-83)
//
// All three dimensions are contracted inside the CTA, so the only large HBM traffic is ONE read of eval:
//   * a CTA ("worker") owns CW = 64 consecutive samples of dimension 1 (a column block b1), a tile t2 of G2 = 4 whole
//     knot spans of dimension 2 and a run of sample planes j3.  The tile's rows of each plane are streamed into an
//     NS-stage shared-memory ring by a dedicated producer warp with bulk async copies (cp.async.bulk, SASS UBLKCP;
//     full/empty mbarriers per stage, no block-wide barrier in the plane loop); stage row (g, q) = q-th row of span g;
//   * consumer thread (column, span g) MARCHES DIMENSION 3 on the raw rows of its span: acc[q][k] += B3[j3,k] * x[q],
//     RPT*(P+1) independent FMAs per plane, the P+1 live control planes per row in registers;
//   * whenever the knot span of dimension 3 advances, one control plane is complete for every (row, column) of the
//     tile; only then it is contracted over dimension 2 (per span, then the <= P+1 span partials of a control slot are
//     added through shared memory) and over dimension 1 (per knot span of the block, then per control index): the
//     separable contractions of dimensions 2 and 1 run once per knot span of dimension 3 instead of once per plane, and
//     only NI1 x S2 numbers per completed control plane leave the CTA.
//   Spans of dimension 2 with more than RPT rows are finished in further passes over the segment that stream only the
//   remaining rows (every sample is still read exactly once) and add to the partials of the first pass.
// Work is split over a PERSISTENT grid of W workers by a linear partition of the (column block, tile, plane) space
// weighted by the rows of each tile, so every worker streams the same number of bytes (no wave quantisation, also for
// the thin slabs of a sharded grid).  A worker's run inside one (block, tile) column is a "segment"; segment k of a
// column stores its control planes at row (i3-1) + k(P+1), block b at (i1-1) + b(P1+1): halos never collide and no
// size depends on the data.  Which segments hold a control plane of a column is recorded as a bit mask (integer OR on
// a small table: order-independent); sg_adj_combine3_kernel then adds the partials of every control point in a fixed
// order: no floating-point atomics, deterministic.  Pathological sample distributions (a column with more than
// `maxseg` segments) raise the header flag and the reference's atomic scatter kernel takes over (decided on device).
#pragma once
#include "sg_fast_adjoint.cuh"

#define SG_M3_CW 128          // columns (samples of dimension 1) per worker
#define SG_M3_G2 4            // knot spans of dimension 2 per tile; consumer threads = CW * G2
#define SG_M3_RPT 5           // rows of a span handled per pass (stage rows = G2 * RPT)
#define SG_M3_NS 5            // ring stages
#define SG_M3_NSPMAX 40       // knot spans of dimension 1 per column block handled by the fast contraction
#define SG_M3_PITCH (SG_M3_CW + SG_M3_CW / 4 + 4)   // skewed row of the park buffer: index j + (j >> 2)
#define SG_M3_MAXPL 64        // planes per staged piece of the dimension-3 tables
#define SG_M3_NCONS (SG_M3_CW * SG_M3_G2)
#define SG_M3_THREADS (SG_M3_NCONS + 32)   // consumers + producer warp

template <typename T>
struct SgAdj3Args {
    const T *X;                 // eval (n1, n2, n3, nout)
    T *part;                    // partials [L1][S2][tiles2][L3][nout]
    unsigned *rowmask;          // [ncols][c3]: bit k set = segment k of the column holds this control plane
    const T *table1, *table2, *table3;   // selected derivative slices (n_d, P_d+1), column-major
    const int32_t *index1, *index3;      // span per sample (1-based)
    const int32_t *start1, *start2;      // span_start arrays (prep kernel)
    SgAdjointHeader *hdr;
    int64_t n1, n2, n3, c1, c2, c3;
    int P1;
    int nb1, tiles2, nout;      // columns = nout * nb1 * tiles2, tile index fastest
    int L1, L3, maxseg;
};

// TMA tensor maps over eval viewed as (n1, n2, n3*nout): box (CW, h, 1) for h = 16, 8, 4, 2, 1.  A run of rows of one
// plane is fetched as the binary decomposition of its length: 1-2 bulk tensor copies per stage instead of one per row.
#define SG_M3_NMAPS 5
struct alignas(64) SgM3Maps {
    CUtensorMap m[SG_M3_NMAPS];   // m[i]: box height 16 >> i
};
// Executed by the WHOLE (converged) producer warp with warp-uniform operands; one elected lane issues.  Keeping the warp
// converged lets the operands live in uniform registers (a lane-divergent caller costs ~25 instructions per copy).
__device__ __forceinline__ void sg_tma_load_3d_elect(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar)
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
__device__ __forceinline__ void sg_mbar_expect_tx_elect(uint32_t bar, unsigned bytes)
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

__device__ __forceinline__ void sg_m3_consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SG_M3_NCONS) : "memory"); }

// mbarrier wait / arrive on raw 32-bit shared addresses (no address re-derivation in the hot loop)
__device__ __forceinline__ void sg_mbar_wait_u(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAITU_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONEU_%=;\n"
        "bra SG_WAITU_%=;\n"
        "SG_DONEU_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sg_mbar_arrive_u(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// value the compiler must keep in a register instead of re-deriving it (from threadIdx etc.) at every use
__device__ __forceinline__ uint32_t sg_opaque(uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }

// explicit shared-space loads (32-bit shared addresses)
__device__ __forceinline__ double sg_lds(uint32_t a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ float sg_lds(uint32_t a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sg_lds2(uint32_t a, double &w0, double &w1) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w0), "=d"(w1) : "r"(a)); }
__device__ __forceinline__ void sg_lds2(uint32_t a, float &w0, float &w1) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w0), "=f"(w1) : "r"(a)); }
__device__ __forceinline__ void sg_lds4(uint32_t a, double (&w)[4])
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(w[2]), "=d"(w[3]) : "r"(a));
}
__device__ __forceinline__ void sg_lds4(uint32_t a, float (&w)[4])
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "r"(a));
}

// Linear partition: iterates over the segments of worker w.  Every thread that needs the sequence (the producer warp and
// consumer thread 0) runs its own copy.  The unit of work is one pass over one plane of one (block, tile) column -- the
// cost of a plane hardly depends on the number of rows of the tile, only on the number of passes it needs -- so a column
// weighs cumw[t+1] - cumw[t] = passes(t) (0 for tiles without samples) per plane; cumw lives in shared memory.
#define SG_M3_MAXTILES 1024
struct SgM3SegIter {
    int64_t o_begin, o_end, tot, tot1, n3, nob;
    const int *cumw;
    int W, w, tiles2;
    int64_t ob;      // index over (o, b1)
    int t;           // tile of the NEXT candidate column
    // current segment
    int64_t col_ob;
    int col_t, p0, p1, kseg;

    __device__ void init(int w_, int W_, int64_t nob_, int64_t n3_, int tiles2_, const int *cumw_)
    {
        W = W_; w = w_; n3 = n3_; tiles2 = tiles2_; cumw = cumw_; nob = nob_;
        tot1 = (int64_t)cumw[tiles2] * n3;
        tot = nob * tot1;
        o_begin = (int64_t)w * tot / W;
        o_end = (int64_t)(w + 1) * tot / W;
        ob = tot1 > 0 ? o_begin / tot1 : nob;
        const int64_t rq = tot1 > 0 ? (o_begin % tot1) / n3 : 0;   // largest t with cumw[t] <= rq
        int lo = 0, hi = tiles2 - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (cumw[mid] <= rq) lo = mid; else hi = mid - 1;
        }
        t = lo;
        if (o_begin >= o_end) ob = nob;                  // empty worker
    }

    __device__ bool next()
    {
        while (ob < nob) {
            const int c0 = cumw[t], c1 = cumw[t + 1];
            const int64_t A = (ob * cumw[tiles2] + c0) * n3;
            if (A >= o_end) return false;
            const int wt = c1 - c0;
            const int64_t my_ob = ob;
            const int my_t = t;
            if (++t == tiles2) { t = 0; ++ob; }
            if (wt <= 0) continue;
            // plane p of the column belongs to the worker that holds its first unit A + p*wt
            const int64_t d0 = o_begin - A, d1 = o_end - A;
            const int64_t q0 = d0 <= 0 ? 0 : (d0 + wt - 1) / wt;
            const int64_t q1 = (d1 + wt - 1) / wt;
            const int pp0 = (int)min(q0, n3), pp1 = (int)min(q1, n3);
            if (pp0 >= pp1) continue;
            col_ob = my_ob; col_t = my_t; p0 = pp0; p1 = pp1;
            const int64_t w_first = ((A + 1) * W + tot - 1) / tot - 1;   // worker that holds unit A
            kseg = (int)(w - w_first);
            return true;
        }
        return false;
    }
};

// Descriptor of a control plane travelling through the post pipeline
struct SgM3RowMeta {
    long long base_off;   // element offset of the partial row (without the block's first control index)
    int b1;               // column block
    int flags;            // bit 0: store, bit 1: first pass (store; else add), bit 2: valid row, bit 3: block-constants buffer
};
// Column-block constants used by the dimension-1 contraction (two blocks can be in flight)
struct SgM3Block {
    int lo_b, nsp_b, ni1, fast, ncols;
    long long j1_0;
};

template <typename T, int P>
__global__ void __launch_bounds__(SG_M3_THREADS, 1) sg_adj_march3_kernel(const __grid_constant__ SgAdj3Args<T> a, const __grid_constant__ SgM3Maps maps)
{
    if (a.hdr->nonmonotone) return;
    constexpr int G2 = SG_M3_G2;
    constexpr int S2 = G2 + P;
    constexpr int CW = SG_M3_CW;
    constexpr int NCONS = SG_M3_NCONS;
    constexpr int NW = NCONS / 32;                                      // consumer warps
    constexpr int RPT = SG_M3_RPT;
    constexpr int NROWS = G2 * RPT;                                     // row slots of a ring stage
    constexpr int NS = SG_M3_NS;
    constexpr int NQ = 3;                                               // row buffers of the post pipeline
    constexpr int PITCH = SG_M3_PITCH;
    constexpr int MAXPL = SG_M3_MAXPL;
    constexpr int NSPMAX = SG_M3_NSPMAX;
    static_assert(NROWS <= 31, "a run of stage rows is decomposed in boxes of 16, 8, 4, 2, 1 rows");
    static_assert((CW & (CW - 1)) == 0, "CW must be a power of two");
    static_assert(NSPMAX * S2 <= NCONS && (NSPMAX + 3) * S2 <= NCONS, "one dimension-1 unit / output per consumer thread");
    static_assert(S2 <= 2 * G2, "two control slots per consumer thread");
    extern __shared__ __align__(16) unsigned char sg_smem3[];          // NOT declared 128-aligned: the compiler would fold the fix-up below
    // TMA destinations must be 128-byte aligned in the shared window (the dynamic region starts after the static
    // variables at an offset that is only 16-byte aligned): the host adds 128 bytes of slack
    T *ring = reinterpret_cast<T *>(sg_smem3 + ((128u - (sg_smem_u32(sg_smem3) & 127u)) & 127u));   // [NS][NROWS][CW]
    T *tq = ring + (size_t)NS * NROWS * CW;                             // [NQ][G2][4][CW]    span partials of dimension 2
    T *park = tq + (size_t)NQ * G2 * 4 * CW;                            // [NQ][S2][PITCH]    control slots x columns (skewed)
    T *a4 = park + (size_t)NQ * S2 * PITCH;                             // [NQ][S2][4][NSPMAX] span sums of dimension 1
    T *wAB = a4 + (size_t)NQ * S2 * 4 * NSPMAX;                         // [2][2][PITCH][2]   B1[j, 0:2], B1[j, 2:4] per block parity
    T *b2s = wAB + 2 * 4 * PITCH;                                       // [NROWS][4]         B2 rows of the stage rows (0 if absent)
    T *b3s = b2s + NROWS * 4;                                           // [MAXPL][4]
    int *s3s = reinterpret_cast<int *>(b3s + MAXPL * 4);                // [MAXPL]
    __shared__ __align__(8) uint64_t full[NS];
    __shared__ __align__(8) uint64_t empty[NS];
    __shared__ __align__(8) uint64_t pbar[3][NQ];                       // phase k of the row in buffer q is complete (all warps)
    __shared__ SgM3RowMeta qmeta[8];                                    // by (row step & 7): a warp may lag almost two steps behind the writer
    __shared__ SgM3Block blk[2];
    __shared__ int cumw[SG_M3_MAXTILES + 1];                            // prefix sums of the passes per plane of every tile

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) { sg_mbar_init(&full[q], 1); sg_mbar_init(&empty[q], NW); }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < NQ; ++q) sg_mbar_init(&pbar[k][q], NW);
    }
    for (int t = threadIdx.x; t < a.tiles2; t += blockDim.x) {          // passes of tile t: ceil(longest span / RPT)
        int maxnr = 0;
#pragma unroll
        for (int gg = 0; gg < G2; ++gg) {
            const int r_a = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * t + gg, a.c2 + 1)];
            const int r_b = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * t + gg + 1, a.c2 + 1)];
            maxnr = max(maxnr, r_b - r_a);
        }
        cumw[t + 1] = (maxnr + RPT - 1) / RPT;
    }
    __syncthreads();
    if (tid == 0) {
        cumw[0] = 0;
        for (int t = 0; t < a.tiles2; ++t) cumw[t + 1] += cumw[t];
    }
    __syncthreads();

    if (tid >= NCONS) {
        // ======================= producer warp =======================
        // The whole warp runs this loop converged; single-thread operations elect a lane inside the asm.
        // Single-pass tiles (no span with more than RPT rows): the tile's rows are contiguous in the plane and land at
        // stage row (row - first row of the tile).  Otherwise every span's rows of the pass land at stage row g*RPT.
        // Each run of rows is fetched as the binary decomposition of its length in TMA boxes (1-2 copies per stage).
        SgM3SegIter it;
        it.init((int)blockIdx.x, (int)gridDim.x, (int64_t)a.nout * a.nb1, a.n3, a.tiles2, cumw);
        int st = 0;
        unsigned ph = 0;
        bool first_round = true;
        const uint32_t ring_s = sg_smem_u32(ring), full_s = sg_smem_u32(full), empty_s = sg_smem_u32(empty);
        while (it.next()) {
            const int64_t o = it.col_ob / a.nb1;
            const int b1 = (int)(it.col_ob % a.nb1);
            const int j1_0 = b1 * CW;
            int rg[G2 + 1];
#pragma unroll
            for (int gg = 0; gg <= G2; ++gg) rg[gg] = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * it.col_t + gg, a.c2 + 1)];
            int maxnr = 0;
#pragma unroll
            for (int gg = 0; gg < G2; ++gg) maxnr = max(maxnr, rg[gg + 1] - rg[gg]);
            const bool contiguous = maxnr <= RPT;
            const int pl0 = (int)(a.n3 * o) + it.p0, np = it.p1 - it.p0;
            for (int r_off = 0; r_off < maxnr; r_off += RPT) {            // passes over the segment
                int nq[G2], cnt = 0;
#pragma unroll
                for (int gg = 0; gg < G2; ++gg) { nq[gg] = max(0, min(RPT, rg[gg + 1] - rg[gg] - r_off)); cnt += nq[gg]; }
                if (cnt == 0) continue;
                const unsigned stage_tx = (unsigned)(CW * sizeof(T)) * (unsigned)cnt;   // boxes are CW wide (zero fill past n1)
                for (int p = 0; p < np; ++p) {
                    if (!first_round) sg_mbar_wait_u(empty_s + (uint32_t)st * 8u, ph ^ 1u);   // consumers have released the stage
                    const uint32_t fb = full_s + (uint32_t)st * 8u;
                    sg_mbar_expect_tx_elect(fb, stage_tx);
                    const uint32_t dst0 = ring_s + (uint32_t)st * (uint32_t)(NROWS * CW * sizeof(T));
                    auto load_run = [&](int row, int n, uint32_t dst) {   // warp-uniform branches
#pragma unroll
                        for (int i = 0; i < SG_M3_NMAPS; ++i) {
                            const int h = 16 >> i;
                            if (n & h) {
                                sg_tma_load_3d_elect(dst, &maps.m[i], j1_0, row, pl0 + p, fb);
                                row += h;
                                dst += (uint32_t)(h * CW * sizeof(T));
                            }
                        }
                    };
                    if (contiguous) {
                        load_run(rg[0], cnt, dst0);
                    } else {
#pragma unroll
                        for (int gg = 0; gg < G2; ++gg)
                            if (nq[gg] > 0) load_run(rg[gg] + r_off, nq[gg], dst0 + (uint32_t)(gg * RPT * CW * sizeof(T)));
                    }
                    if (++st == NS) { st = 0; ph ^= 1u; first_round = false; }
                }
            }
        }
        return;
    }

    // ======================= consumers: thread = (column, span g of the tile) =======================
    // the segment sequence is advanced by thread 0 on an iterator that lives in shared memory
    __shared__ SgM3SegIter it;
    __shared__ int it_has;
    if (tid == 0) it.init((int)blockIdx.x, (int)gridDim.x, (int64_t)a.nout * a.nb1, a.n3, a.tiles2, cumw);
    const int col = tid & (CW - 1), g = tid / CW;
    const int colskew = col + (col >> 2);
    int st = 0;
    unsigned ph = 0;
    int cur_b1 = -1, blk_par = 1;               // blk_par toggles with every new column block (two can be in flight)
    const uint32_t ring_c = sg_smem_u32(ring) + (uint32_t)(col * sizeof(T));
    const uint32_t b2_u = sg_smem_u32(b2s) + (uint32_t)(g * RPT * 4 * sizeof(T)), b3_u = sg_opaque(sg_smem_u32(b3s));
    const uint32_t full_u = sg_opaque(sg_smem_u32(full)), empty_u = sg_opaque(sg_smem_u32(empty));
    const uint32_t pbar_u = sg_smem_u32(pbar);
    const uint32_t s3_u = sg_opaque(sg_smem_u32(s3s));
    const uint32_t park_u = sg_smem_u32(park), wab_u = sg_smem_u32(wAB);

    // ---- post pipeline: at step e (one per queued control plane) a warp runs
    //   phase 1 of row e   (span partial over dimension 2 -> tq),           arrives pbar[0]
    //   phase 2 of row e-1 (control slots: park = sum of span partials),     arrives pbar[1]
    //   phase 3 of row e-2 (span sums over dimension 1: a4),                 arrives pbar[2]
    //   phase 4 of row e-3 (control indices -> partials in global memory).
    // Every wait is on a barrier the other warps arrived at one step (>= one knot span of planes) earlier, so in steady
    // state nobody blocks; rows use buffer (step % 3), which is provably free again three steps later (warps are at most
    // one step apart because of the waits).
    unsigned step = 0;                          // pipeline steps taken so far by this thread
    auto wait_phase = [&](int k, unsigned row_step) {   // phase k+1 of the row queued at step row_step is complete
        sg_mbar_wait_u(pbar_u + (uint32_t)((k * NQ + (int)(row_step % NQ)) * 8), (row_step / NQ) & 1u);
    };
    auto arrive_phase = [&](int k, unsigned row_step) {
        __syncwarp();
        if (lane == 0) sg_mbar_arrive_u(pbar_u + (uint32_t)((k * NQ + (int)(row_step % NQ)) * 8));
    };
    // t: this thread's span partial of the new row (phase 1 input); meta written by thread 0
    auto pipeline_step = [&](bool new_row, const T (&t)[4], int64_t base_off, int b1, int flags) {
        const unsigned e = step;
        {   // ---- phase 1 (row e)
            const int qb = (int)(e % NQ);
            if (new_row) {
                T *__restrict__ dst = tq + ((size_t)qb * G2 * 4 + g * 4) * CW + col;
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[k * CW] = t[k];
            }
            if (tid == 0) {
                qmeta[e & 7].base_off = base_off;
                qmeta[e & 7].b1 = b1;
                qmeta[e & 7].flags = new_row ? (flags | 4) : 0;
            }
            arrive_phase(0, e);
        }
        if (e >= 1) {   // ---- phase 2 (row e-1): park[slot][col] = sum_{g'+k = slot} tq[g'][k][col], slots g and g+G2
            const unsigned r = e - 1;
            const int qb = (int)(r % NQ);
            wait_phase(0, r);
            if (qmeta[r & 7].flags & 4) {
                const T *__restrict__ tqs = tq + (size_t)qb * G2 * 4 * CW + col;
                T *__restrict__ pk = park + (size_t)qb * S2 * PITCH + colskew;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int slot = g + h * G2;
                    if (slot < S2) {
                        T v = T(0);
#pragma unroll
                        for (int gg = 0; gg < G2; ++gg) {
                            const int k = slot - gg;
                            if (k >= 0 && k <= P) v += tqs[(gg * 4 + k) * CW];
                        }
                        pk[slot * PITCH] = v;
                    }
                }
            }
            arrive_phase(1, r);
        }
        if (e >= 2) {   // ---- phase 3 (row e-2): A[sl][slot][e'] = sum_{j in span sl} B1[j,e'] * park[slot][j]
            const unsigned r = e - 2;
            const int qb = (int)(r % NQ);
            wait_phase(1, r);
            const SgM3RowMeta m = qmeta[r & 7];
            if (m.flags & 4) {
                const SgM3Block &B = blk[(m.flags >> 3) & 1];
                if (B.fast && tid < B.nsp_b * S2) {
                    const int sl = tid % B.nsp_b, slot = tid / B.nsp_b;
                    const int64_t j1_0 = B.j1_0;
                    const int c0 = (int)(max((int64_t)a.start1[B.lo_b + a.P1 + sl], j1_0) - j1_0);
                    const int c1 = (int)(min((int64_t)a.start1[B.lo_b + a.P1 + sl + 1], j1_0 + B.ncols) - j1_0);
                    T A[4] = {T(0), T(0), T(0), T(0)};
                    const uint32_t xb = park_u + (uint32_t)(((size_t)qb * S2 + slot) * PITCH * sizeof(T));
                    const uint32_t wb = wab_u + (uint32_t)(((m.flags >> 3) & 1) * 4 * PITCH * sizeof(T));
                    for (int j = c0; j < c1; ++j) {
                        const uint32_t jsk = (uint32_t)(j + (j >> 2));
                        T w[4];
                        sg_lds2(wb + jsk * (uint32_t)(2 * sizeof(T)), w[0], w[1]);
                        sg_lds2(wb + (PITCH + jsk) * (uint32_t)(2 * sizeof(T)), w[2], w[3]);
                        const T x = sg_lds(xb + jsk * (uint32_t)sizeof(T), T(0));
#pragma unroll
                        for (int ee = 0; ee < 4; ++ee) A[ee] = fma(w[ee], x, A[ee]);
                    }
                    T *__restrict__ ad = a4 + ((size_t)qb * S2 + slot) * 4 * NSPMAX + sl;
#pragma unroll
                    for (int ee = 0; ee < 4; ++ee) ad[ee * NSPMAX] = A[ee];
                } else if (!B.fast) {
                    // sparse sampling / high degree in dimension 1: direct table look-ups on the parked control slots,
                    // written straight to the partials (park is only guaranteed intact during this phase)
                    T *__restrict__ prow = a.part + m.base_off + ((int64_t)(B.lo_b - 1) + (int64_t)m.b1 * (a.P1 + 1));
                    const bool store_ok = m.flags & 1, first_pass = m.flags & 2;
                    const T *__restrict__ pk = park + (size_t)qb * S2 * PITCH;
                    for (int om = tid; om < B.ni1 * S2; om += NCONS) {
                        const int il = om % B.ni1, slot = om / B.ni1;
                        const int64_t i = (int64_t)B.lo_b + il;
                        const int64_t s0 = i > a.P1 + 1 ? i : a.P1 + 1;
                        const int64_t s1 = i + a.P1 < a.c1 ? i + a.P1 : a.c1;
                        const int64_t lo = max((int64_t)a.start1[s0], (int64_t)B.j1_0), hi = min((int64_t)a.start1[s1 + 1], (int64_t)B.j1_0 + B.ncols);
                        T sum = T(0);
                        for (int64_t j = lo; j < hi; ++j) {
                            const int k = (int)(i - sg_ldg(a.index1 + j) + a.P1);
                            const int jj = (int)(j - B.j1_0);
                            sum = fma(sg_ldg(a.table1 + j + a.n1 * k), pk[slot * PITCH + jj + (jj >> 2)], sum);
                        }
                        if (store_ok) { T *dst = prow + il + (int64_t)a.L1 * slot; *dst = first_pass ? sum : *dst + sum; }
                    }
                }
            }
            arrive_phase(2, r);
        }
        if (e >= 3) {   // ---- phase 4 (row e-3): out[il][slot] = sum_e' A[il - e'][slot][e'] -> partials
            const unsigned r = e - 3;
            const int qb = (int)(r % NQ);
            wait_phase(2, r);
            const SgM3RowMeta m = qmeta[r & 7];
            if (m.flags & 4) {
                const SgM3Block &B = blk[(m.flags >> 3) & 1];
                T *__restrict__ prow = a.part + m.base_off + ((int64_t)(B.lo_b - 1) + (int64_t)m.b1 * (a.P1 + 1));
                const bool store_ok = m.flags & 1, first_pass = m.flags & 2;
                if (B.fast) {
                    if (tid < B.ni1 * S2) {
                        const int il = tid % B.ni1, slot = tid / B.ni1;
                        const T *__restrict__ as = a4 + ((size_t)qb * S2 + slot) * 4 * NSPMAX;
                        T sum = T(0);
#pragma unroll
                        for (int ee = 0; ee < 4; ++ee) {
                            const int sl = il - ee;
                            if (ee <= a.P1 && sl >= 0 && sl < B.nsp_b) sum += as[ee * NSPMAX + sl];
                        }
                        if (store_ok) { T *dst = prow + il + (int64_t)a.L1 * slot; *dst = first_pass ? sum : *dst + sum; }
                    }
                }
            }
        }
        ++step;
    };

    while (true) {
        sg_m3_consumer_bar();                   // every consumer is done with the previous segment (descriptor, shared tables)
        if (tid == 0) it_has = it.next() ? 1 : 0;
        sg_m3_consumer_bar();
        if (!it_has) break;
        const int64_t o = it.col_ob / a.nb1;
        const int b1 = (int)(it.col_ob % a.nb1);
        const int tile2 = it.col_t;
        const int kseg = it.kseg;
        const int seg_p0 = it.p0, seg_p1 = it.p1;
        const int64_t colid = it.col_ob * a.tiles2 + tile2;
        const bool store_ok = kseg < a.maxseg;
        if (!store_ok && tid == 0) a.hdr->nonmonotone = 1;   // never with sane sample distributions: the scatter kernel
                                                             // takes over (the ring is still drained in step)
        if (b1 != cur_b1) {
            // ---- new column block: its constants and B1 rows go to the other buffer; rows of the previous block that are
            // still in the post pipeline keep using theirs (a segment queues >= P+1 rows, the pipeline holds 3)
            cur_b1 = b1;
            blk_par ^= 1;
            const int64_t j1_0 = (int64_t)b1 * CW;
            const int ncols = (int)min((int64_t)CW, a.n1 - j1_0);
            const int first = sg_ldg(a.index1 + j1_0), last = sg_ldg(a.index1 + j1_0 + ncols - 1);
            if (tid == 0) {
                SgM3Block &B = blk[blk_par];
                B.lo_b = first - a.P1; B.nsp_b = last - first + 1; B.ni1 = last - first + 1 + a.P1;
                B.fast = (last - first + 1 <= NSPMAX && a.P1 <= 3) ? 1 : 0; B.ncols = ncols; B.j1_0 = j1_0;
            }
            if (tid < ncols) {
                T *__restrict__ wd = wAB + (size_t)blk_par * 4 * PITCH;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    wd[((k >> 1) * PITCH + colskew) * 2 + (k & 1)] = k <= a.P1 ? sg_ldg(a.table1 + j1_0 + tid + a.n1 * k) : T(0);
            }
        }
        // rows of this thread's span g of the tile, and the longest span (block-uniform)
        int rg0 = 0, rg1 = 0, maxnr = 0;
        const int tile_r0 = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2, a.c2 + 1)];
#pragma unroll
        for (int gg = 0; gg < G2; ++gg) {
            const int r_a = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg, a.c2 + 1)];
            const int r_b = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg + 1, a.c2 + 1)];
            if (gg == g) { rg0 = r_a; rg1 = r_b; }
            maxnr = max(maxnr, r_b - r_a);
        }
        const int s3_first = sg_ldg(a.index3 + seg_p0), s3_last = sg_ldg(a.index3 + seg_p1 - 1);
        // partial index = pos1 + L1*(slot + S2*(tile2 + tiles2*(rowpos + L3*o)))
        const int64_t row_stride = (int64_t)a.L1 * S2 * a.tiles2;
        const int64_t col_off = (int64_t)a.L1 * S2 * tile2 + row_stride * ((int64_t)a.L3 * o + (int64_t)kseg * (P + 1));

        for (int r_off = 0; r_off < maxnr; r_off += RPT) {                // passes (one unless a span has > RPT rows)
            sg_m3_consumer_bar();               // the previous pass no longer reads b2s; block constants are visible
            if (tid < NROWS * 4) {              // B2 rows of the stage rows of this pass, zero for absent rows
                const int l = tid >> 2, k = tid & 3;
                const int gg = l / RPT, qq = l % RPT;
                const int r_a = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg, a.c2 + 1)];
                const int r_b = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg + 1, a.c2 + 1)];
                const int r = r_a + r_off + qq;
                b2s[tid] = (r < r_b && k <= P) ? sg_ldg(a.table2 + r + a.n2 * k) : T(0);
            }
            const int nq = max(0, min(RPT, rg1 - rg0 - r_off));          // rows of this thread in this pass (warp-uniform)
            const int flags = (store_ok ? 1 : 0) | (r_off == 0 ? 2 : 0) | (blk_par << 3);

            T acc[RPT][P + 1];
#pragma unroll
            for (int q = 0; q < RPT; ++q)
#pragma unroll
                for (int k = 0; k <= P; ++k) acc[q][k] = T(0);
            int cur = s3_first;

            // One control plane of dimension 3 is complete (acc[.][0]): contract this thread's rows over dimension 2,
            // push the span partial into the post pipeline, slide the window.  No block-wide barrier.
            auto emit_oldest = [&]() {
                T t[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    T w[4];
                    sg_lds4(b2_u + (uint32_t)(q * 4 * sizeof(T)), w);
#pragma unroll
                    for (int k = 0; k <= P; ++k) t[k] = fma(w[k], acc[q][0], t[k]);
                }
                pipeline_step(true, t, col_off + row_stride * (int64_t)(cur - P - 1), b1, flags);
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
#pragma unroll
                    for (int k = 0; k < P; ++k) acc[q][k] = acc[q][k + 1];
                    acc[q][P] = T(0);
                }
                ++cur;
            };

            int cnt = 0;                        // stage rows present in this pass (block-uniform); none -> nothing was streamed
#pragma unroll
            for (int gg = 0; gg < G2; ++gg) {
                const int r_a = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg, a.c2 + 1)];
                const int r_b = a.start2[(int)min((int64_t)P + 1 + (int64_t)G2 * tile2 + gg + 1, a.c2 + 1)];
                cnt += max(0, min(RPT, r_b - r_a - r_off));
            }
            if (cnt == 0) continue;
            // stage row of this thread's first row: contiguous tiles (single pass) keep the plane's row order
            const uint32_t ring_u = sg_opaque(ring_c + (uint32_t)((maxnr <= RPT ? rg0 - tile_r0 : g * RPT) * CW * sizeof(T)));

            for (int pp = seg_p0; pp < seg_p1; pp += MAXPL) {
                const int np = min(MAXPL, seg_p1 - pp);
                sg_m3_consumer_bar();           // previous piece's tables are no longer read
                for (int s = tid; s < np; s += NCONS) {
                    s3s[s] = sg_ldg(a.index3 + pp + s);
#pragma unroll
                    for (int k = 0; k < 4; ++k) b3s[s * 4 + k] = k <= P ? sg_ldg(a.table3 + pp + s + a.n3 * k) : T(0);
                }
                sg_m3_consumer_bar();
                for (int s = 0; s < np; ++s) {
                    int sp;
                    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(sp) : "r"(s3_u + (uint32_t)s * 4u));
#pragma unroll 1
                    while (cur < sp) emit_oldest();
                    T b[4];
                    sg_lds4(b3_u + (uint32_t)s * (uint32_t)(4 * sizeof(T)), b);
                    sg_mbar_wait_u(full_u + (uint32_t)st * 8u, ph);
                    const uint32_t xs_u = ring_u + (uint32_t)st * (uint32_t)(NROWS * CW * sizeof(T));
                    T x[RPT];                   // rows past the span's end (other rows / stale data) are dropped
#pragma unroll
                    for (int q = 0; q < RPT; ++q) x[q] = sg_lds(xs_u + (uint32_t)(q * CW * sizeof(T)), T(0));
#pragma unroll
                    for (int q = 0; q < RPT; ++q) x[q] = q < nq ? x[q] : T(0);
                    __syncwarp();
                    if (lane == 0) sg_mbar_arrive_u(empty_u + (uint32_t)st * 8u);   // this warp no longer needs the stage
                    if (++st == NS) { st = 0; ph ^= 1u; }
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
#pragma unroll
                        for (int k = 0; k <= P; ++k) acc[q][k] = fma(b[k], x[q], acc[q][k]);
                }
            }
            // flush the P+1 live planes (partial: the neighbouring segments add theirs in the combine)
#pragma unroll 1
            for (int k = 0; k <= P; ++k) emit_oldest();
        }
        // control planes this segment holds: bit kseg of the column's row mask (integer OR: order-independent)
        if (store_ok)
            for (int r = s3_first - P - 1 + tid; r < s3_last; r += NCONS) atomicOr(a.rowmask + colid * a.c3 + r, 1u << kseg);
    }
    // drain the post pipeline
    {
        const T t0[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 1
        for (int k = 0; k < 3; ++k) pipeline_step(false, t0, 0, 0, 0);
    }
}

// cp[i1,i2,i3,o] = sum over the column blocks b covering i1, the tiles t covering i2 and the segments of column (b,t)
// that hold control plane i3 (bits of the row mask, ascending), i.e. in a fixed order.  Also the zero fill for the
// scatter fallback.  grid = (ceil(c1/128), c2, c3*nout)
template <typename T>
__global__ void __launch_bounds__(128) sg_adj_combine3_kernel(T *__restrict__ cp, const T *__restrict__ part, const unsigned *__restrict__ rowmask,
                                                              const SgAdjointHeader *hdr, const int32_t *__restrict__ index1, int64_t n1,
                                                              int64_t c1, int64_t c2, int64_t c3, int P1, int P, int G2, int nb1, int tiles2,
                                                              int L1, int L3)
{
    const int64_t i1 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 0-based control indices
    if (i1 >= c1) return;
    const int64_t i2 = blockIdx.y;
    const int64_t i3 = blockIdx.z % c3, o = blockIdx.z / c3;
    T *__restrict__ out = cp + i1 + c1 * (i2 + c2 * (i3 + c3 * o));
    if (hdr->nonmonotone) { *out = T(0); return; }                       // the scatter kernel accumulates into zeros
    const int S2 = G2 + P;
    const int i = (int)i1 + 1;
    // first block whose last span >= i (blocks are SG_M3_CW columns wide)
    int b_lo = 0;
    {
        int lo = 0, hi = nb1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int last = sg_ldg(index1 + min((int64_t)(mid + 1) * SG_M3_CW, n1) - 1);
            if (last >= i) hi = mid; else lo = mid + 1;
        }
        b_lo = lo;
    }
    // tile t holds the control rows [G2*t, G2*t + S2) (0-based)
    const int t_hi = (int)min(i2 / G2, (int64_t)tiles2 - 1);
    const int t_lo = i2 > P ? (int)((i2 - P) / G2) : 0;
    const int64_t row_stride = (int64_t)L1 * S2 * tiles2;
    T acc = T(0);
    for (int b = b_lo; b < nb1; ++b) {
        const int first = sg_ldg(index1 + (int64_t)b * SG_M3_CW);
        if (first - P1 > i) break;
        const int64_t pos1 = i1 + (int64_t)b * (P1 + 1);
        for (int t = t_lo; t <= t_hi; ++t) {
            const int slot = (int)i2 - G2 * t;
            if (slot < 0 || slot >= S2) continue;
            const int64_t colid = (o * nb1 + b) * tiles2 + t;
            unsigned mask = sg_ldg(rowmask + colid * c3 + i3);
            const T *__restrict__ pc = part + pos1 + (int64_t)L1 * (slot + (int64_t)S2 * (t + (int64_t)tiles2 * ((int64_t)L3 * o)));
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                acc += pc[row_stride * (i3 + (int64_t)k * (P + 1))];
            }
        }
    }
    *out = acc;
}
