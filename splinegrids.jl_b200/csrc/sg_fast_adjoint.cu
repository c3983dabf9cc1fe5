// sg_fast_adjoint.cu -- plan + dispatch of the atomics-free adjoint passes (kernels: sg_fast_adjoint.cuh).
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <type_traits>

#include "sg_evaluate_generic.cuh"
#include "sg_fast.cuh"
#include "sg_fast_adjoint.cuh"
#include "sg_adjoint_post2.cuh"
#include "sg_adjoint_march2g.cuh"
#include "sg_fast_eval.cuh"

static int sg_env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// ---------------------------------------------------------------------------------------------
// adjoint: plan of per-dimension passes (shared by the scratch-size query and the launcher)
// ---------------------------------------------------------------------------------------------
struct SgAdjPass {
    int d;                       // 0-based dimension contracted by this pass
    int64_t inner, n_d, c_d, outer;
    int P, G, nchunks;
    size_t part_off, out_off;    // byte offsets into the scratch (part only if nchunks > 1)
};
struct SgAdjPlan {
    int npassA;
    SgAdjPass pass[SG_MAX_DIMS];
    size_t bytes;
};

static size_t sg_al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Number of chunks of the marching axis of a 2-D grid's first pass from the kernel's RESIDENT capacity (CTAs per SM x SMs):
// a grid of 1.26 waves runs as long as one of 2 waves, so the chunk count is chosen for full waves, with a mild
// preference for long chunks (less halo, fewer pipeline fills).  Measured on C4: 23 chunks (368 CTAs, capacity 296) 0.443 ms,
// 37 chunks (592 CTAs = 2 full waves) 0.332 ms.
static int64_t sg_pick_nchunks(int64_t nspans, int64_t ctas_per_chunk, int64_t capacity, int P)
{
    int64_t best = 1;
    double best_score = -1.0;
    for (int64_t nch = 1; nch <= std::min<int64_t>(nspans, 96); ++nch) {
        const int64_t G = (nspans + nch - 1) / nch, ne = (nspans + G - 1) / G;
        if (ne != nch) continue;
        const double ctas = (double)ctas_per_chunk * ne, waves = std::ceil(ctas / capacity);
        const double score = ctas / (waves * capacity) * (double)G / ((double)G + 0.25 * P);
        if (score > best_score) { best_score = score; best = nch; }
    }
    return best;
}
#define SG_ADJ_NCH_MAX 96

// capacity0 > 0 (2-D grids): resident CTAs of the first pass's kernel, ctas_per_chunk0 its CTAs per chunk;
// worst = true (workspace query): size the first pass's partials for ANY chunk count the run may pick.
static SgAdjPlan sg_adjoint_pass_plan(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout, const int *degree,
                                      int elem_size, int64_t capacity0 = 0, int64_t ctas_per_chunk0 = 0, bool worst = false)
{
    SgAdjPlan pl{};
    const int V = elem_size == 4 ? 4 : 2;
    size_t off = 0;
    int64_t outer = nout;
    for (int d = nin - 1; d >= 1; --d) {
        SgAdjPass &ps = pl.pass[pl.npassA++];
        ps.d = d;
        ps.inner = 1;
        for (int e = 0; e < d; ++e) ps.inner *= n_samples[e];
        ps.n_d = n_samples[d];
        ps.c_d = n_cp[d];
        ps.outer = outer;
        ps.P = degree[d];
        const int64_t nspans = ps.c_d - ps.P;
        // columns that will really be active: a slab of a sharded grid touches at most n_D + p_D of the c_D control
        // planes of the slowest axis (the kernels skip the others on device)
        int64_t outer_est = outer;
        if (d < nin - 1) {
            const int64_t cD = n_cp[nin - 1], actD = std::min<int64_t>(cD, n_samples[nin - 1] + degree[nin - 1]);
            outer_est = outer / cD * actD;
        }
        const int64_t threads = ((ps.inner + V - 1) / V) * outer_est;
        int64_t nchunks = 1;
        if (threads < 24576) nchunks = std::min<int64_t>(nspans, (148 * 8 * 128 + threads - 1) / threads);
        if (nin == 2 && capacity0 > 0 && ctas_per_chunk0 > 0 && sg_env_int("SG_ADJ_WAVES", 1))
            nchunks = sg_pick_nchunks(nspans, ctas_per_chunk0, capacity0, ps.P);
        const int forced = sg_env_int("SG_ADJ_CHUNKS", 0);
        if (forced > 0) nchunks = std::min<int64_t>(nspans, forced);
        ps.G = (int)((nspans + nchunks - 1) / nchunks);
        ps.nchunks = (int)((nspans + ps.G - 1) / ps.G);
        ps.part_off = off;
        size_t part_elems = ps.nchunks > 1 ? (size_t)ps.inner * (ps.G + ps.P) * ps.nchunks * outer : 0;
        if (worst && nin == 2) {                                        // any chunk count up to SG_ADJ_NCH_MAX
            for (int64_t nch = 2; nch <= std::min<int64_t>(nspans, std::max<int64_t>(SG_ADJ_NCH_MAX, forced)); ++nch) {
                const int64_t G = (nspans + nch - 1) / nch, ne = (nspans + G - 1) / G;
                part_elems = std::max(part_elems, (size_t)ps.inner * (G + ps.P) * ne * outer);
            }
        }
        if (part_elems > 0) off += sg_al256(part_elems * elem_size);
        ps.out_off = off;
        off += sg_al256((size_t)ps.inner * ps.c_d * outer * elem_size);
        outer *= ps.c_d;
    }
    pl.bytes = off;
    return pl;
}

static bool sg_adjoint_fast_supported(int nin, const int *degree, bool rational)
{
    if (nin < 1 || nin > 4) return false;
    for (int d = 0; d < nin; ++d)
        if (degree[d] < 1 || degree[d] > 5) return false;
    if (rational) {
        if (nin != 2 || degree[0] != degree[1] || degree[0] > 3) return false;
    }
    return true;
}

template <typename T, int P, int NT, bool RAT2D>
static void sg_launch_adj_march_nt(const SgAdjPassArgs<T> &pa, int64_t outer, cudaStream_t st)
{
    constexpr int VV = sizeof(T) == 4 ? 4 : 2;
    const bool vec_ok = (pa.inner % VV == 0) && (reinterpret_cast<uintptr_t>(pa.X) % 16 == 0) &&
                        (reinterpret_cast<uintptr_t>(pa.Y) % 16 == 0);
    if constexpr (!RAT2D) {
        if (pa.bt_hdr != nullptr) {                                    // fused 2-D march (the caller checked vec_ok)
            dim3 grid((unsigned)pa.nb1, (unsigned)pa.nchunks, (unsigned)(outer / NT));
            sg_adj_march_kernel<T, P, VV, NT, false, true><<<grid, 128, 0, st>>>(pa);
            g_sg_launches.fetch_add(1);
            return;
        }
    }
    if (vec_ok) {
        dim3 grid((unsigned)((pa.inner + 128 * VV - 1) / (128 * VV)), (unsigned)pa.nchunks, (unsigned)(outer / NT));
        sg_adj_march_kernel<T, P, VV, NT, RAT2D><<<grid, 128, 0, st>>>(pa);
    } else {
        dim3 grid((unsigned)((pa.inner + 127) / 128), (unsigned)pa.nchunks, (unsigned)(outer / NT));
        sg_adj_march_kernel<T, P, 1, NT, RAT2D><<<grid, 128, 0, st>>>(pa);
    }
    g_sg_launches.fetch_add(1);
}

// nt = channels per thread (must divide outer)
template <typename T, int P, bool RAT2D>
static void sg_launch_adj_march(const SgAdjPassArgs<T> &pa, int64_t outer, int nt, cudaStream_t st)
{
    if constexpr (!RAT2D) {
        sg_launch_adj_march_nt<T, P, 1, false>(pa, outer, st);
    } else {
    switch (nt) {
        case 4: sg_launch_adj_march_nt<T, P, 4, RAT2D>(pa, outer, st); break;
        case 3: sg_launch_adj_march_nt<T, P, 3, RAT2D>(pa, outer, st); break;
        case 2: sg_launch_adj_march_nt<T, P, 2, RAT2D>(pa, outer, st); break;
        default: sg_launch_adj_march_nt<T, P, 1, RAT2D>(pa, outer, st); break;
    }
    }
}

// Resident CTAs (all SMs) of the pass-A kernel instantiation a 2-D grid will run: occupancy API, cached per instantiation.
template <typename T, int P, int NT, bool RAT, bool F1>
static int64_t sg_adj_march_capacity_inst()
{
    static const int64_t cap = [] {
        constexpr int VV = sizeof(T) == 4 ? 4 : 2;
        int per_sm = 0, dev = 0, sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sg_adj_march_kernel<T, P, VV, NT, RAT, F1>, 128, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            return (int64_t)0;
        }
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (int64_t)per_sm * sms;
    }();
    return cap;
}
template <typename T>
static int64_t sg_adj_march_capacity(int P, int nt, bool rational, bool fused)
{
    if (rational) {
        if (P < 1 || P > 3) return 0;
#define SG_CAP_RAT(PP)                                                               \
    switch (nt) {                                                                    \
        case 4: return sg_adj_march_capacity_inst<T, PP, 4, true, false>();          \
        case 3: return sg_adj_march_capacity_inst<T, PP, 3, true, false>();          \
        case 2: return sg_adj_march_capacity_inst<T, PP, 2, true, false>();          \
        default: return sg_adj_march_capacity_inst<T, PP, 1, true, false>();         \
    }
        if (P == 1) { SG_CAP_RAT(1) } else if (P == 2) { SG_CAP_RAT(2) } else { SG_CAP_RAT(3) }
#undef SG_CAP_RAT
    }
    switch (P) {
        case 1: return fused ? sg_adj_march_capacity_inst<T, 1, 1, false, true>() : sg_adj_march_capacity_inst<T, 1, 1, false, false>();
        case 2: return fused ? sg_adj_march_capacity_inst<T, 2, 1, false, true>() : sg_adj_march_capacity_inst<T, 2, 1, false, false>();
        case 3: return fused ? sg_adj_march_capacity_inst<T, 3, 1, false, true>() : sg_adj_march_capacity_inst<T, 3, 1, false, false>();
        case 4: return fused ? sg_adj_march_capacity_inst<T, 4, 1, false, true>() : sg_adj_march_capacity_inst<T, 4, 1, false, false>();
        case 5: return fused ? sg_adj_march_capacity_inst<T, 5, 1, false, true>() : sg_adj_march_capacity_inst<T, 5, 1, false, false>();
        default: return 0;
    }
}

// ---- one generic pass A (+ chunk combine) ---------------------------------------------------------
template <typename T>
static void sg_run_pass_a(const SgAdjPass &ps, SgAdjPassArgs<T> &pa, T *out, T *part, bool rat_here, int path,
                          const SgAdjointHeader *hdr, cudaStream_t st)
{
    const int nt = (rat_here && ps.outer <= 4) ? (int)ps.outer : 1;   // shares the denominators between the outputs
    if (rat_here) {
        switch (ps.P) {
            case 1: sg_launch_adj_march<T, 1, true>(pa, ps.outer, nt, st); break;
            case 2: sg_launch_adj_march<T, 2, true>(pa, ps.outer, nt, st); break;
            default: sg_launch_adj_march<T, 3, true>(pa, ps.outer, nt, st); break;
        }
    } else {
        switch (ps.P) {
            case 1: sg_launch_adj_march<T, 1, false>(pa, ps.outer, nt, st); break;
            case 2: sg_launch_adj_march<T, 2, false>(pa, ps.outer, nt, st); break;
            case 3: sg_launch_adj_march<T, 3, false>(pa, ps.outer, nt, st); break;
            case 4: sg_launch_adj_march<T, 4, false>(pa, ps.outer, nt, st); break;
            default: sg_launch_adj_march<T, 5, false>(pa, ps.outer, nt, st); break;
        }
    }
    if (ps.nchunks > 1 && pa.bt_hdr == nullptr) {
        constexpr int VC = sizeof(T) == 4 ? 4 : 2;
        const bool vec_ok = (ps.inner % VC == 0);
        dim3 cgrid((unsigned)((ps.inner + 128 * VC - 1) / (128 * VC)), (unsigned)((ps.c_d + SG_COMBINE_ROWS - 1) / SG_COMBINE_ROWS), (unsigned)ps.outer);
        sg_adj_combine_kernel<T, VC><<<cgrid, 128, 0, st>>>(out, part, hdr, ps.inner, ps.c_d, ps.G, ps.nchunks, ps.P, vec_ok, path,
                                                            pa.dim, pa.restrict_spans, pa.last_dim, pa.last_P, pa.last_div, pa.last_c);
        g_sg_launches.fetch_add(1);
    }
}

// ---- multi-pass pipeline ---------------------------------------------------------------------------
template <typename T>
static int sg_run_multipass(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr, const T *eval,
                            const T *weights, char *ws, int path, const SgAdjKnown &known, const char **variant, cudaStream_t st)
{
    const bool rational = weights != nullptr;
    int64_t capacity0 = 0, cpc0 = 0;
    if (a.nin == 2) {                                                  // full waves for the pass that reads the sample array
        constexpr int VQ = sizeof(T) == 4 ? 4 : 2;
        const bool fused = known.planned && known.fused_ok && ss.bt_hdr != nullptr && !rational;
        const int nt = (rational && a.nout <= 4) ? a.nout : 1;
        capacity0 = sg_adj_march_capacity<T>(a.degree[1], nt, rational, fused);
        cpc0 = ((a.n_samples[0] + 128 * VQ - 1) / (128 * VQ)) * (a.nout / nt);
    }
    const SgAdjPlan pl = sg_adjoint_pass_plan(a.nin, a.n_samples, a.n_cp, a.nout, a.degree, (int)sizeof(T), capacity0, cpc0);
    {
        // fused 2-D march (planned calls): dimension 2 marched per column, dimension 1 contracted in the kernel's epilogue
        constexpr int VV = sizeof(T) == 4 ? 4 : 2;
        const SgAdjPass &ps = pl.pass[0];
        T *part = reinterpret_cast<T *>(ws + (ps.nchunks > 1 ? ps.part_off : ps.out_off));
        const bool nt_ok = !rational || a.nout <= 4;                  // rational: all channels of a thread share the denominators
        // (rational grids keep the multi-pass pipeline: their fused kernel needs 242 registers and measured slower)
        if (a.nin == 2 && !rational && known.planned && known.fused_ok && ss.bt_hdr != nullptr && ss.rfast == 1 && ss.bw == 128 * VV && nt_ok &&
            a.n_samples[0] % VV == 0 && reinterpret_cast<uintptr_t>(eval) % 16 == 0 && ps.outer <= 65535 && ps.nchunks <= 65535) {
            SgAdjPassArgs<T> pa{};
            pa.X = eval; pa.Y = part; pa.table = a.table[1]; pa.index = a.index[1]; pa.span_start = ss.start[1];
            pa.hdr = hdr; pa.inner = ps.inner; pa.n_d = ps.n_d; pa.c_d = ps.c_d; pa.G = ps.G; pa.nchunks = ps.nchunks;
            pa.path = path; pa.dim = 1; pa.last_dim = -1; pa.restrict_spans = 1;
            pa.bt_hdr = ss.bt_hdr; pa.bt_lol = ss.bt_lol; pa.bt_w = ss.bt_w; pa.icap = ss.icap; pa.rmcap = ss.rmcap; pa.nb1 = ss.nb1;
            if (rational) { pa.weights = weights; pa.table1 = a.table[0]; pa.index1 = a.index[0]; pa.c1 = a.n_cp[0]; }
            sg_run_pass_a<T>(ps, pa, part, part, rational, path, hdr, st);
            dim3 cgrid(sg_blocks(a.n_cp[0], 128), (unsigned)a.n_cp[1], (unsigned)a.nout);
            if (rational)
                sg_adj_combine_f2d_kernel<T, true><<<cgrid, 128, 0, st>>>(cp, part, ss.g_lo, ss.bt_hdr, hdr, weights, a.n_cp[0], a.n_cp[1], ps.G,
                                                                          ps.nchunks, ps.P, ss.icap, ss.nb1, ss.bw);
            else
                sg_adj_combine_f2d_kernel<T, false><<<cgrid, 128, 0, st>>>(cp, part, ss.g_lo, ss.bt_hdr, hdr, weights, a.n_cp[0], a.n_cp[1], ps.G,
                                                                           ps.nchunks, ps.P, ss.icap, ss.nb1, ss.bw);
            g_sg_launches.fetch_add(1);
            *variant = rational ? "adjoint_fused2d_rational" : "adjoint_fused2d";
            return SG_OK;
        }
    }
    for (int k = 0; k < pl.npassA; ++k) {
        const SgAdjPass &ps = pl.pass[k];
        if (ps.outer > 65535 || ps.nchunks > 65535 || (ps.nchunks > 1 && ps.c_d > 65535)) return SG_ERR_UNSUPPORTED;
    }
    const T *X = eval;
    for (int k = 0; k < pl.npassA; ++k) {
        const SgAdjPass &ps = pl.pass[k];
        T *out = reinterpret_cast<T *>(ws + ps.out_off);
        T *part = ps.nchunks > 1 ? reinterpret_cast<T *>(ws + ps.part_off) : out;
        SgAdjPassArgs<T> pa{};
        pa.X = X; pa.Y = part; pa.table = a.table[ps.d]; pa.index = a.index[ps.d]; pa.span_start = ss.start[ps.d];
        pa.hdr = hdr; pa.inner = ps.inner; pa.n_d = ps.n_d; pa.c_d = ps.c_d; pa.G = ps.G; pa.nchunks = ps.nchunks;
        pa.path = path;
        pa.dim = ps.d;
        pa.last_dim = -1;
        pa.restrict_spans = (k == 0 && a.nin >= 2) ? 1 : 0;   // only the slowest axis may be a slab of a sharded grid
        if (k > 0) {   // outer = c_{d+1} * ... * c_D * nout: the last dimension's control index is (r / last_div) % c_D
            pa.last_dim = a.nin - 1; pa.last_P = a.degree[a.nin - 1]; pa.last_c = a.n_cp[a.nin - 1];
            pa.last_div = 1;
            for (int e = ps.d + 1; e < a.nin - 1; ++e) pa.last_div *= a.n_cp[e];
        }
        const bool rat_here = rational && k == 0;   // nin == 2: the first pass marches dim 2 over columns of dim 1
        if (rat_here) { pa.weights = weights; pa.table1 = a.table[0]; pa.index1 = a.index[0]; pa.c1 = a.n_cp[0]; }
        sg_run_pass_a<T>(ps, pa, out, part, rat_here, path, hdr, st);
        X = out;
    }
    // pass B: first dimension
    const int64_t outerB = a.cp_total / a.n_cp[0] * a.nout;
    const int64_t avg_range = (int64_t)(a.degree[0] + 1) * a.n_samples[0] / a.n_cp[0];
    int b_last_dim = -1, b_last_P = 0;
    int64_t b_last_div = 1, b_last_c = 1;
    if (a.nin >= 2) {
        b_last_dim = a.nin - 1; b_last_P = a.degree[a.nin - 1]; b_last_c = a.n_cp[a.nin - 1];
        for (int e = 1; e < a.nin - 1; ++e) b_last_div *= a.n_cp[e];
    }
    auto launch_b = [&](auto lanes_tag, auto rat_tag) {
        constexpr int L = decltype(lanes_tag)::value;
        constexpr bool R = decltype(rat_tag)::value;
        const unsigned gy = (unsigned)std::min<int64_t>(outerB, 32768);
        dim3 bgrid(sg_blocks(a.n_cp[0], 256 / L), gy, (unsigned)((outerB + gy - 1) / gy));
        sg_adj_first_dim_kernel<T, L, R><<<bgrid, 256, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                a.n_cp[0], outerB, a.degree[0], weights, a.cp_total, path,
                                                                b_last_dim, b_last_P, b_last_div, b_last_c);
    };
    if (avg_range >= 64) {
        if (rational) launch_b(std::integral_constant<int, 32>{}, std::true_type{});
        else launch_b(std::integral_constant<int, 32>{}, std::false_type{});
    } else if (avg_range <= 18 && sg_env_int("SG_ADJ_PASSB_ROWS", 1)) {
        // short ranges: per-thread weights in registers, many rows per block
        const int rpb = sg_env_int("SG_ADJ_RPB", 8);
        dim3 rgrid(sg_blocks(a.n_cp[0], 128), (unsigned)((outerB + rpb - 1) / rpb));
        if (rgrid.y > 65535) return SG_ERR_UNSUPPORTED;
        if (rational)
            sg_adj_first_dim_rows_kernel<T, 20, true><<<rgrid, 128, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                            a.n_cp[0], outerB, a.degree[0], rpb, weights, a.cp_total, path,
                                                                            b_last_dim, b_last_P, b_last_div, b_last_c);
        else
            sg_adj_first_dim_rows_kernel<T, 20, false><<<rgrid, 128, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                             a.n_cp[0], outerB, a.degree[0], rpb, weights, a.cp_total, path,
                                                                             b_last_dim, b_last_P, b_last_div, b_last_c);
    } else {
        if (rational) launch_b(std::integral_constant<int, 8>{}, std::true_type{});
        else launch_b(std::integral_constant<int, 8>{}, std::false_type{});
    }
    g_sg_launches.fetch_add(1);
    return SG_OK;
}

// ---- 3-D double-march pipeline ---------------------------------------------------------------------------
struct SgMarch2Plan {
    bool ok;
    int G2, tiles2, G3, chunks3;
    size_t part_off, r_off, bytes;
    SgM2gDims g;                 // fused variant (dimension 1 contracted in the march kernel's epilogue)
    size_t bytes_unfused;
};
#define SG_M2_G2 4
#define SG_M2_RS 6

static SgMarch2Plan sg_adjoint_march2_plan(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout, const int *degree,
                                           int elem_size, bool rational)
{
    SgMarch2Plan mp{};
    mp.ok = false;
    // SG_ADJ_MARCH2: 0 = never, 1 = always when eligible, unset = when the slowest axis is densely sampled (a thin slab
    // of a sharded grid touches few control planes: the chunked multi-pass pipeline is better there)
    const int mode = sg_env_int("SG_ADJ_MARCH2", -1);
    if (rational || nin != 3 || mode == 0) return mp;
    // the slab of a sharded grid has few sample planes but the same density as the full grid (its samples sit in a few
    // consecutive spans; chunks without samples exit at once), so only very thin inputs go to the multi-pass pipeline
    if (mode < 0 && n_samples[2] < 16) return mp;
    const int P = degree[1];
    if (degree[2] != P || P < 1 || P > 3) return mp;
    if (n_samples[0] < 128) return mp;
    mp.G2 = SG_M2_G2;
    const int64_t nsp2 = n_cp[1] - P, nsp3 = n_cp[2] - P;
    mp.tiles2 = (int)((nsp2 + mp.G2 - 1) / mp.G2);
    // The number of chunks of dimension 3 sets the CTA count.  The march kernel keeps 3 CTAs per SM resident (launch bounds
    // and shared memory), and what pays is ONE full wave of long CTAs: fewer pipeline fills, fewer chunk halos in the
    // partials (C3: 3 chunks = 384 CTAs 0.227 ms, 10 chunks = 1280 CTAs 0.232 ms, 4 chunks = 512 CTAs = 1.15 waves 0.283 ms;
    // 8-way slab 0.045 vs 0.050 ms).  The chunks adapt on device to the spans that hold samples (sg_m2_chunk_len), G3 is their
    // worst-case length = the row stride of the partials.
    const int64_t base = ((n_samples[0] + 127) / 128) * mp.tiles2 * nout;
    static const int64_t capacity = [] {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (int64_t)3 * sms;
    }();
    int64_t chunks = 1;
    {
        const int target = sg_env_int("SG_ADJ_M2_CTAS", 0);
        if (target > 0) {
            chunks = std::max<int64_t>(1, (target + base / 2) / base);
        } else {
            double best_score = -1.0;
            for (int64_t nch = 1; nch <= std::max<int64_t>(1, nsp3 / std::max(2, P)); ++nch) {
                const int64_t G = (nsp3 + nch - 1) / nch;
                if ((nsp3 + G - 1) / G != nch) continue;
                const double ctas = (double)base * nch, waves = std::ceil(ctas / (double)capacity);
                const double score = ctas / (waves * capacity) * (double)G / ((double)G + P);
                if (score > best_score) { best_score = score; chunks = nch; }
            }
        }
    }
    const int forced = sg_env_int("SG_ADJ_M2_G3", 0);
    if (forced > 0) chunks = (nsp3 + forced - 1) / forced;
    chunks = std::min<int64_t>(chunks, std::max<int64_t>(1, nsp3 / std::max(2, P)));
    const int64_t G3 = std::max<int64_t>(std::max(2, P), (nsp3 + chunks - 1) / chunks);   // G3 >= P: only neighbouring chunks overlap
    mp.G3 = (int)G3;
    mp.chunks3 = (int)chunks;
    if ((int64_t)mp.chunks3 * nout > 65535 || mp.tiles2 > 65535 || n_cp[1] > 65535 || n_cp[2] * nout > 65535) return mp;
    size_t off = 0;
    mp.part_off = off;
    off += sg_al256((size_t)n_samples[0] * (mp.G2 + P) * mp.tiles2 * (mp.G3 + P) * mp.chunks3 * nout * elem_size);
    mp.r_off = off;                                                     // (no intermediate array any more: the post kernel reads the partials)
    mp.bytes = mp.bytes_unfused = off;
    mp.g = sg_m2g_dims(nin, n_samples, n_cp, degree, rational, elem_size);
    if (mp.g.ok) {
        const double elems = (double)mp.g.icap * (mp.G2 + P) * mp.tiles2 * (mp.G3 + P) * mp.chunks3 * mp.g.nb1 * nout;
        if (elems < 4.0e9) mp.bytes = std::max(mp.bytes, sg_al256((size_t)elems * elem_size));
        else mp.g.ok = false;
    }
    mp.ok = true;
    return mp;
}

// Fused double march: eligibility and table sizes from the shape alone (the prep kernel checks the data on device).
// Worth it when dimension 1 has at least ~2 samples per knot span (a block of 128 columns then leaves <= ~70 control
// indices); icap / rmcap leave a factor 2 of slack over equispaced samples.
SgM2gDims sg_m2g_dims(int nin, const int64_t *n_samples, const int64_t *n_cp, const int *degree, bool rational, int elem_size)
{
    SgM2gDims g{};
    g.ok = false;
    const int64_t n1 = n_samples[0], nsp1 = n_cp[0] - degree[0];
    if (nin == 2) {
        // fused 2-D march (sg_adj_march_kernel<.., F1 = true>): a column block is the 128 * V columns of one CTA
        if (sg_env_int("SG_ADJ_F2D", 1) == 0 || !sg_adjoint_fast_supported(nin, degree, rational)) return g;
        const int V = elem_size == 4 ? 4 : 2;
        g.bw = 128 * V;
        g.rfast = 1;
        if (n1 % V != 0 || n1 < g.bw || n1 < 2 * nsp1) return g;
        // the epilogue runs once per knot span of dimension 2 and costs ~ one warp pass per control index of the block:
        // worth it when a block touches few control indices compared with the rows between two emissions
        // (C2: 6 indices / 67 rows -> 0.078 -> 0.056 ms; C5-scaled: 34 / 17 -> slower than the multi-pass pipeline)
        const double est_ni = (double)g.bw * nsp1 / n1 + degree[0];
        const double rows_per_span2 = (double)n_samples[1] / (double)std::max<int64_t>(1, n_cp[1] - degree[1]);
        if (sg_env_int("SG_ADJ_F2D", 1) != 2 && est_ni > 0.5 * rows_per_span2) return g;
    } else if (nin == 3) {
        // opt-in (SG_ADJ_M2G=1): measured slower than the unfused pipeline, see DESIGN.md
        if (rational || sg_env_int("SG_ADJ_M2G", 0) == 0) return g;
        const int P = degree[1];
        if (degree[2] != P || P < 1 || P > 3 || degree[0] < 1 || degree[0] > 5) return g;
        g.bw = 128;
        g.rfast = 0;
        if (n1 < 128 || n1 < 2 * nsp1) return g;
    } else {
        return g;
    }
    const int64_t est_spans = (g.bw * nsp1 + n1 - 1) / n1;
    g.icap = (int)std::min<int64_t>(((2 * est_spans + degree[0] + 2 + 7) / 8) * 8, 136);
    g.rmcap = (int)std::min<int64_t>((degree[0] + 1) * ((2 * n1 + nsp1 - 1) / nsp1) + 2, g.bw);
    g.nb1 = (int)((n1 + g.bw - 1) / g.bw);
    if ((int64_t)g.nb1 * g.rmcap * g.icap > (int64_t)1 << 26) return g;
    g.ok = true;
    return g;
}

size_t sg_adjoint_fast_scratch_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout, const int *degree,
                                     int elem_size)
{
    if (!sg_adjoint_fast_supported(nin, degree, false)) return 0;
    const size_t a = sg_adjoint_pass_plan(nin, n_samples, n_cp, nout, degree, elem_size, 0, 0, true).bytes;
    const SgMarch2Plan mp = sg_adjoint_march2_plan(nin, n_samples, n_cp, nout, degree, elem_size, false);
    // the pipelines never run together: they share the scratch
    return std::max(a, mp.ok ? mp.bytes : (size_t)0);
}

// benchmark hook (include/splinegrids_b200.h): CUDA events around the dominant kernel
static int g_sg_prof_on = 0, g_sg_prof_recorded = 0;
static cudaEvent_t g_sg_prof_ev[2] = {nullptr, nullptr};
extern "C" void sg_profile_adjoint_main(int enable)
{
    g_sg_prof_on = enable != 0;
    g_sg_prof_recorded = 0;
    if (g_sg_prof_on && !g_sg_prof_ev[0]) {
        if (cudaEventCreate(&g_sg_prof_ev[0]) != cudaSuccess || cudaEventCreate(&g_sg_prof_ev[1]) != cudaSuccess) {
            cudaGetLastError();
            g_sg_prof_on = 0;
        }
    }
}
extern "C" float sg_profile_adjoint_main_ms(void)
{
    if (!g_sg_prof_recorded) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(g_sg_prof_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, g_sg_prof_ev[0], g_sg_prof_ev[1]) != cudaSuccess) {
        cudaGetLastError();
        return -1.0f;
    }
    return ms;
}

#define SG_M2_RTMAX 20
#define SG_M2_NS 3

// eval as a 3-D tensor (n1, n2, n3*nout) with boxes of 128 columns x (1..SG_M2_FAST_ROWS) rows x 1 plane
typedef CUresult (*SgEncodeTiledFnA)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename T>
static bool sg_m2_make_eval_maps(SgM2Maps &maps, const T *eval, int64_t n1, int64_t n2, int64_t n3nout)
{
    static SgEncodeTiledFnA enc = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<SgEncodeTiledFnA>(p);
    }();
    if (!enc || (n1 * sizeof(T)) % 16 != 0 || reinterpret_cast<uintptr_t>(eval) % 16 != 0) return false;
    cuuint64_t dims[3] = {(cuuint64_t)n1, (cuuint64_t)n2, (cuuint64_t)n3nout};
    cuuint64_t strides[2] = {(cuuint64_t)n1 * sizeof(T), (cuuint64_t)n1 * n2 * sizeof(T)};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    for (int r = 0; r < SG_M2_FAST_ROWS; ++r) {
        cuuint32_t box[3] = {128, (cuuint32_t)(r + 1), 1};
        if (enc(&maps.m[r], dt, 3, const_cast<T *>(eval), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}
template <typename T, int P>
static int sg_launch_march2(const SgAdj2Args<T> &m, int nout, const SgAdjKnown &known, cudaStream_t st)
{
    dim3 grid((unsigned)((m.n1 + 127) / 128), (unsigned)m.tiles2, (unsigned)(m.chunks3 * nout));
    // TMA-fed ring: bulk copies need 16-byte aligned rows; rows per tile are data dependent, so the expected count must
    // fit the ring (tiles that do not are skipped by the TMA kernel and done by the register kernel right after)
    const bool tma_ok = sg_env_int("SG_ADJ_M2_TMA", 1) && (m.n1 * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(m.X) % 16 == 0 &&
                        (double)m.n2 / (double)std::max<int64_t>(1, m.c2 - P) * SG_M2_G2 <= SG_M2_RTMAX - 2 &&
                        (double)m.n1 * (SG_M2_G2 + P) * m.tiles2 * (m.G3 + P) * m.chunks3 * nout < 4.0e9;   // 32-bit partial offsets
    if (tma_ok) {
        auto kern = sg_adj_march2_tma_kernel<T, P, SG_M2_G2, SG_M2_RTMAX, SG_M2_NS>;
        const size_t smem = sizeof(T) * SG_M2_NS * SG_M2_RTMAX * 128 + 128;    // + slack for the 128-byte alignment of the ring
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g_sg_prof_on) cudaEventRecord(g_sg_prof_ev[0], st);
        SgM2Maps maps{};
        const int use_maps = sg_env_int("SG_ADJ_M2_MAPS", 1) && sg_m2_make_eval_maps<T>(maps, m.X, m.n1, m.n2, m.n3 * nout) ? 1 : 0;
        kern<<<grid, 160, smem, st>>>(m, maps, use_maps);                   // 4 consumer warps + 1 producer warp
        if (g_sg_prof_on) { cudaEventRecord(g_sg_prof_ev[1], st); g_sg_prof_recorded = 1; }
        // tiles with more rows than the ring holds (normally none: one idle launch; a plan knows and skips it)
        if (known.planned && known.rows2_max <= SG_M2_FAST_ROWS) {
            g_sg_launches.fetch_add(1);
            return SG_OK;
        }
        sg_adj_march2_complement_kernel<T, P, SG_M2_G2, SG_M2_RS><<<148 * 4, 128, 0, st>>>(m, SG_M2_RTMAX, grid.x, grid.y, grid.z);
        g_sg_launches.fetch_add(2);
        return SG_OK;
    }
    sg_adj_march2_kernel<T, P, SG_M2_G2, SG_M2_RS><<<grid, 128, 0, st>>>(m, 0);
    g_sg_launches.fetch_add(1);
    return SG_OK;
}

// ---- fused double march (sg_adjoint_march2g.cuh): planned calls only -- the plan has checked on the host that the
// spans are monotone, every span of dimension 2 fits the ring's row slots and the column-block tables fit ------------
size_t sg_m2_uni_bytes(int elem_size) { return elem_size == 4 ? sizeof(SgM2Uni<float>) : sizeof(SgM2Uni<double>); }

template <typename T>
static bool sg_m2_uni_fill_t(SgM2Uni<T> *u, const T *table2, int64_t n2, int P, const int32_t *start2, int64_t c2,
                             const int32_t *start3, int64_t c3, int sf3, int sl3)
{
    if (n2 * (P + 1) > (int64_t)(SG_M2U_B2_BYTES / sizeof(T)) || c2 + 2 > SG_M2U_STARTS || c3 + 2 > SG_M2U_STARTS) return false;
    for (int64_t r = 0; r < n2; ++r)
        for (int k = 0; k <= P; ++k) u->b2[r * (P + 1) + k] = table2[r + n2 * k];
    for (int64_t s = 0; s <= c2 + 1; ++s) u->start2[s] = start2[s];
    for (int64_t s = 0; s <= c3 + 1; ++s) u->start3[s] = start3[s];
    u->span_first3 = sf3; u->span_last3 = sl3;
    return true;
}
bool sg_m2_uni_fill(void *uni, int elem_size, const void *table2_host, int64_t n2, int P, const int32_t *start2_host, int64_t c2,
                    const int32_t *start3_host, int64_t c3, int span_first3, int span_last3)
{
    if (elem_size == 4)
        return sg_m2_uni_fill_t<float>(static_cast<SgM2Uni<float> *>(uni), static_cast<const float *>(table2_host), n2, P, start2_host, c2,
                                       start3_host, c3, span_first3, span_last3);
    return sg_m2_uni_fill_t<double>(static_cast<SgM2Uni<double> *>(uni), static_cast<const double *>(table2_host), n2, P, start2_host, c2,
                                    start3_host, c3, span_first3, span_last3);
}

template <typename T, int P>
static int sg_launch_march2g(T *cp, const SgAdj2gArgs<T> &m, const SgSpanStarts<T> &ss, int nout, int64_t c1, const SgM2Maps &maps,
                             const SgM2Uni<T> *uni, cudaStream_t st)
{
    const size_t smem = sizeof(T) * ((size_t)SG_M2_NS * SG_M2_RTMAX * 128 + (size_t)(SG_M2_G2 + P) * SG_M2G_EP) + 128;
    dim3 grid((unsigned)m.nb1, (unsigned)m.tiles2, (unsigned)(m.chunks3 * nout));
    if (g_sg_prof_on) cudaEventRecord(g_sg_prof_ev[0], st);
    if (uni != nullptr && sg_env_int("SG_ADJ_M2_UNI", 0)) {             // weights of dimension 2 through the uniform datapath
        auto kern = sg_adj_march2g_kernel<T, P, SG_M2_G2, SG_M2_RTMAX, SG_M2_NS, true>;
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 160, smem, st>>>(m, maps, *uni);                   // 4 consumer warps + 1 producer warp
    } else {
        auto kern = sg_adj_march2g_kernel<T, P, SG_M2_G2, SG_M2_RTMAX, SG_M2_NS, false>;
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 160, smem, st>>>(m, maps, SgM2UniNone<T>{});
    }
    if (g_sg_prof_on) { cudaEventRecord(g_sg_prof_ev[1], st); g_sg_prof_recorded = 1; }
    dim3 cgrid((unsigned)((m.tiles2 + 1) * sg_blocks(c1, 128)), (unsigned)m.c3, (unsigned)nout);
    SgPushSpec ps{};                                                    // world == 0: plain adjoint
    if (g_sg_push != nullptr) { ps = *g_sg_push; g_sg_push_done = true; }
    sg_adj_combine2g_kernel<T, P, SG_M2_G2><<<cgrid, 128, 0, st>>>(cp, m.Y, ss.g_lo, m.bt_hdr, m.hdr, c1, m.c2, m.c3, m.tiles2, m.G3, m.chunks3,
                                                                   m.icap, m.nb1, ps);
    g_sg_launches.fetch_add(2);
    return SG_OK;
}

template <typename T>
static int sg_run_march2g(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr, const T *eval,
                          const SgMarch2Plan &mp, char *ws, const SgM2Maps &maps, const SgM2Uni<T> *uni, cudaStream_t st)
{
    SgAdj2gArgs<T> m{};
    m.X = eval; m.Y = reinterpret_cast<T *>(ws + mp.part_off);
    m.table2 = a.table[1]; m.table3 = a.table[2]; m.index3 = a.index[2];
    m.start2 = ss.start[1]; m.start3 = ss.start[2]; m.hdr = hdr;
    m.bt_hdr = ss.bt_hdr; m.bt_lol = ss.bt_lol; m.bt_w = ss.bt_w;
    m.n1 = a.n_samples[0]; m.n2 = a.n_samples[1]; m.n3 = a.n_samples[2]; m.c2 = a.n_cp[1]; m.c3 = a.n_cp[2];
    m.tiles2 = mp.tiles2; m.G3 = mp.G3; m.chunks3 = mp.chunks3; m.icap = ss.icap; m.rmcap = ss.rmcap; m.nb1 = ss.nb1;
    switch (a.degree[1]) {
        case 1: return sg_launch_march2g<T, 1>(cp, m, ss, a.nout, a.n_cp[0], maps, uni, st);
        case 2: return sg_launch_march2g<T, 2>(cp, m, ss, a.nout, a.n_cp[0], maps, uni, st);
        default: return sg_launch_march2g<T, 3>(cp, m, ss, a.nout, a.n_cp[0], maps, uni, st);
    }
}

template <typename T>
static int sg_run_march2(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr, const T *eval,
                         const SgMarch2Plan &mp, char *ws, const SgAdjKnown &known, const char **variant, cudaStream_t st)
{
    const int P = a.degree[1];
    *variant = "adjoint_march2";
    if (known.planned && known.fused_ok && mp.g.ok && ss.bt_hdr != nullptr && known.rows2_max <= SG_M2_FAST_ROWS) {
        SgM2Maps maps{};
        if (sg_m2_make_eval_maps<T>(maps, eval, a.n_samples[0], a.n_samples[1], a.n_samples[2] * a.nout)) {
            *variant = "adjoint_march2_fused";
            return sg_run_march2g<T>(cp, a, ss, hdr, eval, mp, ws, maps, static_cast<const SgM2Uni<T> *>(known.uni), st);
        }
    }
    SgAdj2Args<T> m{};
    T *part = reinterpret_cast<T *>(ws + mp.part_off);
    m.X = eval; m.Y = part; m.table2 = a.table[1]; m.table3 = a.table[2]; m.index3 = a.index[2];
    m.start2 = ss.start[1]; m.start3 = ss.start[2]; m.hdr = hdr;
    m.n1 = a.n_samples[0]; m.n2 = a.n_samples[1]; m.n3 = a.n_samples[2]; m.c2 = a.n_cp[1]; m.c3 = a.n_cp[2];
    m.tiles2 = mp.tiles2; m.G3 = mp.G3; m.chunks3 = mp.chunks3; m.path = SG_PATH_MULTIPASS;
    int lrc;
    switch (P) {
        case 1: lrc = sg_launch_march2<T, 1>(m, a.nout, known, st); break;
        case 2: lrc = sg_launch_march2<T, 2>(m, a.nout, known, st); break;
        default: lrc = sg_launch_march2<T, 3>(m, a.nout, known, st); break;
    }
    if (lrc != SG_OK) return lrc;
    {
        // halo combine + contraction of dimension 1 in one kernel (sg_adjoint_post2.cuh): the partials are read once and
        // the (n1, c2, c3) intermediate never goes to HBM
        dim3 pgrid((unsigned)((mp.tiles2 + 1) * sg_blocks(a.n_cp[0], 128)), (unsigned)m.c3, (unsigned)a.nout);
        SgPushSpec ps{};                                                // world == 0: plain adjoint
        if (g_sg_push != nullptr) { ps = *g_sg_push; g_sg_push_done = true; }
        // A planned call that only PUSHES (keep_local == 0: nothing is written outside the slab's support) launches the planes
        // of the support only: on an 8-way slab 19 of 128 planes, i.e. 85 % of the CTAs would load their tables and exit.
        int i3_top = (int)m.c3;
        if (known.planned && known.sf3 > 0 && ps.world > 0 && ps.keep_local == 0) {
            i3_top = (int)std::min<int64_t>(known.sl3, m.c3);
            const int i3_bot = std::max(known.sf3 - P, 1);
            if (i3_top < i3_bot) return SG_OK;
            pgrid.y = (unsigned)(i3_top - i3_bot + 1);
        }
        // drop the partials' dirty L2 lines after their only read: needs whole 128-byte lines per row (n1 a multiple of a
        // line) and ONE block of control indices (c1 <= 128), else neighbouring blocks read overlapping sample ranges
        const int discard = sg_env_int("SG_ADJ_DISCARD", 1) && (m.n1 * sizeof(T)) % 128 == 0 && reinterpret_cast<uintptr_t>(part) % 128 == 0 &&
                            a.n_cp[0] <= 128;
        switch (P) {
            case 1: sg_adj_post2_kernel<T, 1, SG_M2_G2><<<pgrid, 128, 0, st>>>(cp, part, a.table[0], a.index[0], ss.g_lo, ss.g_w, hdr, m.n1, a.n_cp[0], m.c2, m.c3, a.degree[0], mp.tiles2, mp.G3, mp.chunks3, SG_PATH_MULTIPASS, ps, discard, known.sf3, known.sl3, i3_top); break;
            case 2: sg_adj_post2_kernel<T, 2, SG_M2_G2><<<pgrid, 128, 0, st>>>(cp, part, a.table[0], a.index[0], ss.g_lo, ss.g_w, hdr, m.n1, a.n_cp[0], m.c2, m.c3, a.degree[0], mp.tiles2, mp.G3, mp.chunks3, SG_PATH_MULTIPASS, ps, discard, known.sf3, known.sl3, i3_top); break;
            default: sg_adj_post2_kernel<T, 3, SG_M2_G2><<<pgrid, 128, 0, st>>>(cp, part, a.table[0], a.index[0], ss.g_lo, ss.g_w, hdr, m.n1, a.n_cp[0], m.c2, m.c3, a.degree[0], mp.tiles2, mp.G3, mp.chunks3, SG_PATH_MULTIPASS, ps, discard, known.sf3, known.sl3, i3_top); break;
        }
        g_sg_launches.fetch_add(1);
        return SG_OK;
    }
    return SG_ERR_UNSUPPORTED;   // not reached: the plan guarantees G3 >= P and G2 >= P
}

template <typename T>
int sg_evaluate_adjoint_fast(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr,
                             const T *eval, const T *weights, void *scratch, const SgAdjKnown &known, cudaStream_t st)
{
    const bool rational = weights != nullptr;
    if (!sg_adjoint_fast_supported(a.nin, a.degree, rational)) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy != 2 && a.n_total < 32768) return SG_ERR_UNSUPPORTED;
    char *ws = static_cast<char *>(scratch);
    // zero fill (src/adjoint.jl:61): the double march's post kernel writes every control point itself
    const SgMarch2Plan mp = sg_adjoint_march2_plan(a.nin, a.n_samples, a.n_cp, a.nout, a.degree, (int)sizeof(T), rational);
    if (!mp.ok) SG_CUDA(cudaMemsetAsync(cp, 0, (size_t)a.cp_total * a.nout * sizeof(T), st));
    int rc;
    if (mp.ok) {
        const char *variant = "adjoint_march2";
        rc = sg_run_march2<T>(cp, a, ss, hdr, eval, mp, ws, known, &variant, st);
        g_sg_last_variant = variant;
    } else {
        const char *variant = rational ? "adjoint_passes_rational2d" : "adjoint_passes";
        rc = sg_run_multipass<T>(cp, a, ss, hdr, eval, weights, ws, SG_PATH_MULTIPASS, known, &variant, st);
        g_sg_last_variant = variant;
    }
    if (rc != SG_OK) return rc;
    if (known.planned) {                                               // the plan saw monotone spans: no fallback launch
        cudaError_t e = cudaPeekAtLastError();
        return e == cudaSuccess ? SG_OK : (int)e;
    }
    // non-monotone span indices (decided on device): the reference's atomic scatter
    const unsigned sblocks = (unsigned)std::min<int64_t>(sg_blocks(a.n_total, 256), 148 * 4);   // fallback: fixed small grid
    if (rational)
        sg_adjoint_scatter_kernel<T, true><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
    else
        sg_adjoint_scatter_kernel<T, false><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
    g_sg_launches.fetch_add(1);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? SG_OK : (int)e;
}

template int sg_evaluate_adjoint_fast<float>(float *, const SgGridArgs<float> &, const SgSpanStarts<float> &, SgAdjointHeader *,
                                             const float *, const float *, void *, const SgAdjKnown &, cudaStream_t);
template int sg_evaluate_adjoint_fast<double>(double *, const SgGridArgs<double> &, const SgSpanStarts<double> &, SgAdjointHeader *,
                                              const double *, const double *, void *, const SgAdjKnown &, cudaStream_t);
