// sg_fast_adjoint.cu -- plan + dispatch of the atomics-free adjoint passes (kernels: sg_fast_adjoint.cuh).
#include <cstdlib>

#include <algorithm>
#include <type_traits>

#include "sg_evaluate_generic.cuh"
#include "sg_fast.cuh"
#include "sg_fast_adjoint.cuh"
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

static SgAdjPlan sg_adjoint_plan(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout, const int *degree,
                                 int elem_size)
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
        const int64_t threads = ((ps.inner + V - 1) / V) * outer;
        int64_t nchunks = 1;
        if (threads < 65536) nchunks = std::min<int64_t>(nspans, (148 * 8 * 128 + threads - 1) / threads);
        const int forced = sg_env_int("SG_ADJ_CHUNKS", 0);
        if (forced > 0) nchunks = std::min<int64_t>(nspans, forced);
        ps.G = (int)((nspans + nchunks - 1) / nchunks);
        ps.nchunks = (int)((nspans + ps.G - 1) / ps.G);
        ps.part_off = off;
        if (ps.nchunks > 1) off += sg_al256((size_t)ps.inner * (ps.G + ps.P) * ps.nchunks * outer * elem_size);
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

size_t sg_adjoint_fast_scratch_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout, const int *degree,
                                     int elem_size)
{
    if (!sg_adjoint_fast_supported(nin, degree, false)) return 0;
    return sg_adjoint_plan(nin, n_samples, n_cp, nout, degree, elem_size).bytes;
}

template <typename T, int P, int NT, bool RAT2D>
static void sg_launch_adj_march_nt(const SgAdjPassArgs<T> &pa, int64_t outer, cudaStream_t st)
{
    constexpr int VV = sizeof(T) == 4 ? 4 : 2;
    const bool vec_ok = (pa.inner % VV == 0) && (reinterpret_cast<uintptr_t>(pa.X) % 16 == 0) &&
                        (reinterpret_cast<uintptr_t>(pa.Y) % 16 == 0);
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

template <typename T>
int sg_evaluate_adjoint_fast(T *cp, const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr,
                             const T *eval, const T *weights, void *scratch, cudaStream_t st)
{
    const bool rational = weights != nullptr;
    if (!sg_adjoint_fast_supported(a.nin, a.degree, rational)) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy != 2 && a.n_total < 32768) return SG_ERR_UNSUPPORTED;
    const SgAdjPlan pl = sg_adjoint_plan(a.nin, a.n_samples, a.n_cp, a.nout, a.degree, (int)sizeof(T));
    for (int k = 0; k < pl.npassA; ++k) {
        const SgAdjPass &ps = pl.pass[k];
        if (ps.outer > 65535 || ps.nchunks > 65535 || (ps.nchunks > 1 && ps.c_d > 65535)) return SG_ERR_UNSUPPORTED;   // grid.y / grid.z limits
    }
    char *ws = static_cast<char *>(scratch);
    // zero fill (src/adjoint.jl:61): needed only if the prep kernel flags non-monotone spans (scatter path)
    SG_CUDA(cudaMemsetAsync(cp, 0, (size_t)a.cp_total * a.nout * sizeof(T), st));

    const T *X = eval;
    for (int k = 0; k < pl.npassA; ++k) {
        const SgAdjPass &ps = pl.pass[k];
        T *out = reinterpret_cast<T *>(ws + ps.out_off);
        T *part = ps.nchunks > 1 ? reinterpret_cast<T *>(ws + ps.part_off) : out;
        SgAdjPassArgs<T> pa{};
        pa.X = X; pa.Y = part; pa.table = a.table[ps.d]; pa.index = a.index[ps.d]; pa.span_start = ss.start[ps.d];
        pa.hdr = hdr; pa.inner = ps.inner; pa.n_d = ps.n_d; pa.c_d = ps.c_d; pa.G = ps.G; pa.nchunks = ps.nchunks;
        // the first pass handles the (<= 4) output planes of one column in one thread
        const bool rat_here = rational && k == 0;
        const int nt = (rat_here && ps.outer <= 4) ? (int)ps.outer : 1;   // shares the denominators between the outputs   // nin == 2: the first pass marches dim 2 over columns of dim 1
        if (rat_here) { pa.weights = weights; pa.table1 = a.table[0]; pa.index1 = a.index[0]; pa.c1 = a.n_cp[0]; }
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
        if (ps.nchunks > 1) {
            constexpr int VC = sizeof(T) == 4 ? 4 : 2;
            const bool vec_ok = (ps.inner % VC == 0);
            dim3 cgrid((unsigned)((ps.inner + 128 * VC - 1) / (128 * VC)), (unsigned)ps.c_d, (unsigned)ps.outer);
            sg_adj_combine_kernel<T, VC><<<cgrid, 128, 0, st>>>(out, part, hdr, ps.inner, ps.c_d, ps.G, ps.nchunks, ps.P, vec_ok);
            g_sg_launches.fetch_add(1);
        }
        X = out;
    }
    // pass B: first dimension
    const int64_t outerB = a.cp_total / a.n_cp[0] * a.nout;
    const int64_t avg_range = (int64_t)(a.degree[0] + 1) * a.n_samples[0] / a.n_cp[0];
    auto launch_b = [&](auto lanes_tag, auto rat_tag) {
        constexpr int L = decltype(lanes_tag)::value;
        constexpr bool R = decltype(rat_tag)::value;
        const unsigned gy = (unsigned)std::min<int64_t>(outerB, 32768);
        dim3 bgrid(sg_blocks(a.n_cp[0], 256 / L), gy, (unsigned)((outerB + gy - 1) / gy));
        sg_adj_first_dim_kernel<T, L, R><<<bgrid, 256, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                  a.n_cp[0], outerB, a.degree[0], weights, a.cp_total);
    };
    if (avg_range >= 64) {
        if (rational) launch_b(std::integral_constant<int, 32>{}, std::true_type{});
        else launch_b(std::integral_constant<int, 32>{}, std::false_type{});
    } else if (avg_range <= 24 && sg_env_int("SG_ADJ_PASSB_ROWS", 1)) {
        // short ranges: per-thread weights in registers, many rows per block
        const int rpb = 32;
        dim3 rgrid(sg_blocks(a.n_cp[0], 128), (unsigned)((outerB + rpb - 1) / rpb));
        if (rgrid.y > 65535) return SG_ERR_UNSUPPORTED;
        if (rational)
            sg_adj_first_dim_rows_kernel<T, 32, true><<<rgrid, 128, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                            a.n_cp[0], outerB, a.degree[0], rpb, weights, a.cp_total);
        else
            sg_adj_first_dim_rows_kernel<T, 32, false><<<rgrid, 128, 0, st>>>(cp, X, a.table[0], a.index[0], ss.start[0], hdr, a.n_samples[0],
                                                                             a.n_cp[0], outerB, a.degree[0], rpb, weights, a.cp_total);
    } else {
        if (rational) launch_b(std::integral_constant<int, 8>{}, std::true_type{});
        else launch_b(std::integral_constant<int, 8>{}, std::false_type{});
    }
    g_sg_launches.fetch_add(1);
    // non-monotone span indices (decided on device): the reference's atomic scatter
    const unsigned sblocks = (unsigned)std::min<int64_t>(sg_blocks(a.n_total, 256), 148 * 16);   // fallback: fixed small grid
    if (rational)
        sg_adjoint_scatter_kernel<T, true><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
    else
        sg_adjoint_scatter_kernel<T, false><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
    g_sg_launches.fetch_add(1);
    g_sg_last_variant = rational ? "adjoint_passes_rational2d" : "adjoint_passes";
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? SG_OK : (int)e;
}

template int sg_evaluate_adjoint_fast<float>(float *, const SgGridArgs<float> &, const SgSpanStarts<float> &, SgAdjointHeader *,
                                             const float *, const float *, void *, cudaStream_t);
template int sg_evaluate_adjoint_fast<double>(double *, const SgGridArgs<double> &, const SgSpanStarts<double> &, SgAdjointHeader *,
                                              const double *, const double *, void *, cudaStream_t);
