// sg_evaluate.cu -- C ABI + dispatch for K3 (evaluate!) and K4 (evaluate_adjoint!).
// Reference launch sites: src/spline_grid.jl:200-230, src/adjoint.jl:52-83.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "sg_adjoint_generic.cuh"
#include "sg_evaluate_generic.cuh"
#include "sg_fast.cuh"
#include "sg_eval_multi.cuh"

static inline size_t sg_align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
template <typename T>
static int sg_evaluate_impl(T *eval, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                            const T *const *tables, const int32_t *const *indices, const int *degree,
                            const int *mdo, const int *der, const T *cp, const T *weights, void *stream)
{
    SG_NVTX("sg_evaluate");
    SG_CHECK_ARG(eval && cp);
    SgGridArgs<T> a;
    int rc = sg_fill_grid_args(a, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der);
    if (rc != SG_OK) return rc;
    cudaStream_t st = sg_stream(stream);
    if (g_sg_policy != 1) {
        rc = sg_evaluate_fast<T>(eval, a, cp, weights, st);
        if (rc != SG_ERR_UNSUPPORTED) return rc;  // SG_OK or a real error
    }
    const int threads = 256;
    const unsigned blocks = sg_blocks(a.n_total, threads);
    if (weights)
        sg_evaluate_generic_kernel<T, true><<<blocks, threads, 0, st>>>(eval, a, cp, weights);
    else
        sg_evaluate_generic_kernel<T, false><<<blocks, threads, 0, st>>>(eval, a, cp, weights);
    g_sg_last_variant = "evaluate_generic";
    SG_AFTER_LAUNCH();
    return SG_OK;
}

// K3 for several derivative orders: one fused launch where a multi kernel exists, else one launch per tuple
template <typename T>
static int sg_evaluate_multi_impl(T *const *evals, int n_der, const int *ders, int nin, const int64_t *n_samples, const int64_t *n_cp,
                                  int nout, const T *const *tables, const int32_t *const *indices, const int *degree, const int *mdo,
                                  const T *cp, const T *weights, void *stream)
{
    SG_NVTX("sg_evaluate_multi");
    SG_CHECK_ARG(evals && ders && n_der >= 1 && cp);
    for (int q = 0; q < n_der; ++q) SG_CHECK_ARG(evals[q] != nullptr);
    cudaStream_t st = sg_stream(stream);
    if (!weights && n_der >= 2 && n_der <= SG_MULTI_MAX && nin == 2) {
        SgMultiArgs<T> m{};
        SgGridArgs<T> a{};
        int rc = SG_OK;
        for (int q = 0; q < n_der && rc == SG_OK; ++q) {
            SgGridArgs<T> aq;
            rc = sg_fill_grid_args(aq, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, ders + (size_t)q * nin);
            m.eval[q] = evals[q]; m.table1[q] = aq.table[0]; m.table2[q] = aq.table[1];
            if (q == 0) a = aq;
        }
        if (rc != SG_OK) return rc;
        rc = sg_evaluate_multi_fast<T>(m, n_der, a, cp, st);
        if (rc != SG_ERR_UNSUPPORTED) return rc;
    }
    for (int q = 0; q < n_der; ++q) {
        int rc = sg_evaluate_impl<T>(evals[q], nin, n_samples, n_cp, nout, tables, indices, degree, mdo, ders + (size_t)q * nin, cp,
                                     weights, stream);
        if (rc != SG_OK) return rc;
    }
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
struct SgAdjointLayout {
    // prep region: written by the prep kernel (a plan keeps its own copy of this region)
    size_t header;               // offset 0
    size_t starts[SG_MAX_DIMS];  // int32[n_cp+2] per dimension
    size_t g_lo, g_w;            // gather table of dimension 1: int32[c_1][2], T[SG_GATHER_RMAX][c_1]
    size_t bt_hdr, bt_lol, bt_w; // column-block tables of the fused double march (plans only)
    SgM2gDims g;
    size_t prep_total;
    // scratch region (per call)
    size_t denom;                // T[n_total] (rational only), else 0
    size_t fast;                 // scratch of the tiled fast path
    size_t total;
};

static SgAdjointLayout sg_adjoint_layout(int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                         const int *degree, int elem_size, bool rational, bool block_tables)
{
    SgAdjointLayout L{};
    size_t off = sg_align256(sizeof(SgAdjointHeader));
    int64_t n_total = 1;
    for (int d = 0; d < nin; ++d) {
        L.starts[d] = off;
        off += sg_align256((size_t)(n_cp[d] + 2) * sizeof(int32_t));
        n_total *= n_samples[d];
    }
    L.g_lo = off;
    off += sg_align256((size_t)n_cp[0] * 2 * sizeof(int32_t));
    L.g_w = off;
    off += sg_align256((size_t)n_cp[0] * SG_GATHER_RMAX * elem_size);
    L.g = SgM2gDims{};
    if (block_tables) L.g = sg_m2g_dims(nin, n_samples, n_cp, degree, rational, elem_size);
    if (L.g.ok) {
        L.bt_hdr = off;
        off += sg_align256((size_t)L.g.nb1 * sizeof(SgM2gBlockHdr));
        L.bt_lol = off;
        off += sg_align256((size_t)L.g.nb1 * L.g.icap * sizeof(int32_t));
        L.bt_w = off;
        off += sg_align256((size_t)L.g.nb1 * L.g.rmcap * L.g.icap * elem_size);
    }
    L.prep_total = off;
    L.denom = off;
    if (rational) off += sg_align256((size_t)n_total * elem_size);
    L.fast = off;
    off += sg_align256(sg_adjoint_fast_scratch_bytes(nin, n_samples, n_cp, nout, degree, elem_size));
    L.total = off;
    return L;
}

extern "C" size_t sg_evaluate_adjoint_workspace_bytes(int nin, const int64_t *n_samples, const int64_t *n_cp,
                                                      int nout, const int *degree, int elem_size, int rational)
{
    if (nin < 1 || nin > SG_MAX_DIMS || !n_samples || !n_cp || !degree) return 0;
    return sg_adjoint_layout(nin, n_samples, n_cp, nout, degree, elem_size, rational != 0, false).total;
}

// ---- adjoint plan: the prep kernel's output, kept on the device, plus what the host needs to pick the pipeline ----
struct sg_adjoint_plan {
    int elem_size;
    int nin, nout;
    bool rational;
    int64_t n_samples[SG_MAX_DIMS], n_cp[SG_MAX_DIMS];
    const void *tables[SG_MAX_DIMS];
    const int32_t *indices[SG_MAX_DIMS];
    int degree[SG_MAX_DIMS], mdo[SG_MAX_DIMS], der[SG_MAX_DIMS];
    char *dev;                   // prep region (SgAdjointLayout offsets below prep_total)
    SgAdjointLayout L;
    SgAdjointHeader h;           // host copy of the header after the prep kernel
    void *uni;                   // host copy of dimension 2's weights and span starts (SgM2Uni<T>) or nullptr
};

template <typename T>
static void sg_fill_span_starts(SgSpanStarts<T> &ss, const SgAdjointLayout &L, char *prep, int nin)
{
    for (int d = 0; d < nin; ++d) ss.start[d] = reinterpret_cast<int32_t *>(prep + L.starts[d]);
    ss.g_lo = reinterpret_cast<int32_t *>(prep + L.g_lo);
    ss.g_w = reinterpret_cast<T *>(prep + L.g_w);
    ss.bt_hdr = nullptr; ss.bt_lol = nullptr; ss.bt_w = nullptr; ss.icap = ss.rmcap = ss.nb1 = 0; ss.bw = 128; ss.rfast = 0;
    if (L.g.ok) {
        ss.bt_hdr = reinterpret_cast<SgM2gBlockHdr *>(prep + L.bt_hdr);
        ss.bt_lol = reinterpret_cast<int32_t *>(prep + L.bt_lol);
        ss.bt_w = reinterpret_cast<T *>(prep + L.bt_w);
        ss.icap = L.g.icap; ss.rmcap = L.g.rmcap; ss.nb1 = L.g.nb1; ss.bw = L.g.bw; ss.rfast = L.g.rfast;
    }
}

template <typename T>
static int sg_launch_prep(const SgGridArgs<T> &a, const SgSpanStarts<T> &ss, SgAdjointHeader *hdr, cudaStream_t st)
{
    int64_t max_len = 2;
    for (int d = 0; d < a.nin; ++d) max_len = std::max(max_len, std::max(a.n_samples[d], a.n_cp[d] + 2));
    SG_CUDA(cudaMemsetAsync(hdr, 0, sizeof(SgAdjointHeader), st));
    // rows of blocks: one per dimension, + 1: gather table of dimension 1, + 1: column-block tables (plans of 3-D grids)
    dim3 pgrid((unsigned)std::min<int64_t>((max_len + 255) / 256, 64), a.nin + (ss.bt_hdr ? 2 : 1));
    sg_adjoint_prep_kernel<T><<<pgrid, 256, 0, st>>>(a, ss, hdr);
    g_sg_launches.fetch_add(1);
    return SG_OK;
}

template <typename T>
static int sg_evaluate_adjoint_impl(T *cp, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                    const T *const *tables, const int32_t *const *indices, const int *degree,
                                    const int *mdo, const int *der, const T *eval, const T *weights,
                                    void *workspace, size_t workspace_bytes, void *stream, const sg_adjoint_plan *plan = nullptr)
{
    SG_NVTX("sg_evaluate_adjoint");
    SG_CHECK_ARG(eval && cp);
    SgGridArgs<T> a;
    int rc = sg_fill_grid_args(a, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der);
    if (rc != SG_OK) return rc;
    cudaStream_t st = sg_stream(stream);
    const bool rational = weights != nullptr;
    const size_t cp_bytes = (size_t)a.cp_total * nout * sizeof(T);

    // Tiny problems: the reference's own algorithm (zero fill + atomic scatter), 2 launches.
    const double terms = (double)a.n_total * (double)a.n_window * nout;
    if (g_sg_policy != 2 && terms <= 262144.0) {
        SG_CUDA(cudaMemsetAsync(cp, 0, cp_bytes, st));
        const unsigned blocks = sg_blocks(a.n_total, 256);
        if (rational)
            sg_adjoint_scatter_kernel<T, true><<<blocks, 256, 0, st>>>(cp, a, nullptr, eval, weights);
        else
            sg_adjoint_scatter_kernel<T, false><<<blocks, 256, 0, st>>>(cp, a, nullptr, eval, weights);
        g_sg_last_variant = "adjoint_scatter_small";
        SG_AFTER_LAUNCH();
        return SG_OK;
    }

    // A plan whose prep kernel saw non-monotone spans is of no use: take the unplanned route (device-side decisions).
    if (plan && (plan->h.nonmonotone != 0 || g_sg_policy == 1)) plan = nullptr;
    const SgAdjointLayout L = sg_adjoint_layout(nin, n_samples, n_cp, nout, degree, (int)sizeof(T), rational, false);
    char *ws = static_cast<char *>(workspace);
    bool own = false;
    if (ws) {
        if (workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return SG_ERR_WORKSPACE;
    } else {
        SG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ws), L.total, st));
        own = true;
    }
    // prep region: the plan's (already filled) or the head of the workspace (filled now)
    char *prep = plan ? plan->dev : ws;
    const SgAdjointLayout &PL = plan ? plan->L : L;
    SgAdjointHeader *hdr = reinterpret_cast<SgAdjointHeader *>(prep + PL.header);
    SgSpanStarts<T> ss{};
    sg_fill_span_starts<T>(ss, PL, prep, nin);
    SgAdjKnown known{};
    known.planned = plan != nullptr;
    known.fused_ok = plan != nullptr && PL.g.ok && plan->h.m2g_bad == 0;
    known.rows2_max = plan ? plan->h.rows2_max : 0;
    known.uni = plan ? plan->uni : nullptr;
    known.sf3 = (plan && nin == 3) ? plan->h.span_first[2] : 0;
    known.sl3 = (plan && nin == 3) ? plan->h.span_last[2] : 0;
    rc = SG_OK;
    do {
        cudaError_t e;
        if (!plan) {
            rc = sg_launch_prep<T>(a, ss, hdr, st);
            if (rc != SG_OK) break;
        }
        if (g_sg_policy != 1) {
            int frc = sg_evaluate_adjoint_fast<T>(cp, a, ss, hdr, eval, weights, ws + L.fast, known, st);
            if (frc != SG_ERR_UNSUPPORTED) { rc = frc; break; }
        }
        // generic: zero fill (src/adjoint.jl:61) needed by the scatter branch only, but the branch is
        // chosen on device, so always done.
        e = cudaMemsetAsync(cp, 0, cp_bytes, st);
        if (e != cudaSuccess) { rc = (int)e; break; }
        T *denom = nullptr;
        if (rational) {
            // denom[J] = sum_I prod_d B_d * w[base+I]: a non-rational forward pass with cp := w, Nout := 1
            denom = reinterpret_cast<T *>(ws + L.denom);
            SgGridArgs<T> a1 = a;
            a1.nout = 1;
            sg_evaluate_generic_kernel<T, false><<<sg_blocks(a.n_total, 256), 256, 0, st>>>(denom, a1, weights, nullptr);
            g_sg_launches.fetch_add(1);
        }
        const unsigned gblocks = sg_blocks(a.cp_total, 128);
        const unsigned sblocks = (unsigned)std::min<int64_t>(sg_blocks(a.n_total, 256), 148 * 16);   // fallback: fixed small grid
        if (rational) {
            sg_adjoint_gather_kernel<T, true><<<gblocks, 128, 0, st>>>(cp, a, ss, hdr, eval, weights, denom);
            if (!plan) sg_adjoint_scatter_kernel<T, true><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
        } else {
            sg_adjoint_gather_kernel<T, false><<<gblocks, 128, 0, st>>>(cp, a, ss, hdr, eval, weights, denom);
            if (!plan) sg_adjoint_scatter_kernel<T, false><<<sblocks, 256, 0, st>>>(cp, a, hdr, eval, weights);
        }
        g_sg_launches.fetch_add(plan ? 1 : 2);
        g_sg_last_variant = "adjoint_gather_generic";
        e = cudaPeekAtLastError();
        if (e != cudaSuccess) rc = (int)e;
    } while (0);
    if (own) {
        cudaError_t e = cudaFreeAsync(ws, st);
        if (rc == SG_OK && e != cudaSuccess) rc = (int)e;
    }
    return rc;
}

template <typename T>
static int sg_adjoint_plan_create_impl(sg_adjoint_plan **out, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                       const T *const *tables, const int32_t *const *indices, const int *degree, const int *mdo,
                                       const int *der, int rational, void *stream)
{
    SG_NVTX("sg_adjoint_plan_create");
    SG_CHECK_ARG(out);
    *out = nullptr;
    SgGridArgs<T> a;
    int rc = sg_fill_grid_args(a, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der);
    if (rc != SG_OK) return rc;
    cudaStream_t st = sg_stream(stream);
    sg_adjoint_plan *p = new sg_adjoint_plan{};
    p->elem_size = (int)sizeof(T); p->nin = nin; p->nout = nout; p->rational = rational != 0;
    for (int d = 0; d < nin; ++d) {
        p->n_samples[d] = n_samples[d]; p->n_cp[d] = n_cp[d]; p->tables[d] = tables[d]; p->indices[d] = indices[d];
        p->degree[d] = degree[d]; p->mdo[d] = mdo[d]; p->der[d] = der[d];
    }
    p->L = sg_adjoint_layout(nin, n_samples, n_cp, nout, degree, (int)sizeof(T), p->rational, true);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->dev), p->L.prep_total);
    if (e != cudaSuccess) { delete p; return (int)e; }
    SgSpanStarts<T> ss{};
    sg_fill_span_starts<T>(ss, p->L, p->dev, nin);
    SgAdjointHeader *hdr = reinterpret_cast<SgAdjointHeader *>(p->dev + p->L.header);
    rc = sg_launch_prep<T>(a, ss, hdr, st);
    if (rc == SG_OK) {
        e = cudaMemcpyAsync(&p->h, hdr, sizeof(SgAdjointHeader), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc == SG_OK && nin == 3 && !p->rational && p->h.nonmonotone == 0 && degree[1] == degree[2] && degree[1] >= 1 && degree[1] <= 3) {
        // host copy of dimension 2's selected table slice and span starts: kernel parameter of the fused double march
        const int64_t n2 = n_samples[1], c2 = n_cp[1];
        const int P = degree[1];
        std::vector<T> tb((size_t)n2 * (P + 1));
        std::vector<int32_t> s2((size_t)c2 + 2), s3((size_t)n_cp[2] + 2);
        e = cudaMemcpyAsync(tb.data(), a.table[1], tb.size() * sizeof(T), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s2.data(), ss.start[1], s2.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s3.data(), ss.start[2], s3.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = (int)e;
        else {
            p->uni = std::malloc(sg_m2_uni_bytes((int)sizeof(T)));
            if (p->uni && !sg_m2_uni_fill(p->uni, (int)sizeof(T), tb.data(), n2, P, s2.data(), c2, s3.data(), n_cp[2], p->h.span_first[2], p->h.span_last[2])) {
                std::free(p->uni);
                p->uni = nullptr;
            }
        }
    }
    if (rc != SG_OK) { cudaFree(p->dev); std::free(p->uni); delete p; return rc; }
    *out = p;
    return SG_OK;
}

extern "C" int sg_adjoint_plan_destroy(sg_adjoint_plan *plan)
{
    if (!plan) return SG_OK;
    cudaError_t e = cudaFree(plan->dev);
    std::free(plan->uni);
    delete plan;
    return e == cudaSuccess ? SG_OK : (int)e;
}

extern "C" int sg_adjoint_plan_info(const sg_adjoint_plan *plan, int *monotone, int *fused_tables_fit, int *rows2_max)
{
    SG_CHECK_ARG(plan);
    if (monotone) *monotone = plan->h.nonmonotone == 0;
    if (fused_tables_fit) *fused_tables_fit = plan->L.g.ok && plan->h.m2g_bad == 0;
    if (rows2_max) *rows2_max = plan->h.rows2_max;
    return SG_OK;
}

// Support-plane exchange, host side (pure host code, no device needed): the LOCAL planes [dst_lo[r], dst_hi[r]) of rank my_rank's
// support that rank r's slab touches too, i.e. the overlap of the two supports (ranks whose supports do not meet receive nothing;
// the own slot always gets every plane).
extern "C" int sg_exchange_support_ranges(int world, int my_rank, const int64_t *k0s, const int64_t *nps, int64_t max_planes,
                                          int *dst_lo, int *dst_hi)
{
    if (!k0s || !nps || !dst_lo || !dst_hi || world < 1 || world > SG_MAX_PEERS || my_rank < 0 || my_rank >= world || max_planes < 0)
        return SG_ERR_INVALID_ARGUMENT;
    const int full = (int)std::min<int64_t>(max_planes, INT32_MAX);
    for (int r = 0; r < world; ++r) {
        if (r == my_rank) { dst_lo[r] = 0; dst_hi[r] = full; continue; }
        const int64_t lo = std::max(k0s[r], k0s[my_rank]) - k0s[my_rank];
        const int64_t hi = std::min(k0s[r] + nps[r], k0s[my_rank] + nps[my_rank]) - k0s[my_rank];
        dst_lo[r] = (int)std::min<int64_t>(std::max<int64_t>(lo, 0), full);
        dst_hi[r] = (int)std::max<int64_t>(std::min<int64_t>(hi, full), dst_lo[r]);
    }
    return SG_OK;
}

extern "C" int sg_exchange_push_f32(const float *, void *const *, int, int, int64_t, int64_t, int, int64_t, int64_t, int64_t, void *);
extern "C" int sg_exchange_push_f64(const double *, void *const *, int, int, int64_t, int64_t, int, int64_t, int64_t, int64_t, void *);

// sg_evaluate_adjoint followed by sg_exchange_push as ONE call; the double march's last kernel does the push itself.
template <typename T, typename PushFn>
static int sg_adjoint_push_impl(const sg_adjoint_plan *plan, T *cp, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                                const T *const *tables, const int32_t *const *indices, const int *degree, const int *mdo,
                                const int *der, const T *eval, const T *weights, void *workspace, size_t workspace_bytes,
                                void *const *peer_stage, int world, int my_rank, int64_t k0, int64_t np, int64_t max_planes,
                                int keep_local, void *stream, PushFn push_fn, void *multicast_stage = nullptr,
                                const int64_t *k0s = nullptr, const int64_t *nps = nullptr)
{
    if (!peer_stage || world < 1 || world > SG_MAX_PEERS || my_rank < 0 || my_rank >= world || nin < 1)
        return SG_ERR_INVALID_ARGUMENT;
    SgPushSpec spec{};
    for (int r = 0; r < world; ++r) spec.stage[r] = peer_stage[r];
    spec.world = world; spec.my_rank = my_rank; spec.max_planes = max_planes; spec.keep_local = keep_local;
    spec.n_dst = world;
    for (int r = 0; r < SG_MAX_PEERS; ++r) { spec.dst_lo[r] = 0; spec.dst_hi[r] = (int)std::min<int64_t>(max_planes, INT32_MAX); }
    if (k0s && nps) {
        int rc = sg_exchange_support_ranges(world, my_rank, k0s, nps, max_planes, spec.dst_lo, spec.dst_hi);
        if (rc != SG_OK) return rc;
    } else if (multicast_stage != nullptr) { spec.stage[0] = multicast_stage; spec.n_dst = 1; }   // one store reaches every rank
    g_sg_push = &spec; g_sg_push_done = false;
    int rc = sg_evaluate_adjoint_impl<T>(cp, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der, eval, weights,
                                         workspace, workspace_bytes, stream, plan);
    g_sg_push = nullptr;
    if (rc == SG_OK && !g_sg_push_done) {   /* another pipeline ran: separate push kernel */
        int64_t plane_elems = 1;
        for (int d = 0; d + 1 < nin; ++d) plane_elems *= n_cp[d];
        rc = push_fn(cp, peer_stage, world, my_rank, plane_elems, n_cp[nin - 1], nout, k0, np, max_planes, stream);
    }
    return rc;
}

#define SG_DEFINE_EVAL_API(T, SUF)                                                                                   \
    extern "C" int sg_evaluate_##SUF(T *eval, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,      \
                                     const T *const *tables, const int32_t *const *indices, const int *degree,       \
                                     const int *mdo, const int *der, const T *cp, const T *weights, void *stream)    \
    {                                                                                                                \
        return sg_evaluate_impl<T>(eval, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der, cp, weights, \
                                   stream);                                                                          \
    }                                                                                                                \
    extern "C" int sg_evaluate_multi_##SUF(T *const *evals, int n_der, const int *ders, int nin, const int64_t *n_samples,     \
                                           const int64_t *n_cp, int nout, const T *const *tables,                   \
                                           const int32_t *const *indices, const int *degree, const int *mdo,         \
                                           const T *cp, const T *weights, void *stream)                              \
    {                                                                                                                \
        return sg_evaluate_multi_impl<T>(evals, n_der, ders, nin, n_samples, n_cp, nout, tables, indices, degree,    \
                                         mdo, cp, weights, stream);                                                  \
    }                                                                                                                \
    extern "C" int sg_evaluate_adjoint_##SUF(T *cp, int nin, const int64_t *n_samples, const int64_t *n_cp,          \
                                             int nout, const T *const *tables, const int32_t *const *indices,        \
                                             const int *degree, const int *mdo, const int *der, const T *eval,       \
                                             const T *weights, void *workspace, size_t workspace_bytes,              \
                                             void *stream)                                                           \
    {                                                                                                                \
        return sg_evaluate_adjoint_impl<T>(cp, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der, eval,  \
                                           weights, workspace, workspace_bytes, stream);                             \
    }                                                                                                                \
    extern "C" int sg_evaluate_adjoint_push_##SUF(T *cp, int nin, const int64_t *n_samples, const int64_t *n_cp,     \
                                                  int nout, const T *const *tables, const int32_t *const *indices,   \
                                                  const int *degree, const int *mdo, const int *der, const T *eval,  \
                                                  const T *weights, void *workspace, size_t workspace_bytes,         \
                                                  void *const *peer_stage, int world, int my_rank, int64_t k0,       \
                                                  int64_t np, int64_t max_planes, int keep_local, void *stream)      \
    {                                                                                                                \
        return sg_adjoint_push_impl<T>(nullptr, cp, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der,   \
                                       eval, weights, workspace, workspace_bytes, peer_stage, world, my_rank, k0,    \
                                       np, max_planes, keep_local, stream, sg_exchange_push_##SUF);                  \
    }                                                                                                                \
    extern "C" int sg_adjoint_plan_create_##SUF(sg_adjoint_plan **plan, int nin, const int64_t *n_samples,           \
                                                const int64_t *n_cp, int nout, const T *const *tables,               \
                                                const int32_t *const *indices, const int *degree, const int *mdo,    \
                                                const int *der, int rational, void *stream)                          \
    {                                                                                                                \
        return sg_adjoint_plan_create_impl<T>(plan, nin, n_samples, n_cp, nout, tables, indices, degree, mdo, der,   \
                                              rational, stream);                                                     \
    }                                                                                                                \
    extern "C" int sg_evaluate_adjoint_planned_##SUF(const sg_adjoint_plan *plan, T *cp, const T *eval,              \
                                                     const T *weights, void *workspace, size_t workspace_bytes,      \
                                                     void *const *peer_stage, int world, int my_rank, int64_t k0,    \
                                                     int64_t np, int64_t max_planes, int keep_local,                 \
                                                     void *multicast_stage, void *stream)                            \
    {                                                                                                                \
        if (!plan || plan->elem_size != (int)sizeof(T) || (plan->rational != (weights != nullptr)))                  \
            return SG_ERR_INVALID_ARGUMENT;                                                                          \
        const T *tb[SG_MAX_DIMS];                                                                                    \
        for (int d = 0; d < plan->nin; ++d) tb[d] = static_cast<const T *>(plan->tables[d]);                         \
        if (!peer_stage)                                                                                             \
            return sg_evaluate_adjoint_impl<T>(cp, plan->nin, plan->n_samples, plan->n_cp, plan->nout, tb,           \
                                               plan->indices, plan->degree, plan->mdo, plan->der, eval, weights,     \
                                               workspace, workspace_bytes, stream, plan);                            \
        return sg_adjoint_push_impl<T>(plan, cp, plan->nin, plan->n_samples, plan->n_cp, plan->nout, tb,             \
                                       plan->indices, plan->degree, plan->mdo, plan->der, eval, weights, workspace,  \
                                       workspace_bytes, peer_stage, world, my_rank, k0, np, max_planes, keep_local,  \
                                       stream, sg_exchange_push_##SUF, multicast_stage);                             \
    }                                                                                                                \
    extern "C" int sg_evaluate_adjoint_planned_support_##SUF(const sg_adjoint_plan *plan, T *cp, const T *eval,      \
                                                     const T *weights, void *workspace, size_t workspace_bytes,      \
                                                     void *const *peer_stage, int world, int my_rank,                \
                                                     const int64_t *k0s, const int64_t *nps, int64_t max_planes,     \
                                                     int keep_local, void *stream)                                   \
    {                                                                                                                \
        if (!plan || plan->elem_size != (int)sizeof(T) || (plan->rational != (weights != nullptr)) || !peer_stage || \
            !k0s || !nps || world < 1 || my_rank < 0 || my_rank >= world)                                            \
            return SG_ERR_INVALID_ARGUMENT;                                                                          \
        const T *tb[SG_MAX_DIMS];                                                                                    \
        for (int d = 0; d < plan->nin; ++d) tb[d] = static_cast<const T *>(plan->tables[d]);                         \
        return sg_adjoint_push_impl<T>(plan, cp, plan->nin, plan->n_samples, plan->n_cp, plan->nout, tb,             \
                                       plan->indices, plan->degree, plan->mdo, plan->der, eval, weights, workspace,  \
                                       workspace_bytes, peer_stage, world, my_rank, k0s[my_rank], nps[my_rank],      \
                                       max_planes, keep_local, stream, sg_exchange_push_##SUF, nullptr, k0s, nps);   \
    }

SG_DEFINE_EVAL_API(float, f32)
SG_DEFINE_EVAL_API(double, f64)
