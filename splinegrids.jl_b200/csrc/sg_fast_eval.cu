// sg_fast_eval.cu -- dispatch of the tiled sm_100a fast paths for evaluate! (kernels: sg_fast_eval.cuh).
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <type_traits>

#include "sg_evaluate_generic.cuh"
#include "sg_fast.cuh"
#include "sg_fast_adjoint.cuh"
#include "sg_fast_eval.cuh"
#include "sg_eval_multi.cuh"

static int sg_env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

static bool sg_uniform_degree(const int *deg, int nin, int &p)
{
    p = deg[0];
    for (int d = 1; d < nin; ++d)
        if (deg[d] != p) return false;
    return true;
}

// number of marching chunks so that the grid has a few waves of CTAs on 148 SMs
static int sg_pick_chunk(int64_t n_march, int64_t col_tiles, int min_chunk, int max_chunk, int env_override,
                         int64_t target_ctas = 148 * 8)
{
    if (env_override > 0) return (int)std::min<int64_t>(std::max(env_override, 1), std::min<int64_t>(n_march, max_chunk));
    int64_t nchunks = (target_ctas + col_tiles - 1) / col_tiles;
    nchunks = std::max<int64_t>(1, std::min<int64_t>(nchunks, (n_march + min_chunk - 1) / min_chunk));
    int64_t chunk = (n_march + nchunks - 1) / nchunks;
    chunk = (chunk + 7) / 8 * 8;                       // multiples of 8 steps: fewer ragged tail chunks
    chunk = std::min<int64_t>(std::max<int64_t>(chunk, 1), max_chunk);
    return (int)chunk;
}

// ---- TMA tensor map for the control points (c1, c2, c3, nout), box (B1, B2, B3, 1) ------------------------
typedef CUresult (*SgEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SgEncodeTiledFn sg_get_encoder()
{
    static SgEncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<SgEncodeTiledFn>(p);
    }();
    return fn;
}

template <typename T>
static bool sg_make_cp_tensor_map(CUtensorMap &tm, const T *cp, const SgGridArgs<T> &a)
{
    SgEncodeTiledFn enc = sg_get_encoder();
    if (!enc) return false;
    const cuuint64_t c1 = a.n_cp[0], c2 = a.n_cp[1], c3 = a.n_cp[2];
    if ((c1 * sizeof(T)) % 16 != 0 || reinterpret_cast<uintptr_t>(cp) % 16 != 0) return false;   // TMA stride/base alignment
    cuuint64_t dims[4] = {c1, c2, c3, (cuuint64_t)a.nout};
    cuuint64_t strides[3] = {c1 * sizeof(T), c1 * c2 * sizeof(T), c1 * c2 * c3 * sizeof(T)};   // bytes, dims 1..3
    cuuint32_t box[4] = {SG_TMA_B1, SG_TMA_B2, SG_TMA_B3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(&tm, dt, 4, const_cast<T *>(cp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Chunk of the marching axis from the kernel's RESIDENT capacity: full waves, with a preference for long chunks (a CTA pays
// ~8 planes' worth of start-up: window load + first window contraction).  Measured on the 8-way slab of C3 (64 planes, 256
// column tiles, capacity 296): one chunk of 64 planes 0.0315 ms, two of 32 (512 CTAs = 1.7 waves) 0.0356 ms.
static int sg_pick_chunk_waves(int64_t n_march, int64_t col_tiles, int64_t capacity, int max_chunk)
{
    int best = (int)std::min<int64_t>(n_march, max_chunk);
    double best_score = -1.0;
    for (int64_t nch = 1; nch <= std::min<int64_t>(n_march, 64); ++nch) {
        int64_t chunk = (n_march + nch - 1) / nch;
        chunk = (chunk + 7) / 8 * 8;
        if (chunk > max_chunk) continue;
        const int64_t ne = (n_march + chunk - 1) / chunk;
        const double ctas = (double)col_tiles * ne, waves = std::ceil(ctas / (double)capacity);
        const double score = ctas / (waves * capacity) * (double)chunk / ((double)chunk + 8.0);
        if (score > best_score) { best_score = score; best = (int)chunk; }
    }
    return best;
}

template <typename T, int P, int V1, int V2, int TY>
static int sg_launch_eval3d(T *eval, const SgGridArgs<T> &a, const T *cp, cudaStream_t st)
{
    const int64_t gx = (a.n_samples[0] + 32 * V1 - 1) / (32 * V1);
    const int64_t gy = (a.n_samples[1] + TY * V2 - 1) / (TY * V2);
    // (the TMA variant stages at most SG_TMA_B3 control planes per CTA: a finer cut of the marching axis suits it)
    int chunk = sg_pick_chunk(a.n_samples[2], gx * gy, 32, 1024, sg_env_int("SG_CHUNK3D", 0), 148 * 14);
    if (sg_env_int("SG_CHUNK3D", 0) == 0 && sg_env_int("SG_EVAL_WAVES", 1)) {
        static const int64_t capacity = [] {
            int per_sm = 0, dev = 0, sms = 148;
            auto kern = sg_eval3d_march_kernel<T, P, V1, V2, TY, true>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            const size_t smem = sizeof(T) * SG_TMA_B1 * SG_TMA_B2 * SG_TMA_B3 + 64 * ((P + 1) * sizeof(T) + sizeof(int)) + 32;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * TY, smem) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                return (int64_t)0;
            }
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            return (int64_t)per_sm * sms;
        }();
        if (capacity > 0) chunk = sg_pick_chunk_waves(a.n_samples[2], gx * gy, capacity, 96);
    }
    const int64_t gz = (a.n_samples[2] + chunk - 1) / chunk;
    if (gy > 65535 || gz > 65535) return SG_ERR_UNSUPPORTED;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(eval) % 16 == 0) && (a.n_samples[0] % V1 == 0);
    const size_t smem_march = (((size_t)chunk * ((P + 1) * sizeof(T) + sizeof(int)) + 15) & ~(size_t)15) + 16;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32, TY);
    CUtensorMap tm{};
    if (sg_env_int("SG_EVAL_TMA", 1) && sg_make_cp_tensor_map<T>(tm, cp, a)) {
        // control-point window staged by one TMA bulk tensor copy per CTA
        const size_t smem = sizeof(T) * SG_TMA_B1 * SG_TMA_B2 * SG_TMA_B3 + smem_march;
        auto kern = sg_eval3d_march_kernel<T, P, V1, V2, TY, true>;
        SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (smem <= 160 * 1024) {
            kern<<<grid, block, smem, st>>>(eval, a, cp, chunk, a.nout, vec_ok, tm);
            g_sg_last_variant = "evaluate_march3d_tma";
            SG_AFTER_LAUNCH();
            return SG_OK;
        }
    }
    sg_eval3d_march_kernel<T, P, V1, V2, TY, false><<<grid, block, smem_march, st>>>(eval, a, cp, chunk, a.nout, vec_ok, tm);
    g_sg_last_variant = "evaluate_march3d";
    SG_AFTER_LAUNCH();
    return SG_OK;
}

template <typename T, int P, int V1, int NT, bool NURBS>
static int sg_launch_eval2d(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st)
{
    const int threads = 128;
    const int64_t gx = (a.n_samples[0] + threads * V1 - 1) / (threads * V1);
    const int64_t gz = (a.nout + NT - 1) / NT;
    const int chunk = sg_pick_chunk(a.n_samples[1], gx * gz, 32, 1024, sg_env_int("SG_CHUNK2D", 0));
    const int64_t gy = (a.n_samples[1] + chunk - 1) / chunk;
    if (gy > 65535 || gz > 65535) return SG_ERR_UNSUPPORTED;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(eval) % 16 == 0) && (a.n_samples[0] % V1 == 0);
    const size_t smem = (size_t)chunk * ((P + 1) * sizeof(T) + sizeof(int));
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
    sg_eval2d_march_kernel<T, P, V1, NT, NURBS><<<grid, threads, smem, st>>>(eval, a, cp, weights, chunk, vec_ok);
    g_sg_last_variant = NURBS ? "evaluate_march2d_nurbs" : "evaluate_march2d";
    SG_AFTER_LAUNCH();
    return SG_OK;
}

template <typename T, int P, bool NURBS>
static int sg_eval2d_by_nout(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st)
{
    constexpr int V1 = sizeof(T) == 4 ? 4 : 2;
    switch (a.nout >= 4 ? 4 : a.nout) {
        case 1: return sg_launch_eval2d<T, P, V1, 1, NURBS>(eval, a, cp, weights, st);
        case 2: return sg_launch_eval2d<T, P, V1, 2, NURBS>(eval, a, cp, weights, st);
        case 3: return sg_launch_eval2d<T, P, V1, 3, NURBS>(eval, a, cp, weights, st);
        default: return sg_launch_eval2d<T, P, V1, 4, NURBS>(eval, a, cp, weights, st);
    }
}

template <typename T, int P>
static int sg_eval3d_variant(T *eval, const SgGridArgs<T> &a, const T *cp, cudaStream_t st)
{
    constexpr int V1 = sizeof(T) == 4 ? 4 : 2;
    if (P == 3) {   // tuning variants for the cubic case
        switch (sg_env_int("SG_EVAL3D_VARIANT", 0)) {
            case 1: return sg_launch_eval3d<T, 3, V1, 2, 4>(eval, a, cp, st);
            case 2: return sg_launch_eval3d<T, 3, V1, 4, 8>(eval, a, cp, st);
            case 3: return sg_launch_eval3d<T, 3, V1, 2, 8>(eval, a, cp, st);
            case 4: return sg_launch_eval3d<T, 3, V1, 1, 8>(eval, a, cp, st);
            default: return sg_launch_eval3d<T, 3, V1, 4, 4>(eval, a, cp, st);
        }
    }
    return sg_launch_eval3d<T, P, V1, 2, 4>(eval, a, cp, st);
}

template <typename T>
static int sg_evaluate_fast_uniform(T *eval, const SgGridArgs<T> &a, int p, const T *cp, const T *weights, cudaStream_t st)
{
    if (a.nin == 3) {
        if (weights) return SG_ERR_UNSUPPORTED;
        switch (p) {
            case 1: return sg_eval3d_variant<T, 1>(eval, a, cp, st);
            case 2: return sg_eval3d_variant<T, 2>(eval, a, cp, st);
            default: return sg_eval3d_variant<T, 3>(eval, a, cp, st);
        }
    }
    if (weights) {
        switch (p) {
            case 1: return sg_eval2d_by_nout<T, 1, true>(eval, a, cp, weights, st);
            case 2: return sg_eval2d_by_nout<T, 2, true>(eval, a, cp, weights, st);
            default: return sg_eval2d_by_nout<T, 3, true>(eval, a, cp, weights, st);
        }
    }
    switch (p) {
        case 1: return sg_eval2d_by_nout<T, 1, false>(eval, a, cp, weights, st);
        case 2: return sg_eval2d_by_nout<T, 2, false>(eval, a, cp, weights, st);
        default: return sg_eval2d_by_nout<T, 3, false>(eval, a, cp, weights, st);
    }
}

template <typename T>
int sg_evaluate_fast(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st)
{
    if (a.nin != 2 && a.nin != 3) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy != 2 && a.n_total < 32768) return SG_ERR_UNSUPPORTED;   // launch-latency regime: generic kernel
    int p;
    if (sg_uniform_degree(a.degree, a.nin, p)) {
        if (p < 1 || p > 3) return SG_ERR_UNSUPPORTED;
        return sg_evaluate_fast_uniform<T>(eval, a, p, cp, weights, st);
    }
    // Mixed degrees (the README grid has (2, 3, 2)): pad every dimension's table to the largest degree with leading zero
    // columns and run the uniform-degree march kernel of that degree.  Costs one tiny launch per padded dimension and a
    // stream-ordered scratch allocation, so only grids past the launch-latency regime take it.
    int pmax = 0;
    for (int d = 0; d < a.nin; ++d) pmax = std::max(pmax, a.degree[d]);
    if (pmax < 1 || pmax > 3 || sg_env_int("SG_EVAL_MIXED", 1) == 0) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy != 2 && a.n_total < 262144) return SG_ERR_UNSUPPORTED;
    SgGridArgs<T> a2 = a;
    size_t off[SG_MAX_DIMS], total = 0;
    for (int d = 0; d < a.nin; ++d) {
        off[d] = total;
        if (a.degree[d] != pmax) total += ((size_t)a.n_samples[d] * (pmax + 1) * sizeof(T) + 255) & ~(size_t)255;
    }
    char *scratch = nullptr;
    SG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), total, st));
    a2.n_window = 1;
    for (int d = 0; d < a.nin; ++d) {
        if (a.degree[d] != pmax) {
            T *dst = reinterpret_cast<T *>(scratch + off[d]);
            sg_pad_table_kernel<T><<<sg_blocks(a.n_samples[d], 128), 128, 0, st>>>(dst, a.table[d], a.n_samples[d], a.degree[d], pmax);
            g_sg_launches.fetch_add(1);
            a2.table[d] = dst;
            a2.degree[d] = pmax;
        }
        a2.n_window *= pmax + 1;
    }
    int rc = sg_evaluate_fast_uniform<T>(eval, a2, pmax, cp, weights, st);
    cudaError_t e = cudaFreeAsync(scratch, st);
    if (rc == SG_OK && e != cudaSuccess) rc = (int)e;
    if (rc == SG_OK) {
        static thread_local char name[64];
        snprintf(name, sizeof(name), "%s_mixed", g_sg_last_variant);
        g_sg_last_variant = name;
    }
    return rc;
}

// ---- several derivative orders in one launch (2-D, uniform degree 1..3, not rational) ----------------------------
template <typename T, int P, int ND>
static int sg_launch_eval2d_multi(const SgMultiArgs<T> &m, const SgGridArgs<T> &a, const T *cp, cudaStream_t st)
{
    constexpr int V1 = sizeof(T) == 4 ? 4 : 2;
    const int threads = 128;
    const int64_t gx = (a.n_samples[0] + threads * V1 - 1) / (threads * V1);
    const int chunk = sg_pick_chunk(a.n_samples[1], gx * a.nout, 32, 512, sg_env_int("SG_CHUNK2D", 0));
    const int64_t gy = (a.n_samples[1] + chunk - 1) / chunk;
    if (gy > 65535 || a.nout > 65535) return SG_ERR_UNSUPPORTED;
    bool vec_ok = a.n_samples[0] % V1 == 0;
    for (int q = 0; q < ND; ++q) vec_ok = vec_ok && reinterpret_cast<uintptr_t>(m.eval[q]) % 16 == 0;
    const size_t smem = (size_t)chunk * (ND * (P + 1) * sizeof(T) + sizeof(int));
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)a.nout);
    sg_eval2d_multi_kernel<T, P, V1, ND><<<grid, threads, smem, st>>>(m, a, cp, chunk, vec_ok);
    g_sg_last_variant = "evaluate_multi2d";
    SG_AFTER_LAUNCH();
    return SG_OK;
}

template <typename T>
int sg_evaluate_multi_fast(const SgMultiArgs<T> &m, int n_der, const SgGridArgs<T> &a, const T *cp, cudaStream_t st)
{
    int p;
    if (a.nin != 2 || !sg_uniform_degree(a.degree, a.nin, p) || p < 1 || p > 3) return SG_ERR_UNSUPPORTED;
    if (n_der < 2 || n_der > SG_MULTI_MAX) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy == 1 || (g_sg_policy != 2 && a.n_total < 32768)) return SG_ERR_UNSUPPORTED;
    // evaluate! is bound by its output write, which fusing cannot reduce: on C2 (3 x 201 MB out) the fused kernel takes
    // 0.122 ms against 0.112 ms for three single launches (82 % of HBM each), so large outputs keep the single kernels and
    // the fused launch serves the sizes where launches and table/control-point re-reads matter (SG_EVAL_MULTI=2 forces it).
    const double out_bytes = (double)a.n_total * a.nout * n_der * sizeof(T);
    const int mode = sg_env_int("SG_EVAL_MULTI", 1);
    if (mode == 0 || (mode == 1 && g_sg_policy != 2 && out_bytes > 64.0e6)) return SG_ERR_UNSUPPORTED;
#define SG_MULTI_CASE(PP)                                                              \
    switch (n_der) {                                                                   \
        case 2: return sg_launch_eval2d_multi<T, PP, 2>(m, a, cp, st);                 \
        case 3: return sg_launch_eval2d_multi<T, PP, 3>(m, a, cp, st);                 \
        default: return sg_launch_eval2d_multi<T, PP, 4>(m, a, cp, st);                \
    }
    if (p == 1) { SG_MULTI_CASE(1) } else if (p == 2) { SG_MULTI_CASE(2) } else { SG_MULTI_CASE(3) }
#undef SG_MULTI_CASE
}
template int sg_evaluate_multi_fast<float>(const SgMultiArgs<float> &, int, const SgGridArgs<float> &, const float *, cudaStream_t);
template int sg_evaluate_multi_fast<double>(const SgMultiArgs<double> &, int, const SgGridArgs<double> &, const double *, cudaStream_t);

template int sg_evaluate_fast<float>(float *, const SgGridArgs<float> &, const float *, const float *, cudaStream_t);
template int sg_evaluate_fast<double>(double *, const SgGridArgs<double> &, const double *, const double *, cudaStream_t);
