// sg_fast_eval.cu -- dispatch of the tiled sm_100a fast paths for evaluate! (kernels: sg_fast_eval.cuh).
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

static bool sg_uniform_degree(const int *deg, int nin, int &p)
{
    p = deg[0];
    for (int d = 1; d < nin; ++d)
        if (deg[d] != p) return false;
    return true;
}

// number of marching chunks so that the grid has a few waves of CTAs on 148 SMs
static int sg_pick_chunk(int64_t n_march, int64_t col_tiles, int min_chunk, int max_chunk, int env_override)
{
    if (env_override > 0) return (int)std::min<int64_t>(std::max(env_override, 1), std::min<int64_t>(n_march, max_chunk));
    const int64_t target_ctas = 148 * 8;
    int64_t nchunks = (target_ctas + col_tiles - 1) / col_tiles;
    nchunks = std::max<int64_t>(1, std::min<int64_t>(nchunks, (n_march + min_chunk - 1) / min_chunk));
    int64_t chunk = (n_march + nchunks - 1) / nchunks;
    chunk = std::min<int64_t>(std::max<int64_t>(chunk, 1), max_chunk);
    return (int)chunk;
}

template <typename T, int P, int V1, int V2, int TY>
static int sg_launch_eval3d(T *eval, const SgGridArgs<T> &a, const T *cp, cudaStream_t st)
{
    const int64_t gx = (a.n_samples[0] + 32 * V1 - 1) / (32 * V1);
    const int64_t gy = (a.n_samples[1] + TY * V2 - 1) / (TY * V2);
    const int chunk = sg_pick_chunk(a.n_samples[2], gx * gy, 32, 1024, sg_env_int("SG_CHUNK3D", 0));
    const int64_t gz = (a.n_samples[2] + chunk - 1) / chunk;
    if (gy > 65535 || gz > 65535) return SG_ERR_UNSUPPORTED;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(eval) % 16 == 0) && (a.n_samples[0] % V1 == 0);
    const size_t smem = (size_t)chunk * ((P + 1) * sizeof(T) + sizeof(int));
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32, TY);
    sg_eval3d_march_kernel<T, P, V1, V2, TY><<<grid, block, smem, st>>>(eval, a, cp, chunk, a.nout, vec_ok);
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
int sg_evaluate_fast(T *eval, const SgGridArgs<T> &a, const T *cp, const T *weights, cudaStream_t st)
{
    int p;
    if (!sg_uniform_degree(a.degree, a.nin, p)) return SG_ERR_UNSUPPORTED;
    if (a.nin != 2 && a.nin != 3) return SG_ERR_UNSUPPORTED;
    if (p < 1 || p > 3) return SG_ERR_UNSUPPORTED;
    if (g_sg_policy != 2 && a.n_total < 32768) return SG_ERR_UNSUPPORTED;   // launch-latency regime: generic kernel
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

template int sg_evaluate_fast<float>(float *, const SgGridArgs<float> &, const float *, const float *, cudaStream_t);
template int sg_evaluate_fast<double>(double *, const SgGridArgs<double> &, const double *, const double *, cudaStream_t);
