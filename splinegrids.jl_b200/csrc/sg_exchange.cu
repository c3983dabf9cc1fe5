// sg_exchange.cu -- gradient exchange for slab-sharded multi-GPU adjoints over NVLink peer memory.
//
// A rank's partial control-point gradient is non-zero only on the control planes its slab touches
// (planes [k0, k0+np) of the slowest control axis; neighbouring slabs overlap in p planes).  Instead of an
// all-reduce of the whole (mostly zero) gradient, every rank PUSHES its support planes straight into its slot
// of every peer's staging buffer with peer-to-peer stores (sg_exchange_push), the ranks meet at a barrier
// (provided by the host framework on the same stream), and each rank sums the <= few slots covering every
// plane locally (sg_exchange_reduce).  Deterministic: slots are summed in rank order.
// Staging layout per rank: [world][nout][max_planes][plane_elems].
#include <algorithm>
#include <type_traits>

#include "sg_common.cuh"

template <typename T>
struct SgPeerPtrs {
    T *stage[SG_MAX_PEERS];
};
struct SgSupports {
    int64_t k0[SG_MAX_PEERS];
    int64_t np[SG_MAX_PEERS];
};

// grad: (plane_elems, c_last, nout) column-major.  One thread moves 16 bytes to every peer.
template <typename T, int V>
__global__ void __launch_bounds__(256) sg_exchange_push_kernel(const T *__restrict__ grad, const __grid_constant__ SgPeerPtrs<T> peers,
                                                               int world, int my_rank, int64_t plane_elems, int64_t c_last, int nout,
                                                               int64_t k0, int64_t np, int64_t max_planes)
{
    const int64_t per_out = np * plane_elems;                        // contiguous in grad for a fixed output
    const int64_t n_vec = per_out / V;
    const int o = blockIdx.y;
    const T *__restrict__ src = grad + plane_elems * (k0 + c_last * o);
    const int64_t slot_off = (int64_t)max_planes * plane_elems * (o + (int64_t)nout * my_rank);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        typename std::conditional<sizeof(T) * V == 16, uint4, typename std::conditional<sizeof(T) * V == 8, uint2, unsigned>::type>::type v =
            *reinterpret_cast<const decltype(v) *>(src + i * V);
#pragma unroll 1
        for (int r = 0; r < world; ++r) *reinterpret_cast<decltype(v) *>(peers.stage[r] + slot_off + i * V) = v;
    }
}

// grad[:, k, o] = sum over ranks r with k0_r <= k < k0_r + np_r of stage[r][o][k - k0_r][:]
template <typename T, int V>
__global__ void __launch_bounds__(256) sg_exchange_reduce_kernel(T *__restrict__ grad, const T *__restrict__ stage,
                                                                 const __grid_constant__ SgSupports sup, int world, int64_t plane_elems,
                                                                 int64_t c_last, int nout, int64_t max_planes)
{
    const int64_t k = blockIdx.y;                                     // control plane of the slowest axis
    const int o = blockIdx.z;
    T *__restrict__ dst = grad + plane_elems * (k + c_last * o);
    const int64_t n_vec = plane_elems / V;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = T(0);
        for (int r = 0; r < world; ++r) {
            const int64_t l = k - sup.k0[r];
            if (l < 0 || l >= sup.np[r]) continue;
            const T *__restrict__ s = stage + plane_elems * (l + max_planes * (o + (int64_t)nout * r)) + i * V;
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] += s[v];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) dst[i * V + v] = acc[v];
    }
}

template <typename T>
static int sg_exchange_push_impl(const T *grad, void *const *peer_stage, int world, int my_rank, int64_t plane_elems,
                                 int64_t c_last, int nout, int64_t k0, int64_t np, int64_t max_planes, void *stream)
{
    SG_CHECK_ARG(grad && peer_stage && world >= 1 && my_rank >= 0 && my_rank < world && plane_elems >= 1 && nout >= 1);
    SG_CHECK_ARG(k0 >= 0 && np >= 0 && k0 + np <= c_last && np <= max_planes);
    if (world > SG_MAX_PEERS) return SG_ERR_UNSUPPORTED;
    if (np == 0) return SG_OK;
    SgPeerPtrs<T> p{};
    for (int r = 0; r < world; ++r) { SG_CHECK_ARG(peer_stage[r]); p.stage[r] = static_cast<T *>(peer_stage[r]); }
    constexpr int VV = 16 / sizeof(T);
    const bool vec = (plane_elems % VV == 0) && (reinterpret_cast<uintptr_t>(grad) % 16 == 0);
    const int64_t n = np * plane_elems;
    dim3 grid((unsigned)std::min<int64_t>((n / (vec ? VV : 1) + 255) / 256, 148 * 8), (unsigned)nout);
    if (vec)
        sg_exchange_push_kernel<T, VV><<<grid, 256, 0, sg_stream(stream)>>>(grad, p, world, my_rank, plane_elems, c_last, nout, k0, np, max_planes);
    else
        sg_exchange_push_kernel<T, 1><<<grid, 256, 0, sg_stream(stream)>>>(grad, p, world, my_rank, plane_elems, c_last, nout, k0, np, max_planes);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

template <typename T>
static int sg_exchange_reduce_impl(T *grad, const T *stage, int world, const int64_t *k0s, const int64_t *nps, int64_t plane_elems,
                                   int64_t c_last, int nout, int64_t max_planes, void *stream)
{
    SG_CHECK_ARG(grad && stage && k0s && nps && world >= 1 && plane_elems >= 1 && c_last >= 1 && nout >= 1);
    if (world > SG_MAX_PEERS || c_last > 65535 || nout > 65535) return SG_ERR_UNSUPPORTED;
    SgSupports sup{};
    for (int r = 0; r < world; ++r) { sup.k0[r] = k0s[r]; sup.np[r] = nps[r]; }
    constexpr int VV = 16 / sizeof(T);
    const bool vec = (plane_elems % VV == 0) && (reinterpret_cast<uintptr_t>(grad) % 16 == 0) && (reinterpret_cast<uintptr_t>(stage) % 16 == 0);
    dim3 grid((unsigned)std::min<int64_t>((plane_elems / (vec ? VV : 1) + 255) / 256, 64), (unsigned)c_last, (unsigned)nout);
    if (vec)
        sg_exchange_reduce_kernel<T, VV><<<grid, 256, 0, sg_stream(stream)>>>(grad, stage, sup, world, plane_elems, c_last, nout, max_planes);
    else
        sg_exchange_reduce_kernel<T, 1><<<grid, 256, 0, sg_stream(stream)>>>(grad, stage, sup, world, plane_elems, c_last, nout, max_planes);
    SG_AFTER_LAUNCH();
    return SG_OK;
}

#define SG_DEFINE_EXCHANGE_API(T, SUF)                                                                                          \
    extern "C" int sg_exchange_push_##SUF(const T *grad, void *const *peer_stage, int world, int my_rank, int64_t plane_elems, \
                                          int64_t c_last, int nout, int64_t k0, int64_t np, int64_t max_planes, void *stream)  \
    {                                                                                                                           \
        return sg_exchange_push_impl<T>(grad, peer_stage, world, my_rank, plane_elems, c_last, nout, k0, np, max_planes,        \
                                        stream);                                                                                \
    }                                                                                                                           \
    extern "C" int sg_exchange_reduce_##SUF(T *grad, const T *stage, int world, const int64_t *k0s, const int64_t *nps,        \
                                            int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream)   \
    {                                                                                                                           \
        return sg_exchange_reduce_impl<T>(grad, stage, world, k0s, nps, plane_elems, c_last, nout, max_planes, stream);         \
    }

SG_DEFINE_EXCHANGE_API(float, f32)
SG_DEFINE_EXCHANGE_API(double, f64)
