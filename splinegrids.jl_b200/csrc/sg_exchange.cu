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
    SG_NVTX("sg_exchange_push");
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
    SG_NVTX("sg_exchange_reduce");
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

// ---------------------------------------------------------------------------------------------
// Flag-synchronised exchange: the barrier between push and reduce lives in the C ABI (no host framework needed).
//
//   flags  (peer-mapped, `world` uint64 per rank, zero-initialised): flags[r] on rank q = number of exchanges whose
//          push from rank r has completely landed in q's staging buffer.  Monotone counters, never reset.
//   sync   (local device memory, SG_EXCHANGE_SYNC_BYTES, zero-initialised): { epoch = exchanges completed by this rank,
//          done = block counter of the running reduce, timed_out }.  The epoch lives on the device, so a CUDA graph
//          that contains the exchange can be replayed: every launch reads epoch + 1 as "its" exchange number.
//
// sg_exchange_signal:      (stream-ordered after the push) fence.sys + release-store of epoch + 1 into flags[my_rank] of
//                          every rank.
// sg_exchange_wait_reduce: [optionally the same signal first, by block 0] every block acquire-spins until all `world`
//                          flags of this rank have reached epoch + 1, then reduces its part of the gradient exactly like
//                          sg_exchange_reduce; the last block to finish publishes epoch + 1.  Two staging buffers must
//                          alternate between consecutive exchanges (a rank can start pushing exchange e + 1 while a
//                          peer is still reducing exchange e).
// A peer that never signals would hang the GPU: the spin gives up after SG_EXCHANGE_TIMEOUT_NS, raises sync->timed_out
// (sg_exchange_status) and lets the kernel finish.
// ---------------------------------------------------------------------------------------------
#define SG_EXCHANGE_TIMEOUT_NS 20000000000ull   // 20 s

struct SgExchangeSync {
    unsigned long long epoch;
    unsigned int done;
    unsigned int timed_out;
    unsigned long long pad[6];
};
static_assert(sizeof(SgExchangeSync) == 64, "SG_EXCHANGE_SYNC_BYTES");

struct SgFlagPtrs {
    unsigned long long *flags[SG_MAX_PEERS];
};

__device__ __forceinline__ void sg_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sg_ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long sg_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// thread r < world of ONE block: tell rank r that this rank's push number `e` has landed
__device__ __forceinline__ void sg_exchange_signal_peers(const SgFlagPtrs &pf, int world, int my_rank, unsigned long long e,
                                                         unsigned peer_mask = 0xffffffffu)
{
    if ((int)threadIdx.x < world && ((peer_mask >> threadIdx.x) & 1u)) {
        __threadfence_system();                                      // the pushes of the preceding kernels, cumulatively
        sg_st_release_sys(pf.flags[threadIdx.x] + my_rank, e);
    }
}

__global__ void sg_exchange_signal_kernel(const __grid_constant__ SgFlagPtrs pf, int world, int my_rank, const SgExchangeSync *sync)
{
    sg_exchange_signal_peers(pf, world, my_rank, sync->epoch + 1);
}

// thread r < world of every block waits for flags[r] >= e; returns after a block barrier (all threads may read the stage)
__device__ __forceinline__ void sg_exchange_wait_all(const unsigned long long *my_flags, int world, unsigned long long e, SgExchangeSync *sync,
                                                     unsigned peer_mask = 0xffffffffu)
{
    if ((int)threadIdx.x < world && ((peer_mask >> threadIdx.x) & 1u)) {
        const unsigned long long t0 = sg_globaltimer();
        while (sg_ld_acquire_sys(my_flags + threadIdx.x) < e) {
            __nanosleep(64);
            if (sg_globaltimer() - t0 > SG_EXCHANGE_TIMEOUT_NS) { sync->timed_out = 1; break; }
        }
    }
    __syncthreads();
}

// A block owns one control plane k (blockIdx.y) of one output (blockIdx.z) and strides over its 16-byte vectors; the
// ranks whose support covers the plane (block-uniform, <= 2 for slabs thicker than p spans) are found once per block.
template <typename T, int V>
__global__ void __launch_bounds__(256) sg_exchange_wait_reduce_kernel(T *__restrict__ grad, const T *stage, const unsigned long long *my_flags,
                                                                      SgExchangeSync *sync, const __grid_constant__ SgFlagPtrs pf, int do_signal,
                                                                      int my_rank, const __grid_constant__ SgSupports sup, int world,
                                                                      int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes,
                                                                      unsigned peer_mask, int64_t k_first)
{
    // peer_mask / k_first: the support-plane exchange signals and waits for the ranks whose supports meet this rank's only,
    // and reduces planes k_first .. k_first + gridDim.y - 1 (this rank's support); the replicated exchange passes ~0 / 0
    const unsigned long long e = sync->epoch + 1;                    // nobody writes epoch while blocks of this launch can still read it
    if (do_signal && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) sg_exchange_signal_peers(pf, world, my_rank, e, peer_mask);
    sg_exchange_wait_all(my_flags, world, e, sync, peer_mask);

    const int64_t k = k_first + blockIdx.y;
    const int o = blockIdx.z;
    T *__restrict__ dst = grad + plane_elems * (k + c_last * o);
    const int64_t n_vec = plane_elems / V;
    const T *src[SG_MAX_PEERS];
    int ns = 0;
    for (int r = 0; r < world; ++r) {
        const int64_t l = k - sup.k0[r];
        if (l >= 0 && l < sup.np[r]) src[ns++] = stage + plane_elems * (l + max_planes * (o + (int64_t)nout * r));
    }
    using VecT = typename std::conditional<sizeof(T) * V == 16, double2, typename std::conditional<sizeof(T) * V == 8, double, float>::type>::type;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = T(0);
        if (ns == 1) {                                               // interior plane: a copy
            const VecT x = __ldcg(reinterpret_cast<const VecT *>(src[0] + i * V));
            const T *xv = reinterpret_cast<const T *>(&x);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = xv[v];
        } else if (ns == 2) {                                        // halo plane shared by two neighbouring slabs
            const VecT x = __ldcg(reinterpret_cast<const VecT *>(src[0] + i * V));
            const VecT y = __ldcg(reinterpret_cast<const VecT *>(src[1] + i * V));
            const T *xv = reinterpret_cast<const T *>(&x), *yv = reinterpret_cast<const T *>(&y);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = xv[v] + yv[v];
        } else {
            for (int q = 0; q < ns; ++q) {                           // rank order: deterministic
                const VecT x = __ldcg(reinterpret_cast<const VecT *>(src[q] + i * V));
                const T *xv = reinterpret_cast<const T *>(&x);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += xv[v];
            }
        }
        *reinterpret_cast<VecT *>(dst + i * V) = *reinterpret_cast<const VecT *>(acc);
    }
    // the last block to finish publishes the new epoch (all blocks have read the old one before they got here)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        __threadfence();
        if (atomicAdd(&sync->done, 1u) == total - 1) {
            sync->done = 0;
            __threadfence();
            sync->epoch = e;
        }
    }
}

static int sg_fill_flag_ptrs(SgFlagPtrs &pf, void *const *peer_flags, int world)
{
    for (int r = 0; r < world; ++r) {
        SG_CHECK_ARG(peer_flags[r] && reinterpret_cast<uintptr_t>(peer_flags[r]) % 8 == 0);
        pf.flags[r] = static_cast<unsigned long long *>(peer_flags[r]);
    }
    return SG_OK;
}

extern "C" int sg_exchange_signal(void *const *peer_flags, int world, int my_rank, const void *local_sync, void *stream)
{
    SG_NVTX("sg_exchange_signal");
    SG_CHECK_ARG(peer_flags && local_sync && world >= 1 && my_rank >= 0 && my_rank < world);
    if (world > SG_MAX_PEERS) return SG_ERR_UNSUPPORTED;
    SgFlagPtrs pf{};
    int rc = sg_fill_flag_ptrs(pf, peer_flags, world);
    if (rc != SG_OK) return rc;
    sg_exchange_signal_kernel<<<1, 32, 0, sg_stream(stream)>>>(pf, world, my_rank, static_cast<const SgExchangeSync *>(local_sync));
    SG_AFTER_LAUNCH();
    return SG_OK;
}

extern "C" int sg_exchange_status(const void *local_sync, unsigned long long *epoch, int *timed_out, void *stream)
{
    SG_CHECK_ARG(local_sync);
    SgExchangeSync h{};
    SG_CUDA(cudaMemcpyAsync(&h, local_sync, sizeof(h), cudaMemcpyDeviceToHost, sg_stream(stream)));
    SG_CUDA(cudaStreamSynchronize(sg_stream(stream)));
    if (epoch) *epoch = h.epoch;
    if (timed_out) *timed_out = (int)h.timed_out;
    return SG_OK;
}

template <typename T>
static int sg_exchange_wait_reduce_impl(T *grad, const T *stage, const void *my_flags, void *local_sync, void *const *peer_flags_or_null,
                                        int world, int my_rank, const int64_t *k0s, const int64_t *nps, int64_t plane_elems,
                                        int64_t c_last, int nout, int64_t max_planes, void *stream, bool support_only = false)
{
    SG_NVTX("sg_exchange_wait_reduce");
    SG_CHECK_ARG(grad && stage && my_flags && local_sync && k0s && nps && world >= 1 && plane_elems >= 1 && c_last >= 1 && nout >= 1);
    SG_CHECK_ARG(my_rank >= 0 && my_rank < world && reinterpret_cast<uintptr_t>(my_flags) % 8 == 0);
    if (world > SG_MAX_PEERS || c_last > 65535 || nout > 65535) return SG_ERR_UNSUPPORTED;
    SgSupports sup{};
    for (int r = 0; r < world; ++r) { sup.k0[r] = k0s[r]; sup.np[r] = nps[r]; }
    SgFlagPtrs pf{};
    if (peer_flags_or_null) {
        int rc = sg_fill_flag_ptrs(pf, peer_flags_or_null, world);
        if (rc != SG_OK) return rc;
    }
    constexpr int VV = 16 / sizeof(T);
    const bool vec = (plane_elems % VV == 0) && (reinterpret_cast<uintptr_t>(grad) % 16 == 0) && (reinterpret_cast<uintptr_t>(stage) % 16 == 0);
    // ~4 vectors per thread; the whole grid is resident at once for the usual sizes (the spin needs no particular order:
    // a block only waits for PEERS, and every peer's signal is issued by the first block of its own launch)
    const int64_t nv = plane_elems / (vec ? VV : 1);
    unsigned mask = 0xffffffffu;
    int64_t k_first = 0, n_planes = c_last;
    if (support_only) {                                              // ranks whose supports overlap this rank's (and this rank)
        SG_CHECK_ARG(k0s[my_rank] >= 0 && nps[my_rank] >= 1 && k0s[my_rank] + nps[my_rank] <= c_last);
        mask = 0;
        for (int r = 0; r < world; ++r)
            if (r == my_rank || (k0s[r] < k0s[my_rank] + nps[my_rank] && k0s[my_rank] < k0s[r] + nps[r])) mask |= 1u << r;
        k_first = k0s[my_rank]; n_planes = nps[my_rank];
    }
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((nv + 1023) / 1024, 16)), (unsigned)n_planes, (unsigned)nout);
    auto *sy = static_cast<SgExchangeSync *>(local_sync);
    auto *fl = static_cast<const unsigned long long *>(my_flags);
    if (vec)
        sg_exchange_wait_reduce_kernel<T, VV><<<grid, 256, 0, sg_stream(stream)>>>(grad, stage, fl, sy, pf, peer_flags_or_null ? 1 : 0, my_rank, sup, world, plane_elems, c_last, nout, max_planes, mask, k_first);
    else
        sg_exchange_wait_reduce_kernel<T, 1><<<grid, 256, 0, sg_stream(stream)>>>(grad, stage, fl, sy, pf, peer_flags_or_null ? 1 : 0, my_rank, sup, world, plane_elems, c_last, nout, max_planes, mask, k_first);
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
    extern "C" int sg_exchange_wait_reduce_##SUF(T *grad, const T *stage, const void *my_flags, void *local_sync,               \
                                                 void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s,   \
                                                 const int64_t *nps, int64_t plane_elems, int64_t c_last, int nout,             \
                                                 int64_t max_planes, void *stream)                                             \
    {                                                                                                                           \
        return sg_exchange_wait_reduce_impl<T>(grad, stage, my_flags, local_sync, peer_flags_or_null, world, my_rank, k0s, nps, \
                                               plane_elems, c_last, nout, max_planes, stream);                                  \
    }                                                                                                                           \
    extern "C" int sg_exchange_wait_reduce_support_##SUF(T *grad, const T *stage, const void *my_flags, void *local_sync,       \
                                                 void *const *peer_flags_or_null, int world, int my_rank, const int64_t *k0s,   \
                                                 const int64_t *nps, int64_t plane_elems, int64_t c_last, int nout,             \
                                                 int64_t max_planes, void *stream)                                             \
    {                                                                                                                           \
        return sg_exchange_wait_reduce_impl<T>(grad, stage, my_flags, local_sync, peer_flags_or_null, world, my_rank, k0s, nps, \
                                               plane_elems, c_last, nout, max_planes, stream, true);                            \
    }                                                                                                                           \
    extern "C" int sg_exchange_reduce_##SUF(T *grad, const T *stage, int world, const int64_t *k0s, const int64_t *nps,        \
                                            int64_t plane_elems, int64_t c_last, int nout, int64_t max_planes, void *stream)   \
    {                                                                                                                           \
        return sg_exchange_reduce_impl<T>(grad, stage, world, k0s, nps, plane_elems, c_last, nout, max_planes, stream);         \
    }

SG_DEFINE_EXCHANGE_API(float, f32)
SG_DEFINE_EXCHANGE_API(double, f64)
