// sg_common.cuh -- shared definitions for the sm_100a kernels of libsplinegrids_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "splinegrids_b200.h"

// NVTX range around every exported call (tracing): shows up in Nsight Systems / ncu --nvtx timelines
// under the domain-less name of the entry point; a no-op (one predictable branch) when no tool is attached.
struct SgNvtxRange {
    explicit SgNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~SgNvtxRange() { nvtxRangePop(); }
    SgNvtxRange(const SgNvtxRange &) = delete;
    SgNvtxRange &operator=(const SgNvtxRange &) = delete;
};
#define SG_NVTX(name) SgNvtxRange sg_nvtx_range_(name)

#define SG_MAXW (SG_MAX_DEGREE + 1)
#define SG_GEN_OCHUNK 4   // outputs accumulated in registers per pass over the window (generic kernels)

// Process-wide count of kernel launches made by this library (sg_launch_count()).
extern std::atomic<int64_t> g_sg_launches;
extern int g_sg_policy;            // 0 auto, 1 generic only, 2 prefer fast paths
extern const char *g_sg_last_variant;

#define SG_CHECK_ARG(cond)                          \
    do {                                            \
        if (!(cond)) return SG_ERR_INVALID_ARGUMENT; \
    } while (0)

#define SG_CUDA(call)                           \
    do {                                        \
        cudaError_t e__ = (call);               \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

// Count the launch and surface launch-configuration errors without synchronising.
#define SG_AFTER_LAUNCH()                        \
    do {                                         \
        g_sg_launches.fetch_add(1);              \
        cudaError_t e__ = cudaPeekAtLastError(); \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

static inline cudaStream_t sg_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline unsigned sg_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Grid description passed BY VALUE to the evaluation kernels (no H2D copy, no allocation).
template <typename T>
struct SgGridArgs {
    int nin;
    int nout;
    int64_t n_samples[SG_MAX_DIMS];
    int64_t n_cp[SG_MAX_DIMS];
    int64_t cp_stride[SG_MAX_DIMS];   // element stride of control-point dim d
    const T *table[SG_MAX_DIMS];      // already offset to the selected derivative slice: (n_d, p_d+1)
    const int32_t *index[SG_MAX_DIMS];
    int degree[SG_MAX_DIMS];
    int64_t n_total;                  // prod n_samples
    int64_t cp_total;                 // prod n_cp
    int64_t n_window;                 // prod (p_d + 1)
};

template <typename T>
static inline int sg_fill_grid_args(SgGridArgs<T> &a, int nin, const int64_t *n_samples, const int64_t *n_cp,
                                    int nout, const T *const *tables, const int32_t *const *indices,
                                    const int *degree, const int *mdo, const int *der)
{
    SG_CHECK_ARG(n_samples && n_cp && tables && indices && degree && mdo && der);
    if (nin < 1 || nin > SG_MAX_DIMS) return SG_ERR_UNSUPPORTED;
    SG_CHECK_ARG(nout >= 1);
    a.nin = nin;
    a.nout = nout;
    a.n_total = 1;
    a.cp_total = 1;
    a.n_window = 1;
    for (int d = 0; d < SG_MAX_DIMS; ++d) {
        a.n_samples[d] = 1; a.n_cp[d] = 1; a.cp_stride[d] = 0; a.table[d] = nullptr; a.index[d] = nullptr; a.degree[d] = 0;
    }
    for (int d = 0; d < nin; ++d) {
        SG_CHECK_ARG(n_samples[d] >= 1 && n_cp[d] >= 1 && tables[d] && indices[d]);
        if (degree[d] < 0 || degree[d] > SG_MAX_DEGREE) return SG_ERR_UNSUPPORTED;
        SG_CHECK_ARG(n_cp[d] >= degree[d] + 1);
        SG_CHECK_ARG(mdo[d] >= 0 && der[d] >= 0 && der[d] <= mdo[d]);
        a.n_samples[d] = n_samples[d];
        a.n_cp[d] = n_cp[d];
        a.cp_stride[d] = a.cp_total;
        a.table[d] = tables[d] + (int64_t)der[d] * (degree[d] + 1) * n_samples[d];
        a.index[d] = indices[d];
        a.degree[d] = degree[d];
        a.n_total *= n_samples[d];
        a.cp_total *= n_cp[d];
        a.n_window *= degree[d] + 1;
    }
    return SG_OK;
}

// Fused gradient push of a slab-sharded adjoint (sg_evaluate_adjoint_push_*, sg_exchange.cu): the last kernel of the
// adjoint stores the control planes of this rank's support straight into this rank's slot of EVERY peer's staging buffer
// (peer-to-peer stores over NVLink), so the transfer overlaps the computation and no separate push kernel runs.
#define SG_MAX_PEERS 16
struct SgPushSpec {
    void *stage[SG_MAX_PEERS];   // staging buffer of every rank (peer-mapped), layout [world][nout][max_planes][plane_elems]
    int world, my_rank;          // world == 0: no push
    int n_dst;                   // destinations the kernel stores to: `world` peer pointers, or 1 = stage[0] is a MULTICAST address
                                 // (NVLS: one store, the NVSwitch replicates it into every rank's buffer -- 1/world of the egress)
    long long max_planes;
    int dst_lo[SG_MAX_PEERS];    // destination r receives the LOCAL planes [dst_lo[r], dst_hi[r]) of this rank's support: everything
    int dst_hi[SG_MAX_PEERS];    // (0, max_planes) in the replicated exchange; only the planes rank r's own slab touches in the
                                 // support-plane exchange (sg_evaluate_adjoint_planned_support_*)
    int keep_local;              // 0: the caller does not need the local partial gradient (the reduce overwrites it): the fused
                                 // pipeline then skips its own writes of the control-point array (zeros and results)
};
extern thread_local const SgPushSpec *g_sg_push;   // set around sg_evaluate_adjoint_impl by the push entry point
extern thread_local bool g_sg_push_done;           // the pipeline that ran did the push itself

// Adjoint workspace header (first 256 bytes of the workspace)
struct SgAdjointHeader {
    int nonmonotone;  // set to 1 by the prep kernel if any dimension's span indices decrease
    int m2_skipped;   // set by the TMA-fed double march when it leaves a tile to the register kernel (complement pass)
    int span_first[SG_MAX_DIMS];   // span of the first / last sample of every dimension (1-based), written by the
    int span_last[SG_MAX_DIMS];    // prep kernel: the control indices a (slab of a) grid can touch are [first-p, last]
    int m2g_bad;      // set by the prep kernel when a column block's table of dimension 1 does not fit (fused double march)
    int rows2_max;    // largest number of samples in one knot span of dimension 2 (prep kernel; 3-D grids)
    int pad[60 - 2 * SG_MAX_DIMS];
};
static_assert(sizeof(SgAdjointHeader) == 256, "adjoint workspace header is 256 bytes");
// Fused double march (sg_adjoint_march2g.cuh): per column block of 128 samples of dimension 1, the control indices it
// touches [i1_lo, i1_lo + ni) (1-based) and the longest support (in samples of the block) of one of them.
struct SgM2gBlockHdr {
    int i1_lo, ni, rm, pad;
};
// Fused adjoint: control-index slots reserved per warp tile of TS = 32*V samples of dimension 1.  A tile
// touches at most TS + p control indices (every sample in its own span), p <= 5 -> TS + 8 (keeps 16-byte alignment).
#define SG_ADJ_SLOT_PAD 8

// Read-only (non-coherent) load helper
template <typename T>
__device__ __forceinline__ T sg_ldg(const T *p) { return __ldg(p); }

// rounding-exact (non-contracted) arithmetic, so K2 reproduces an IEEE evaluation of the reference
__device__ __forceinline__ float sg_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double sg_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float sg_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sg_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sg_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sg_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sg_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sg_div(double a, double b) { return __ddiv_rn(a, b); }
