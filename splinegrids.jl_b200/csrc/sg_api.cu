// sg_api.cu -- library info, launch counter, and the NCCL gradient all-reduce entry points.
// NCCL is resolved at run time with dlopen("libnccl.so.2") so the library has no link-time NCCL
// dependency (a process that already loaded NCCL -- e.g. through torch -- shares that copy).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "sg_common.cuh"

std::atomic<int64_t> g_sg_launches{0};
int g_sg_policy = 0;
const char *g_sg_last_variant = "none";
thread_local const SgPushSpec *g_sg_push = nullptr;
thread_local bool g_sg_push_done = false;

extern "C" int sg_version(void) { return 100; /* 0.1.0 */ }

extern "C" int64_t sg_launch_count(void) { return g_sg_launches.load(); }
extern "C" void sg_launch_count_reset(void) { g_sg_launches.store(0); }
extern "C" void sg_set_kernel_policy(int policy) { g_sg_policy = policy; }
extern "C" const char *sg_last_variant(void) { return g_sg_last_variant; }

extern "C" const char *sg_status_string(int status)
{
    switch (status) {
        case SG_OK: return "SG_OK";
        case SG_ERR_INVALID_ARGUMENT: return "SG_ERR_INVALID_ARGUMENT";
        case SG_ERR_UNSUPPORTED: return "SG_ERR_UNSUPPORTED";
        case SG_ERR_WORKSPACE: return "SG_ERR_WORKSPACE";
        case SG_ERR_NCCL: return "SG_ERR_NCCL";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown sg_status";
}

// ---------------------------------------------------------------------------------------------
// NCCL (minimal ABI subset, stable since NCCL 2.x)
// ---------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct SgNccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

static SgNccl &sg_nccl()
{
    static SgNccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.handle, "ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.handle, "ncclCommInitRank"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce;
    });
    return n;
}

struct sg_comm {
    ncclComm_t comm;
    int world_size;
    int rank;
};

extern "C" int sg_comm_unique_id(void *unique_id_128_bytes)
{
    SG_CHECK_ARG(unique_id_128_bytes);
    SgNccl &n = sg_nccl();
    if (!n.ok) return SG_ERR_NCCL;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != 0) return SG_ERR_NCCL;
    std::memcpy(unique_id_128_bytes, &id, sizeof(id));
    return SG_OK;
}

extern "C" int sg_comm_create(sg_comm **comm, int world_size, int rank, const void *unique_id_128_bytes)
{
    SG_CHECK_ARG(comm && unique_id_128_bytes && world_size >= 1 && rank >= 0 && rank < world_size);
    SgNccl &n = sg_nccl();
    if (!n.ok) return SG_ERR_NCCL;
    ncclUniqueId id;
    std::memcpy(&id, unique_id_128_bytes, sizeof(id));
    ncclComm_t c;
    if (n.CommInitRank(&c, world_size, id, rank) != 0) return SG_ERR_NCCL;
    *comm = new sg_comm{c, world_size, rank};
    return SG_OK;
}

extern "C" int sg_comm_destroy(sg_comm *comm)
{
    if (!comm) return SG_OK;
    SgNccl &n = sg_nccl();
    int rc = (n.ok && n.CommDestroy(comm->comm) == 0) ? SG_OK : SG_ERR_NCCL;
    delete comm;
    return rc;
}

static int sg_allreduce(void *buf, int64_t count, int dtype, sg_comm *comm, void *stream)
{
    SG_NVTX("sg_allreduce_sum");
    SG_CHECK_ARG(buf && comm && count >= 0);
    SgNccl &n = sg_nccl();
    if (!n.ok) return SG_ERR_NCCL;
    if (count == 0) return SG_OK;
    return n.AllReduce(buf, buf, (size_t)count, dtype, ncclSum, comm->comm, sg_stream(stream)) == 0 ? SG_OK : SG_ERR_NCCL;
}

extern "C" int sg_allreduce_sum_f32(float *buf, int64_t count, sg_comm *comm, void *stream)
{
    return sg_allreduce(buf, count, ncclFloat32, comm, stream);
}

extern "C" int sg_allreduce_sum_f64(double *buf, int64_t count, sg_comm *comm, void *stream)
{
    return sg_allreduce(buf, count, ncclFloat64, comm, stream);
}
