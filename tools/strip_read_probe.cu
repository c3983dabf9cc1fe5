// strip_read_probe.cu -- how fast can a B200 READ a 2-D sample array in the access pattern of the 2-D adjoint march?
// (tools only; not part of the library).  Pattern "strip": CTA (x, c, r) marches rows [c*R, (c+1)*R) of channel r and reads,
// per row, the 128*V contiguous elements of column strip x (one 16-byte load per thread), U rows per batch -- exactly what
// sg_adj_march_kernel does, minus tables, spans and emission.  Pattern "flat": the same bytes, each CTA a contiguous range.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o strip_read_probe strip_read_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int U, int FMAS, bool PIPE, int MINB = 1>
__global__ void __launch_bounds__(128, MINB) strip_kernel(const float4 *__restrict__ X, float4 *__restrict__ out, int64_t inner4, int64_t n_d, int R,
                                                    const float *__restrict__ wts)
{
    __shared__ float ws[1024];
    for (int i = threadIdx.x; i < 1024; i += 128) ws[i] = wts[i];
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t j_lo = (int64_t)blockIdx.y * R, j_hi = min(n_d, j_lo + R);
    const float4 *p = X + q + inner4 * (j_lo + n_d * blockIdx.z);
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0, 0, 0, 0);
    auto consume = [&](const float4 &x, int64_t j) {
#pragma unroll
        for (int k = 0; k < FMAS; ++k) {
            const float w = ws[(j * 4 + (k & 3)) & 1023];
            acc[k & 3].x = fmaf(w, x.x, acc[k & 3].x); acc[k & 3].y = fmaf(w, x.y, acc[k & 3].y);
            acc[k & 3].z = fmaf(w, x.z, acc[k & 3].z); acc[k & 3].w = fmaf(w, x.w, acc[k & 3].w);
        }
    };
    if (!PIPE) {
        for (int64_t j = j_lo; j < j_hi; j += U) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + u < j_hi) x[u] = __ldcs(p + inner4 * u);
            p += inner4 * U;
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + u < j_hi) consume(x[u], j + u);
        }
    } else {
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (j_lo + u < j_hi) a[u] = __ldcs(p + inner4 * u);
        p += inner4 * U;
        for (int64_t j = j_lo; j < j_hi; j += 2 * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + U + u < j_hi) b[u] = __ldcs(p + inner4 * u);
            p += inner4 * U;
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + u < j_hi) consume(a[u], j + u);
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + 2 * U + u < j_hi) a[u] = __ldcs(p + inner4 * u);
            p += inner4 * U;
#pragma unroll
            for (int u = 0; u < U; ++u) if (j + U + u < j_hi) consume(b[u], j + U + u);
        }
    }
    float4 s = make_float4(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y,
                           acc[0].z + acc[1].z + acc[2].z + acc[3].z, acc[0].w + acc[1].w + acc[2].w + acc[3].w);
    out[q + inner4 * (blockIdx.y + (int64_t)gridDim.y * blockIdx.z)] = s;
}

// per-thread cp.async ring: D rows in flight per thread without holding them in registers; every thread reads back its OWN 16 bytes
template <int D, int FMAS, int MINB>
__global__ void __launch_bounds__(128, MINB) strip_cpasync_kernel(const float4 *__restrict__ X, float4 *__restrict__ out, int64_t inner4, int64_t n_d, int R,
                                                                  const float *__restrict__ wts)
{
    __shared__ float ws[1024];
    __shared__ float4 ring[D][128];
    for (int i = threadIdx.x; i < 1024; i += 128) ws[i] = wts[i];
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t j_lo = (int64_t)blockIdx.y * R, j_hi = min(n_d, j_lo + R);
    const float4 *p = X + q + inner4 * (j_lo + n_d * blockIdx.z);
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0, 0, 0, 0);
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(&ring[0][threadIdx.x]);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (j_lo + d < j_hi) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + d * 2048), "l"(p + inner4 * d) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    p += inner4 * D;
    for (int64_t j = j_lo; j < j_hi; ++j) {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        const int slot = (int)(j - j_lo) & (D - 1);
        const float4 x = ring[slot][threadIdx.x];
#pragma unroll
        for (int k = 0; k < FMAS; ++k) {
            const float w = ws[(j * 4 + (k & 3)) & 1023];
            acc[k & 3].x = fmaf(w, x.x, acc[k & 3].x); acc[k & 3].y = fmaf(w, x.y, acc[k & 3].y);
            acc[k & 3].z = fmaf(w, x.z, acc[k & 3].z); acc[k & 3].w = fmaf(w, x.w, acc[k & 3].w);
        }
        if (j + D < j_hi) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + slot * 2048), "l"(p) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        p += inner4;
    }
    float4 s = make_float4(acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y,
                           acc[0].z + acc[1].z + acc[2].z + acc[3].z, acc[0].w + acc[1].w + acc[2].w + acc[3].w);
    out[q + inner4 * (blockIdx.y + (int64_t)gridDim.y * blockIdx.z)] = s;
}

// same bytes, every CTA reads one contiguous range of R*128 float4 (U loads in flight per thread)
template <int U, int FMAS>
__global__ void __launch_bounds__(128) flat_kernel(const float4 *__restrict__ X, float4 *__restrict__ out, int64_t per_cta4, const float *__restrict__ wts)
{
    const int64_t cta = blockIdx.x + (int64_t)gridDim.x * (blockIdx.y + (int64_t)gridDim.y * blockIdx.z);
    const float4 *p = X + cta * per_cta4 + threadIdx.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int64_t i = 0; i < per_cta4; i += 128 * U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + 128 * u + threadIdx.x < per_cta4) x[u] = __ldcs(p + i + 128 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + 128 * u + threadIdx.x < per_cta4) {
#pragma unroll
            for (int k = 0; k < FMAS; ++k) { const float w = wts[(i + k) & 1023]; acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y); acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w); }
        }
    }
    out[cta * 128 + threadIdx.x] = acc;
}

// evict the array from L2 by READING another buffer (a memset would leave up to 126 MB of dirty lines whose write-back then
// competes with the timed kernel)
__global__ void flush_read_kernel(const float4 *__restrict__ f, size_t n4, float4 *out)
{
    float4 a = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { float4 x = f[i]; a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; }
    if (a.x == 123.f) out[0] = a;
}
template <typename F>
static float time_ms(F launch, int iters, char *flush, size_t flush_bytes)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> t;
    for (int it = 0; it < iters + 2; ++it) {
        flush_read_kernel<<<148 * 8, 256>>>(reinterpret_cast<const float4 *>(flush), flush_bytes / 16, reinterpret_cast<float4 *>(flush));
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 2) t.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char **argv)
{
    const int64_t n1 = argc > 1 ? atoll(argv[1]) : 4096, n2 = argc > 2 ? atoll(argv[2]) : 4096, nout = 3;
    const int64_t inner4 = n1 / 4;
    const size_t bytes = (size_t)n1 * n2 * nout * 4;
    float4 *X, *out; float *w; char *flush;
    const size_t flush_bytes = 512u << 20;
    CK(cudaMalloc(&X, bytes)); CK(cudaMemset(X, 0, bytes));
    CK(cudaMalloc(&out, 64u << 20)); CK(cudaMalloc(&w, 4096)); CK(cudaMemset(w, 0, 4096)); CK(cudaMalloc(&flush, flush_bytes)); CK(cudaMemset(flush, 0, flush_bytes)); CK(cudaDeviceSynchronize());
    const int strips = (int)(inner4 / 128);
    printf("array %lld x %lld x %lld float32 = %.1f MB, %d column strips of 2 KB\n", (long long)n1, (long long)n2, (long long)nout, bytes / 1e6, strips);
    for (int chunks : {24, 37, 48}) {
        const int R = (int)((n2 + chunks - 1) / chunks);
        dim3 grid(strips, chunks, (unsigned)nout);
#define RUN(name, kern) { int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0)); \
        float ms = time_ms([&] { kern<<<grid, 128>>>(X, out, inner4, n2, R, w); }, 9, flush, flush_bytes); \
        printf("{\"pattern\": \"%s\", \"chunks\": %d, \"ctas\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBs\": %.0f}\n", name, chunks, strips * chunks * (int)nout, occ, ms, bytes / ms / 1e6); }
        RUN("regs U=8 minb4 fma4", (strip_kernel<8, 4, false, 4>));
        RUN("regs U=8 minb5 fma4", (strip_kernel<8, 4, false, 5>));
        RUN("regs U=4 minb8 fma4", (strip_kernel<4, 4, false, 8>));
        RUN("regs U=16 minb3 fma4", (strip_kernel<16, 4, false, 3>));
        RUN("regs U=8 pipelined minb4 fma4", (strip_kernel<8, 4, true, 4>));
        RUN("regs U=8 minb4 fma8", (strip_kernel<8, 8, false, 4>));
        RUN("cpasync D=8 minb8 fma4", (strip_cpasync_kernel<8, 4, 8>));
        RUN("cpasync D=16 minb5 fma4", (strip_cpasync_kernel<16, 4, 5>));
        RUN("cpasync D=16 minb8 fma4", (strip_cpasync_kernel<16, 4, 8>));
        RUN("cpasync D=16 minb8 fma8", (strip_cpasync_kernel<16, 8, 8>));
        RUN("cpasync D=16 minb12 fma4", (strip_cpasync_kernel<16, 4, 12>));
    }
    return 0;
}
