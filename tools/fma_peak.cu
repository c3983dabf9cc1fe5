// fma_peak.cu -- register-resident FMA micro-benchmark: measures the FP32 / FP64 CUDA-core FMA peak of
// the device (the roofline's compute denominator; MEASURED_PEAKS.json records no such number).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak tools/fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T *out, T a, T b, int iters)
{
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == T(-1.2345)) out[0] = s;   // never true; keeps the loop alive
}

template <typename T>
double run(const char *name, int iters)
{
    constexpr int ILP = 16;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    T *out;
    cudaMalloc(&out, sizeof(T));
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<T, ILP><<<grid, block>>>(out, T(1.0000001), T(1e-9), iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double fmas = (double)grid.x * block.x * ILP * (double)iters;
        double rate = fmas / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    printf("{\"what\": \"%s\", \"fma_per_s\": %.4e, \"tflops\": %.2f, \"sms\": %d}\n", name, best, 2 * best / 1e12, sms);
    cudaFree(out);
    return best;
}

int main()
{
    run<float>("fp32_fma_peak", 20000);
    run<double>("fp64_fma_peak", 4000);
    return 0;
}
