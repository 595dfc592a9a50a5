// sitator_b200 -- pipe micro-benchmarks used by bench.py for the roofline denominators that
// MEASURED_PEAKS.json does not carry (FP32 lane-ops, FP64 lane-ops, SFU ops per second).
#include "../../include/sitator_b200.h"
#include <cuda_runtime.h>

namespace sitb {
int set_error(int code, const char* fmt, ...);

template <typename T>
__global__ void k_fma_chain(T* out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_sfu_chain(float* out, int iters) {
    float x0 = 1.0f + threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
}  // namespace sitb

using namespace sitb;

// rates in operations (one FMA / one SFU op = 1) per second over the whole device
extern "C" int sitb_microbench(int device, double* fp32_ops, double* fp64_ops, double* sfu_ops) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    void* buf = nullptr;
    e = cudaMalloc(&buf, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    const double lanes = (double)blocks * threads;
    for (int rep = 0; rep < 2; ++rep) {
        const int it32 = 4096, it64 = 1024, its = 2048;
        cudaEventRecord(a);
        k_fma_chain<float><<<blocks, threads>>>((float*)buf, it32, 1.000001f, 1e-7f);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        if (fp32_ops) *fp32_ops = lanes * it32 * 64.0 / (ms * 1e-3);
        cudaEventRecord(a);
        k_fma_chain<double><<<blocks, threads>>>((double*)buf, it64, 1.000001, 1e-7);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        if (fp64_ops) *fp64_ops = lanes * it64 * 64.0 / (ms * 1e-3);
        cudaEventRecord(a);
        k_sfu_chain<<<blocks, threads>>>((float*)buf, its);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        if (sfu_ops) *sfu_ops = lanes * its * 32.0 / (ms * 1e-3);
    }
    e = cudaGetLastError();
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf);
    if (e != cudaSuccess) return set_error(SITB_E_CUDA, "sitb_microbench: %s", cudaGetErrorString(e));
    return SITB_OK;
}
