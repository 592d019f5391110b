// Roofline denominators measured on the device itself: FP64 FMA throughput (dependent-free DFMA chains
// on every SM) and device-to-device copy bandwidth.  MEASURED_PEAKS.json holds the HBM and bf16 numbers
// of this pool but no FP64 figure, and FP64 is what bounds the pair and k-space
// kernels.
#include "context.hpp"

namespace lumol {

constexpr int PEAK_THREADS = 256;
constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(PEAK_THREADS) fp64_peak_kernel(double seed, double* __restrict__ out) {
    double v[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) v[k] = seed + (double)(threadIdx.x + k);
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; k++) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) s += v[k];
    if (s == 12345.678) out[0] = s;  // keeps the chains alive without a store in the common case
}

int measure_fp64_peak(Context* ctx, double* tflops) {
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve(16));
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t start, stop;
    LUMOL_CUDA_CHECK(ctx, cudaEventCreate(&start));
    LUMOL_CUDA_CHECK(ctx, cudaEventCreate(&stop));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(start, ctx->stream);
        for (int k = 0; k < 4; k++) {
            fp64_peak_kernel<<<blocks, PEAK_THREADS, 0, ctx->stream>>>(1.0 + rep, ctx->partials.ptr);
        }
        cudaEventRecord(stop, ctx->stream);
        cudaEventSynchronize(stop);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, start, stop);
        const double flops = 4.0 * (double)blocks * PEAK_THREADS * PEAK_CHAINS * (double)PEAK_ITERS * 2.0;
        const double t = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    ctx->launches += 24;
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    *tflops = best;
    return 0;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t count) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x) {
        dst[k] = src[k];
    }
}

int measure_copy_bandwidth(Context* ctx, double* gbs) {
    const size_t count = (size_t)64 << 20;  // 64 Mi double2 = 1 GiB read + 1 GiB written
    DeviceBuffer<double2> a, b;
    LUMOL_CUDA_CHECK(ctx, a.reserve(count));
    LUMOL_CUDA_CHECK(ctx, b.reserve(count));
    LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(a.ptr, 0, count * sizeof(double2), ctx->stream));
    cudaEvent_t start, stop;
    LUMOL_CUDA_CHECK(ctx, cudaEventCreate(&start));
    LUMOL_CUDA_CHECK(ctx, cudaEventCreate(&stop));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(start, ctx->stream);
        copy_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(a.ptr, b.ptr, count);
        cudaEventRecord(stop, ctx->stream);
        cudaEventSynchronize(stop);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, start, stop);
        const double t = 2.0 * (double)count * sizeof(double2) / (ms * 1e-3) / 1e9;
        if (rep > 0 && t > best) best = t;
    }
    ctx->launches += 6;
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    a.release();
    b.release();
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    *gbs = best;
    return 0;
}

}  // namespace lumol
