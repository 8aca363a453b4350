// FP64 FMA peak microbenchmark (the roofline denominator for the lean ELBO kernel, which is
// FP64-pipe bound; MEASURED_PEAKS.json has no FP64 figure).  8 independent DFMA chains per
// thread, 2 CTAs of 256 threads per SM, timed with CUDA events.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) pfb_dfma_chain(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// returns TFLOP/s (2 flops per FMA) in *tflops; best of `reps`.
extern "C" int pfb_measure_fp64_fma_tflops(int device, int reps, double* tflops) {
    if (!tflops) return -1;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int)e;
    const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 4096;
    double* d = nullptr;
    e = cudaMalloc(&d, (size_t)blocks * threads * 8);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(a);
        pfb_dfma_chain<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return e == cudaSuccess ? 0 : (int)e;
}

// FP64 tensor-core peak (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4): 8 independent accumulator pairs per
// warp, 2 CTAs of 256 threads per SM.  256 FMAs per warp-level instruction.
__global__ void __launch_bounds__(256) pfb_dmma_chain(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int pfb_measure_fp64_dmma_tflops(int device, int reps, double* tflops) {
    if (!tflops) return -1;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int)e;
    const int blocks = prop.multiProcessorCount * 2, threads = 256, iters = 4096;
    double* d = nullptr;
    e = cudaMalloc(&d, (size_t)blocks * threads * 8);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(a);
        pfb_dmma_chain<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        double fl = 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32);
        double tf = fl / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return e == cudaSuccess ? 0 : (int)e;
}
