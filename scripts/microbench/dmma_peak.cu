// Microbenchmark: FP64 tensor-core (mma.sync m8n8k4 f64, SASS DMMA) vs DFMA throughput on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, (size_t)sms * 8 * 1024 * 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int blocks = sms * bps;
            double t = timeit([&] { k_dmma<8><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 256 * 8.0 * iters * (double)blocks * (threads / 32);
            double t1 = timeit([&] { k_dmma<1><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
            double fl1 = 2.0 * 256 * 1.0 * iters * (double)blocks * (threads / 32);
            double t2 = timeit([&] { k_dfma<8><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
            double fl2 = 2.0 * 8.0 * iters * (double)blocks * threads;
            printf("threads/CTA %4d CTAs/SM %d : DMMA(8 acc) %.2f TF  DMMA(1 acc, latency-bound) %.2f TF (%.1f clk/dep-MMA @1.965GHz)  DFMA %.2f TF\n",
                   threads, bps, fl / t / 1e12, fl1 / t1 / 1e12, t1 / iters * 1.965e9, fl2 / t2 / 1e12);
        }
    }
    return 0;
}
