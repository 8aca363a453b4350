// Microbenchmark: how DMMA (m8n8k4 f64) overlaps with integer ALU / IMAD / DFMA work on B200.
// Per loop iteration a warp issues ND DMMAs (independent accumulators) and NX "other" ops per DMMA
// (LOP3 chains, IMAD.WIDE chains or DFMAs, 2 independent chains).  Reports cycles per iteration
// per SM sub-partition for 1/2/4 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix dmma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// KIND 0: LOP3 (xor chains)  1: IMAD.WIDE chains  2: DFMA chains   3: mixed philox-like (imad.wide + lop3)
template <int ND, int NX, int KIND>
__global__ void __launch_bounds__(1024) k_mix(double* out, int iters, double a, double b, unsigned k) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    unsigned x0 = threadIdx.x, x1 = threadIdx.x * 3 + 1, x2 = k, x3 = k + 7;
    double f0 = threadIdx.x, f1 = 1.0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            dmma(c[i & 7][0], c[i & 7][1], a, b);
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (KIND == 0) {
                    if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x0) : "r"(x2), "r"(k));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x1) : "r"(x3), "r"(k));
                } else if (KIND == 1) {
                    unsigned long long p;
                    if (j & 1) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x0), "r"(0xD2511F53u)); x0 = (unsigned)(p >> 32); x2 ^= (unsigned)p; }
                    else { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x1), "r"(0xCD9E8D57u)); x1 = (unsigned)(p >> 32); x3 ^= (unsigned)p; }
                } else if (KIND == 2) {
                    if (j & 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f0) : "d"(a), "d"(b));
                    else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f1) : "d"(a), "d"(b));
                } else if (KIND == 4) {
                    // philox round with separate mul.hi / mul.lo instead of mul.wide
                    unsigned ph, pl, qh, ql;
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(ph) : "r"(x0), "r"(0xD2511F53u));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(pl) : "r"(x0), "r"(0xD2511F53u));
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(qh) : "r"(x2), "r"(0xCD9E8D57u));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(ql) : "r"(x2), "r"(0xCD9E8D57u));
                    unsigned n0, n2;
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n0) : "r"(qh), "r"(x1), "r"(k));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n2) : "r"(ph), "r"(x3), "r"(k));
                    x0 = n0; x1 = ql; x2 = n2; x3 = pl;
                } else if (KIND == 5) {
                    // 16-bit split multiplies on the full-rate 32-bit IMAD path (x * M = (xh * M) << 16 + xl * M needs carries: cost model only)
                    unsigned a0 = x0 & 0xFFFFu, a1 = x0 >> 16, b0 = x2 & 0xFFFFu, b1 = x2 >> 16;
                    unsigned p0, p1, q0, q1;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(p0) : "r"(a0), "r"(0x1F53u), "r"(x1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(p1) : "r"(a1), "r"(0xD251u), "r"(p0));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(q0) : "r"(b0), "r"(0x8D57u), "r"(x3));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(q1) : "r"(b1), "r"(0xCD9Eu), "r"(q0));
                    x0 = q1 ^ k; x1 = p0; x2 = p1 ^ k; x3 = q0;
                } else {
                    unsigned long long p, q;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x0), "r"(0xD2511F53u));
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(x2), "r"(0xCD9E8D57u));
                    unsigned n0, n2;
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n0) : "r"((unsigned)(q >> 32)), "r"(x1), "r"(k));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n2) : "r"((unsigned)(p >> 32)), "r"(x3), "r"(k));
                    x0 = n0; x1 = (unsigned)q; x2 = n2; x3 = (unsigned)p;
                }
            }
        }
    }
    double s = f0 + f1 + x0 + x1 + x2 + x3;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

template <int ND, int NX, int KIND>
void run(const char* name, double* d, int sms) {
    const int iters = 2048;
    printf("%-28s ND=%d NX=%2d :", name, ND, NX);
    for (int threads : {128, 256, 512, 1024}) {
        double t = timeit([&] { k_mix<ND, NX, KIND><<<sms, threads>>>(d, iters, 1.0000001, 1e-9, 12345u); });
        double cyc = t * 1.965e9 / iters;  // cycles per iteration (per warp, wall)
        int wps = threads / 128;           // warps per scheduler
        printf("  %dw/sched %.0f clk/iter (%.1f per sched-iter)", wps, cyc, cyc / wps);
    }
    printf("\n");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, (size_t)sms * 1024 * 8);
    run<8, 0, 0>("DMMA only", d, sms);
    run<0 + 8, 4, 0>("DMMA + 4 LOP3 each", d, sms);
    run<8, 8, 0>("DMMA + 8 LOP3 each", d, sms);
    run<8, 16, 0>("DMMA + 16 LOP3 each", d, sms);
    run<8, 8, 1>("DMMA + 8 IMAD.WIDE each", d, sms);
    run<8, 16, 1>("DMMA + 16 IMAD.WIDE each", d, sms);
    run<8, 2, 2>("DMMA + 2 DFMA each", d, sms);
    run<8, 4, 2>("DMMA + 4 DFMA each", d, sms);
    run<8, 2, 3>("DMMA + 2 philox rounds each", d, sms);
    run<8, 4, 3>("DMMA + 4 philox rounds each", d, sms);
    run<8, 5, 3>("DMMA + 5 philox rounds each", d, sms);
    run<8, 2, 4>("DMMA + 2 philox(hi/lo) rounds each", d, sms);
    run<8, 5, 4>("DMMA + 5 philox(hi/lo) rounds each", d, sms);
    run<1, 20, 4>("1 DMMA + 20 philox(hi/lo) rounds", d, sms);
    run<8, 2, 5>("DMMA + 2x4 mad.lo.u32 each", d, sms);
    run<8, 5, 5>("DMMA + 5x4 mad.lo.u32 each", d, sms);
    run<1, 20, 5>("1 DMMA + 20x4 mad.lo.u32", d, sms);
    run<1, 20, 3>("1 DMMA + 20 philox rounds", d, sms);
    run<1, 40, 0>("1 DMMA + 40 LOP3", d, sms);
    run<1, 40, 1>("1 DMMA + 40 IMAD.WIDE", d, sms);
    return 0;
}
