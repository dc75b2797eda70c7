// single-warp dependent-chain latencies on B200 (cycles): DFMA, DMUL, rsqrt, 1/x, sqrt, shfl(double), LDS, DMMA
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0, int iters) {
    __shared__ double sm[64];
    sm[threadIdx.x] = x0 + threadIdx.x; sm[threadIdx.x + 32] = 1.0;
    __syncwarp();
    double x = x0 + threadIdx.x * 1e-3; long long t0, t1; double acc = 0;
#define RUN(id, ...) { x = x0 + threadIdx.x*1e-3; t0 = clock64(); for (int i = 0; i < iters; ++i) { __VA_ARGS__; } t1 = clock64(); acc += x; if (threadIdx.x==0) cyc[id] = (t1 - t0); }
    RUN(0, x = fma(x, 1.0000001, 1e-9))
    RUN(1, x = x * 1.0000001)
    RUN(2, x = rsqrt(x) + 1.5)
    RUN(3, x = 1.0 / x + 1.5)
    RUN(4, x = sqrt(x) + 1.5)
    RUN(5, x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31))
    RUN(6, x = sm[((int)x) & 31] )
    RUN(7, { double c0 = x, c1 = x; asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(1e-9), "d"(1e-9)); x = c0; })
    RUN(8, x = (double)rsqrtf((float)x) + 1.5)
    RUN(9, { float f = (float)x; f = rsqrtf(f); double y = (double)f; double t = x * y; double e = fma(-t, y, 1.0); y = fma(0.5 * y, e, y); t = x * y; e = fma(-t, y, 1.0); y = fma(0.5 * y, e, y); x = y + 1.5; })
    out[threadIdx.x] = acc;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 128);
    int iters = 2000;
    k<<<1, 32>>>(out, cyc, 1.3, iters); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, cyc, 1.3, iters); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 80, cudaMemcpyDeviceToHost);
    const char* names[] = {"dfma", "dmul", "rsqrt(double)+add", "1/x+add", "sqrt+add", "shfl double", "LDS dependent (incl cvt)", "dmma dependent", "rsqrtf via f32 +cvt+add", "f32 rsqrt + 2 NR + add"};
    for (int i = 0; i < 10; ++i) printf("%-28s %.1f cycles\n", names[i], (double)h[i] / iters);
    return 0;
}
