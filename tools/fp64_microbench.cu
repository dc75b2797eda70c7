// FP64 pipe microbenchmarks for B200 (sm_100a): DMMA.8x8x4 vs DFMA issue-rate ceilings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
        for (int threads = 128; threads <= 256; threads *= 2) {
            int iters = 20000;
            // DMMA
            dmma_loop<16><<<nsm * ctas_per_sm, threads>>>(out, 100, 1.0, 1.0);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int r = 0; r < 3; ++r) {
                CK(cudaEventRecord(e0));
                dmma_loop<16><<<nsm * ctas_per_sm, threads>>>(out, iters, 1.0, 1.0);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            double flops = 2.0 * 256.0 * 16.0 * iters * (threads / 32) * nsm * ctas_per_sm;
            printf("DMMA.8x8x4  ctas/sm %d threads %d : %.3f ms  %.2f TFLOP/s\n", ctas_per_sm, threads, best, flops / best * 1e-9);
            // DFMA
            dfma_loop<16><<<nsm * ctas_per_sm, threads>>>(out, 100, 1.0, 1.0);
            CK(cudaDeviceSynchronize());
            best = 1e30f;
            for (int r = 0; r < 3; ++r) {
                CK(cudaEventRecord(e0));
                dfma_loop<16><<<nsm * ctas_per_sm, threads>>>(out, iters, 0.999999, 1e-3);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            flops = 2.0 * 16.0 * iters * threads * nsm * ctas_per_sm;
            printf("DFMA        ctas/sm %d threads %d : %.3f ms  %.2f TFLOP/s\n", ctas_per_sm, threads, best, flops / best * 1e-9);
        }
    }
    return 0;
}
