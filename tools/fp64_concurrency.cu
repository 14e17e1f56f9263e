// Do DMMA (mma.sync.m8n8k4.f64) and DFMA share one execution unit on B200?  Half the warps of every CTA run a
// DMMA loop, the other half a DFMA loop; if the combined rate exceeds either peak, the pipes are separate.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__global__ void __launch_bounds__(256) mixed(double* out, int iters, int mode, double fa, double fb) {
    const int warp = threadIdx.x >> 5;
    const bool do_mma = (mode == 1) || (mode == 2 && (warp & 1));
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    if (do_mma) {
        double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0}, c4[2] = {0, 0}, c5[2] = {0, 0}, c6[2] = {0, 0}, c7[2] = {0, 0};
        double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-9;
        for (int it = 0; it < iters; ++it) {
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
            MMA(c0) MMA(c1) MMA(c2) MMA(c3) MMA(c4) MMA(c5) MMA(c6) MMA(c7)
        }
        acc[0] = c0[0] + c1[1] + c2[0] + c3[1] + c4[0] + c5[1] + c6[0] + c7[1];
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int blocks = p.multiProcessorCount * 4, iters = 20000;
    double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"dfma_only", "dmma_only", "half_dmma_half_dfma"};
    for (int mode = 0; mode < 3; ++mode) {
        mixed<<<blocks, 256>>>(out, 100, mode, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); mixed<<<blocks, 256>>>(out, iters, mode, 1.0000001, 1e-9); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        // per warp per iteration: DFMA path 64 DFMA x 32 lanes x 2 flop = 4096 flop; DMMA path 8 x 512 = 4096 flop
        double flop = 4096.0 * iters * blocks * 8;
        printf("%s: %.3f ms  %.2f TFLOP/s\n", names[mode], best, flop / (best * 1e-3) / 1e12);
    }
    return 0;
}
