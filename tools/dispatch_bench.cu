// Does every non-FP64 instruction cost the FP64 pipe a dispatch cycle?  28 independent DFMAs per iteration
// mixed with K independent integer ops (K = 0, 8, 16, 32), and the same with DMMA instead of DFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, bool MMA>
__global__ void __launch_bounds__(256) mix(double* out, int* iout, int iters, double fa, double fb, int ia) {
    double acc[14];
    int r[8];
#pragma unroll
    for (int c = 0; c < 14; ++c) acc[c] = threadIdx.x * 1e-3 + c;
#pragma unroll
    for (int c = 0; c < 8; ++c) r[c] = threadIdx.x + c;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
        if (MMA) {
            // 4 DMMA = 4 x 8 = 32 FMA per lane ~ 28 DFMA worth of pipe time (x 32/28)
#define MMAI(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(fa), "d"(fb));
            MMAI(c0) MMAI(c1) MMAI(c2) MMAI(c3)
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int c = 0; c < 14; ++c) acc[c] = fma(acc[c], fa, fb);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) r[k & 7] = (r[k & 7] ^ ia) + (r[(k + 3) & 7] & 0x3f);   // LOP3 + IADD-ish (2 ALU ops)
    }
    double s = c0[0] + c1[1] + c2[0] + c3[1];
    int si = 0;
#pragma unroll
    for (int c = 0; c < 14; ++c) s += acc[c];
#pragma unroll
    for (int c = 0; c < 8; ++c) si += r[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = si;
}
template <int K, bool MMA>
void run(double* out, int* iout, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 2, iters = 20000;
    mix<K, MMA><<<blocks, 256>>>(out, iout, 10, 1.0000001, 1e-9, 5); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<K, MMA><<<blocks, 256>>>(out, iout, iters, 1.0000001, 1e-9, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_lane = (MMA ? 32.0 : 28.0) * iters * blocks * 256;
    printf("%s + %2d int-op pairs/iter: %7.3f ms  FP64 %.1f%% of 64 FMA/clk/SM\n", MMA ? "4 DMMA " : "28 DFMA", K, ms,
           100.0 * fma_lane / (ms * 1e-3) / (sms * 64.0 * 1.965e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    double* out; int* iout; cudaMalloc(&out, 8 * sms * 2 * 256); cudaMalloc(&iout, 4 * sms * 2 * 256);
    run<0, false>(out, iout, sms); run<4, false>(out, iout, sms); run<8, false>(out, iout, sms); run<16, false>(out, iout, sms);
    run<0, true>(out, iout, sms); run<4, true>(out, iout, sms); run<8, true>(out, iout, sms); run<16, true>(out, iout, sms);
    return 0;
}
