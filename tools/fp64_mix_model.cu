// DMMA.8x8x4 + DFMA mixing on the shared FP64 pipe of sm_100a (developer micro-benchmark).
// Per iteration every warp issues 12 DMMA (4 accumulators x 3 chained) and 120 DFMA (8 chains x 15), the ratio of
// the d = 11 Matern32 sweep, either SPREAD (1 DMMA every 10 DFMA) or CLUSTERED (12 DMMA, then 120 DFMA).
#include <cstdio>
#include <cuda_runtime.h>
#define MMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
template <int MODE>   // 0 spread, 1 clustered, 2 DFMA only, 3 DMMA only
__global__ void mix(double* out, const double* in, int iters) {
    double a[8], c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = in[c] + threadIdx.x;
    const double X = in[100], Y = in[101], fa = in[102], fb = in[103];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int s = 0; s < 12; ++s) {
                if (s % 4 == 0) { MMA(c0, fa, fb) } else if (s % 4 == 1) { MMA(c1, fa, fb) } else if (s % 4 == 2) { MMA(c2, fa, fb) } else { MMA(c3, fa, fb) }
#pragma unroll
                for (int u = 0; u < 10; ++u) a[(s * 10 + u) % 8] = fma(a[(s * 10 + u) % 8], X, Y);
            }
        } else {
            if (MODE != 2) {
#pragma unroll
                for (int s = 0; s < 3; ++s) { MMA(c0, fa, fb) MMA(c1, fa, fb) MMA(c2, fa, fb) MMA(c3, fa, fb) }
            }
            if (MODE != 3) {
#pragma unroll
                for (int u = 0; u < 120; ++u) a[u % 8] = fma(a[u % 8], X, Y);
            }
        }
    }
    double s = c0[0] + c1[1] + c2[0] + c3[1];
#pragma unroll
    for (int c = 0; c < 8; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(double* out, double* in, int sms, int w) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 128 * w, iters = 4000;
    mix<MODE><<<sms, threads>>>(out, in, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<MODE><<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double per_lane = (MODE == 2 ? 120.0 : MODE == 3 ? 96.0 : 216.0);   // 12 DMMA = 12 * 256 / 32 = 96 FMA per lane
    const char* names[] = {"spread 12 DMMA + 120 DFMA", "clustered 12 DMMA + 120 DFMA", "120 DFMA only", "12 DMMA only"};
    printf("%-30s warps/smsp=%d : %6.1f %% of 64 FMA/clk/SM\n", names[MODE], w,
           100.0 * per_lane * iters * (double)sms * threads / (ms * 1e-3) / (sms * 64.0 * 1.965e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, 8 * sms * 1024); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {1, 2, 4}) { run<0>(out, in, sms, w); run<1>(out, in, sms, w); run<2>(out, in, sms, w); run<3>(out, in, sms, w); }
    return 0;
}
