// FP64 issue model of sm_100a (developer micro-benchmark): throughput of dependent DFMA chains as a function of
// the number of independent chains per warp (ILP), warps per SM sub-partition (TLP) and how many DISTINCT
// register operands each DFMA reads (register-file bandwidth).   nvcc -arch=sm_100a -O3 fp64_issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a = fma(a, X, Y)  X, Y shared by all chains (1 distinct operand per DFMA, reuse cache hits)
// MODE 1: a = fma(a, x[c], Y)  (2 distinct)      MODE 2: a = fma(a, x[c], y[c])  (3 distinct)
// MODE 3: a = fma(x[c], y[c], a) with x,y rotating (3 distinct, no operand shared with the previous DFMA)
template <int NCH, int MODE>
__global__ void chains(double* out, const double* in, int iters) {
    double a[NCH], x[NCH], y[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { a[c] = in[c] + threadIdx.x; x[c] = in[32 + c]; y[c] = in[64 + c]; }
    const double X = in[100], Y = in[101];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (MODE == 0) a[c] = fma(a[c], X, Y);
                else if (MODE == 1) a[c] = fma(a[c], x[c], Y);
                else if (MODE == 2) a[c] = fma(a[c], x[c], y[c]);
                else a[c] = fma(x[(c + u) % NCH], y[(c + 2 * u + 1) % NCH], a[c]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH, int MODE>
void run(double* out, double* in, int sms, int warps_per_smsp) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 128 * warps_per_smsp, iters = 4000;
    chains<NCH, MODE><<<sms, threads>>>(out, in, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); chains<NCH, MODE><<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma_lane = 8.0 * NCH * iters * (double)sms * threads;
    printf("chains=%2d mode=%d warps/smsp=%d : %6.1f %% of 64 FMA/clk/SM\n", NCH, MODE, warps_per_smsp,
           100.0 * fma_lane / (ms * 1e-3) / (sms * 64.0 * 1.965e9));
}
template <int MODE>
void sweep(double* out, double* in, int sms) {
    for (int w : {1, 2, 4}) {
        run<1, MODE>(out, in, sms, w); run<2, MODE>(out, in, sms, w); run<4, MODE>(out, in, sms, w);
        run<8, MODE>(out, in, sms, w); run<16, MODE>(out, in, sms, w);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, 8 * sms * 1024); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    sweep<0>(out, in, sms); sweep<1>(out, in, sms); sweep<2>(out, in, sms); sweep<3>(out, in, sms);
    return 0;
}
