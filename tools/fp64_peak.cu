// FP64 denominators for the roofline: sustained DFMA and DMMA issue rates, DFMA latency,
// and the accuracy of rsqrt.approx.ftz.f64 (MUFU.RSQ64H) which the K*V kernel's sqrt is seeded from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CH>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double c4[2] = {0, 0}, c5[2] = {0, 0}, c6[2] = {0, 0}, c7[2] = {0, 0};
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
        MMA(c0) MMA(c1) MMA(c2) MMA(c3) MMA(c4) MMA(c5) MMA(c6) MMA(c7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[0] + c3[1] + c4[0] + c5[1] + c6[0] + c7[1];
}

__global__ void dfma_latency(double* out, long long* cyc, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsqrt_acc(const double* in, double* relerr, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = in[i], y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double exact = 1.0 / sqrt(x);
    relerr[i] = fabs(y - exact) / exact;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // DFMA throughput: blocks of 256 threads, 4 per SM
    {
        int iters = 20000; int blocks = sms * 4;
        dfma_loop<8><<<blocks, 256>>>(out, 100, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        double best = 0, sustained = 0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            dfma_loop<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 8 * 8 * (double)iters * blocks * 256;
            double tf = fl / (ms * 1e-3) / 1e12; if (tf > best) best = tf;
        }
        // sustained: ~3 seconds back to back
        cudaEventRecord(e0); int reps = 0;
        for (; reps < 200; ++reps) dfma_loop<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        sustained = 2.0 * 8 * 8 * (double)iters * blocks * 256 * reps / (ms * 1e-3) / 1e12;
        printf(" \"dfma_tflops_burst\": %.3f, \"dfma_tflops_sustained\": %.3f, \"dfma_sustained_seconds\": %.2f,\n", best, sustained, ms * 1e-3);
    }
    {
        int iters = 20000; int blocks = sms * 4;
        dmma_loop<<<blocks, 256>>>(out, 100); CK(cudaDeviceSynchronize());
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            dmma_loop<<<blocks, 256>>>(out, iters);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * 8;  // per warp: 8 mma x 512 flop
            double tf = fl / (ms * 1e-3) / 1e12; if (tf > best) best = tf;
        }
        printf(" \"dmma_tflops_burst\": %.3f,\n", best);
    }
    {
        long long* cyc; CK(cudaMalloc(&cyc, 8)); long long h;
        dfma_latency<<<1, 32>>>(out, cyc, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        dfma_latency<<<1, 32>>>(out, cyc, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" \"dfma_dependent_latency_cycles\": %.2f,\n", (double)h / 16000.0);
    }
    {
        int n = 1 << 20; double* hin = (double*)malloc(n * 8); double* herr = (double*)malloc(n * 8);
        srand(1); for (int i = 0; i < n; ++i) { double u = (rand() + 0.5) / (RAND_MAX + 1.0); hin[i] = exp(-40.0 + 80.0 * u); }
        double *din, *derr; CK(cudaMalloc(&din, n * 8)); CK(cudaMalloc(&derr, n * 8));
        cudaMemcpy(din, hin, n * 8, cudaMemcpyHostToDevice);
        rsqrt_acc<<<n / 256, 256>>>(din, derr, n); CK(cudaDeviceSynchronize());
        cudaMemcpy(herr, derr, n * 8, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < n; ++i) if (herr[i] > mx) mx = herr[i];
        printf(" \"rsqrt_approx_f64_max_relerr\": %.4e, \"rsqrt_log2\": %.2f\n", mx, log2(mx));
    }
    printf("}\n");
    return 0;
}
