// Context lifetime, error text, workspaces.
#include <math.h>
#include <stdlib.h>
#include <stdarg.h>

#include "common.cuh"

namespace cglb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ensure_vpad(Context* ctx, long n_pad) {
    if (n_pad <= ctx->vpad_cap) return CGLB_OK;
    // growing a workspace is a (rare) synchronising operation: wait for in-flight users first
    CGLB_CUDA_OK(cudaDeviceSynchronize());
    if (ctx->vpad) cudaFree(ctx->vpad);
    if (ctx->upad) cudaFree(ctx->upad);
    if (ctx->rsum) cudaFree(ctx->rsum);
    ctx->vpad = ctx->upad = ctx->rsum = nullptr;
    ctx->vpad_cap = 0;
    long cap = n_pad + n_pad / 8;
    cap = (cap + 1023) / 1024 * 1024;
    CGLB_CUDA_OK(cudaMalloc(&ctx->vpad, sizeof(double) * cap));
    CGLB_CUDA_OK(cudaMalloc(&ctx->upad, sizeof(double) * cap));
    CGLB_CUDA_OK(cudaMalloc(&ctx->rsum, sizeof(double) * cap));
    ctx->vpad_cap = cap;
    return CGLB_OK;
}

int ensure_scratch(Context* ctx, long n_doubles) {
    if (n_doubles <= ctx->scratch_cap) return CGLB_OK;
    CGLB_CUDA_OK(cudaDeviceSynchronize());
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_cap = 0;
    long cap = n_doubles + n_doubles / 4 + 1024;
    CGLB_CUDA_OK(cudaMalloc(&ctx->scratch, sizeof(double) * cap));
    ctx->scratch_cap = cap;
    return CGLB_OK;
}

int ensure_ypart(Context* ctx, long n_doubles) {
    if (n_doubles <= ctx->ypart_cap) return CGLB_OK;
    CGLB_CUDA_OK(cudaDeviceSynchronize());
    if (ctx->ypart) cudaFree(ctx->ypart);
    ctx->ypart = nullptr;
    ctx->ypart_cap = 0;
    long cap = n_doubles + n_doubles / 16 + 1024;
    CGLB_CUDA_OK(cudaMalloc(&ctx->ypart, sizeof(double) * cap));
    ctx->ypart_cap = cap;
    return CGLB_OK;
}

}  // namespace cglb

using namespace cglb;

extern "C" int cglb_abi_version(void) { return CGLB_ABI_VERSION; }
extern "C" const char* cglb_last_error(void) { return g_err; }

extern "C" int cglb_create(cglb_context** out, int device) {
    if (!out) {
        set_error("cglb_create: null output pointer");
        return CGLB_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("cglb_create: no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return CGLB_ERR_CUDA;
    }
    CGLB_CHECK_ARG(device >= 0 && device < count, "device index");
    CGLB_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CGLB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("cglb_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
        return CGLB_ERR_UNSUPPORTED;
    }
    Context* ctx = new Context();
    memset(ctx, 0, sizeof(Context));
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char* e = getenv("CGLB_DSWEEP");
        ctx->opt_dsweep = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
        e = getenv("CGLB_GEMM_STAGING");
        ctx->opt_gemm_staging = (e && e[0] == '2') ? 2 : 1;
    }
    CGLB_CUDA_OK(cudaMalloc(&ctx->counters, sizeof(int) * 16));
    CGLB_CUDA_OK(cudaMemset(ctx->counters, 0, sizeof(int) * 16));
    static double tab[kExpTabSmall + kExpTabBig];
    for (int j = 0; j < kExpTabSmall; ++j) tab[j] = exp2((double)j / kExpTabSmall);
    for (int j = 0; j < kExpTabBig; ++j) {
        // fast_exp_neg<10> adds n * 2^10 = ((n >> 10) << 20) + (j << 10) to the high word: pre-subtract j << 10
        const double t = exp2((double)j / kExpTabBig);
        unsigned long long bits;
        memcpy(&bits, &t, sizeof(bits));
        bits -= (unsigned long long)j << 42;
        memcpy(&tab[kExpTabSmall + j], &bits, sizeof(bits));
    }
    CGLB_CUDA_OK(cudaMalloc(&ctx->exp_table, sizeof(tab)));
    CGLB_CUDA_OK(cudaMemcpy(ctx->exp_table, tab, sizeof(tab), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<cglb_context*>(ctx);
    return CGLB_OK;
}

extern "C" int cglb_set_option(cglb_context* c, const char* name, long value) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && name, "null pointer");
    if (strcmp(name, "dsweep") == 0) {
        CGLB_CHECK_ARG(value >= 0 && value <= 2, "dsweep: 0, 1 or 2");
        ctx->opt_dsweep = (int)value;
        return CGLB_OK;
    }
    if (strcmp(name, "superrow") == 0) {
        CGLB_CHECK_ARG(value >= 0, "superrow: chunks per super-row, 0 = automatic");
        ctx->opt_superrow = value;
        return CGLB_OK;
    }
    if (strcmp(name, "gemm_staging") == 0) {
        CGLB_CHECK_ARG(value == 1 || value == 2, "gemm_staging: 1 (cp.async) or 2 (TMA)");
        ctx->opt_gemm_staging = (int)value;
        return CGLB_OK;
    }
    set_error("cglb_set_option: unknown option '%s'", name);
    return CGLB_ERR_ARG;
}

extern "C" int cglb_destroy(cglb_context* c) {
    Context* ctx = reinterpret_cast<Context*>(c);
    if (!ctx) return CGLB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->vpad) cudaFree(ctx->vpad);
    if (ctx->upad) cudaFree(ctx->upad);
    if (ctx->rsum) cudaFree(ctx->rsum);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->ypart) cudaFree(ctx->ypart);
    if (ctx->exp_table) cudaFree(ctx->exp_table);
    delete ctx;
    return CGLB_OK;
}

extern "C" unsigned long long cglb_launch_count(const cglb_context* c) {
    return c ? reinterpret_cast<const Context*>(c)->launches : 0ULL;
}
extern "C" int cglb_num_sms(const cglb_context* c) { return c ? reinterpret_cast<const Context*>(c)->num_sms : 0; }
