// extern "C" entry points of the matrix-free sweeps (K1, K2) + input packing.
#include <stdlib.h>

#include "kmv_impl.cuh"
#include "f32sweep_impl.cuh"

namespace cglb {

__global__ void pack_inputs_kernel(int kind, const double* __restrict__ x, long n, long n_pad, int d, int dp,
                                   const double* __restrict__ ls, const double* __restrict__ shift,
                                   double* __restrict__ xp) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double* o = xp + i * dp;
    if (i >= n) {
        for (int k = 0; k < dp; ++k) o[k] = 0.0;
        return;
    }
    const double c = (kind == CGLB_MATERN32) ? 1.7320508075688772935 : 0.70710678118654752440;
    double nrm = 0.0;
    for (int k = 0; k < d; ++k) {
        double a = c * (x[i * d + k] - (shift ? shift[k] : 0.0)) / ls[k];
        o[k] = a;
        nrm = fma(a, a, nrm);
    }
    for (int k = d; k < dp; ++k) o[k] = 0.0;
    o[norm_index(d)] = nrm;
}

// fp32 packed layout of the fp32-pair sweeps (f32sweep_impl.cuh): coordinates scaled in fp64, rounded to float;
// the squared norm is that of the ROUNDED coordinates (the expanded form needs |a|^2 consistent with a)
__global__ void pack_inputs_f32_kernel(int kind, const double* __restrict__ x, long n, long n_pad, int d, int dpf,
                                       const double* __restrict__ ls, const double* __restrict__ shift, float* __restrict__ xp) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float* o = xp + i * dpf;
    if (i >= n) {
        for (int k = 0; k < dpf; ++k) o[k] = 0.0f;
        return;
    }
    // the fp64 scaling times the base-2 factor of f32sweep_impl.cuh (the exponentials run on MUFU.EX2)
    const double c = ((kind == CGLB_MATERN32) ? 1.7320508075688772935 : 0.70710678118654752440) * f32_input_scale(kind);
    double nrm = 0.0;
    for (int k = 0; k < d; ++k) {
        const float a = (float)(c * (x[i * d + k] - (shift ? shift[k] : 0.0)) / ls[k]);
        o[k] = a;
        nrm = fma((double)a, (double)a, nrm);
    }
    for (int k = d; k < dpf; ++k) o[k] = 0.0f;
    o[dpf - 1] = (float)nrm;
}

// vpad32 = float([v, 0...]); y = diag * v
__global__ void kmv_prologue_f32_kernel(const double* __restrict__ v, long n, long n_pad, float* __restrict__ vpad,
                                        double* __restrict__ y, double diag) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) vpad[i] = (i < n) ? (float)v[i] : 0.0f;
    if (i < n) y[i] = diag * v[i];
}

__global__ void bwd_prologue_f32_kernel(const double* __restrict__ u, const double* __restrict__ w, long n, long n_pad,
                                        float* __restrict__ upad, float* __restrict__ wpad, double* __restrict__ rsum) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) {
        upad[i] = (i < n) ? (float)u[i] : 0.0f;
        wpad[i] = (i < n) ? (float)w[i] : 0.0f;
        rsum[i] = 0.0;
    }
}

// vpad = [v, 0...]; y = diag * v (or 0)
__global__ void kmv_prologue_kernel(const double* __restrict__ v, long n, long n_pad, double* __restrict__ vpad,
                                    double* __restrict__ y, long ny, double diag) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) vpad[i] = (i < n) ? v[i] : 0.0;
    if (y != nullptr && i < ny) y[i] = (i < n) ? diag * v[i] : 0.0;
}

__global__ void bwd_prologue_kernel(const double* __restrict__ u, const double* __restrict__ w, long n, long n_pad,
                                    double* __restrict__ upad, double* __restrict__ wpad, double* __restrict__ rsum) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) {
        upad[i] = (i < n) ? u[i] : 0.0;
        wpad[i] = (i < n) ? w[i] : 0.0;
        rsum[i] = 0.0;
    }
}

// Vpad[i][r] = V[i][r] for r < t, 0 for t <= r < T and for padded rows (multi-RHS sweep: T = 2 or 4 columns per group)
__global__ void multi_prologue_kernel(const double* __restrict__ v, long n, long n_pad, int t, int t0, int T, double* __restrict__ vpad) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pad * T) return;
    const long i = idx / T;
    const int r = (int)(idx % T);
    vpad[idx] = (i < n && t0 + r < t) ? v[i * t + t0 + r] : 0.0;
}

// Y[i][t0 + r] = diag * V[i][t0 + r] + sum_c part[c][i][r]   (fixed CTA order)
__global__ void multi_reduce_kernel(const double* __restrict__ part, long stride, int nslots, const double* __restrict__ v, double diag,
                                    double* __restrict__ y, long n, int t, int t0, int T) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * T) return;
    const long i = idx / T;
    const int r = (int)(idx % T);
    if (t0 + r >= t) return;
    double s = diag * v[i * t + t0 + r];
#pragma unroll 8
    for (int c = 0; c < nslots; ++c) s += part[(long)c * stride + idx];
    y[i * t + t0 + r] = s;
}

// Second stage of the fixed-order reductions: the sweeps' CTAs accumulate into their own copies of the output vector
// (part + c * stride); here the copies are summed in CTA order.  y[i] = diag * v[i] + sum_c part[c][i].
__global__ void reduce_parts_kernel(const double* __restrict__ part, long stride, int nslots, const double* __restrict__ v,
                                    double diag, double* __restrict__ y, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = (v != nullptr) ? diag * v[i] : 0.0;
#pragma unroll 8
    for (int c = 0; c < nslots; ++c) s += part[(long)c * stride + i];
    y[i] = s;
}

// backward sweep: R[i] = sum_c part[c][i]; the last block also sums the per-CTA gradient slots into gtmp[0 .. d]
__global__ void reduce_bwd_parts_kernel(const double* __restrict__ part, long stride, int nslots, double* __restrict__ rsum, long n,
                                        const double* __restrict__ gpart, long gstride, int d, double* __restrict__ gtmp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double s = 0.0;
#pragma unroll 8
        for (int c = 0; c < nslots; ++c) s += part[(long)c * stride + i];
        rsum[i] = s;
    }
    if (blockIdx.x == gridDim.x - 1 && (int)threadIdx.x <= d) {
        double s = 0.0;
        for (int c = 0; c < nslots; ++c) s += gpart[(long)c * gstride + threadIdx.x];
        gtmp[threadIdx.x] = s;
    }
}

// out[q] += (variance*cfac/ls_q) * ( sum_i a_iq^2 R_i + gq[q] ) ; out[d] += gvar      (single CTA per q chunk)
__global__ void bwd_epilogue_kernel(const double* __restrict__ xp, long n, int d, int dp, const double* __restrict__ rsum,
                                    const double* __restrict__ gtmp, const double* __restrict__ ls, double variance,
                                    double cfac, double* __restrict__ out, int add_sweep_terms) {
    // grid: d blocks (one per lengthscale) ; block 256 threads, deterministic tree reduction
    const int q = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        double a = xp[i * dp + q];
        s = fma(a * a, rsum[i], s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double tot = sh[0] + (add_sweep_terms ? gtmp[q] : 0.0);
        out[q] += variance * cfac / ls[q] * tot;
        if (q == 0 && add_sweep_terms) out[d] += gtmp[d];
    }
}


// CGLB_KMV_DIMS_LIST = "X(1) X(2) ..." : the dimensions this build instantiates (build.py passes it)
#ifndef CGLB_KMV_DIMS_LIST
#define CGLB_KMV_DIMS_LIST                                                                                 \
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) \
        X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)
#endif
#define X(DD)                                                              \
    int sweep_d##DD(Context*, int, int, const SweepArgs&, cudaStream_t);   \
    int f32_d##DD(Context*, int, const SweepArgsF32&, cudaStream_t);       \
    int f32_bwd_d##DD(Context*, int, const BwdArgsF32&, cudaStream_t);     \
    int knm_d##DD(Context*, int, int, const KnmArgs&, cudaStream_t);
CGLB_KMV_DIMS_LIST
#undef X

knm_fn get_knm_fn(int d) {
    switch (d) {
#define X(DD) \
    case DD:  \
        return knm_d##DD;
        CGLB_KMV_DIMS_LIST
#undef X
        default:
            return nullptr;
    }
}

f32_fn get_f32_fn(int d) {
    switch (d) {
#define X(DD) \
    case DD:  \
        return f32_d##DD;
        CGLB_KMV_DIMS_LIST
#undef X
        default:
            return nullptr;
    }
}

f32_bwd_fn get_f32_bwd_fn(int d) {
    switch (d) {
#define X(DD) \
    case DD:  \
        return f32_bwd_d##DD;
        CGLB_KMV_DIMS_LIST
#undef X
        default:
            return nullptr;
    }
}

sweep_fn get_sweep_fn(int d) {
    switch (d) {
#define X(DD) \
    case DD:  \
        return sweep_d##DD;
        CGLB_KMV_DIMS_LIST
#undef X
        default:
            return nullptr;
    }
}

int knm_build_wide(Context* ctx, int kind, const double* zp, long m, const double* xp, long n, int d, double variance,
                   double* out, long ld, cudaStream_t st);
int knm_backward_wide(Context* ctx, int kind, const double* zp, long m, const double* xp, long ncols, int d, double variance,
                      const double* lengthscale, double* t, long ldt, const double* wt, const double* zvec, double* out_ls,
                      double* out_var, double* out_z, cudaStream_t st);
int wide_sweep(Context* ctx, int kind, bool sym, const double* xp_rows, long nrows, const double* xp_cols, long ncols, int d,
               const double* vcol, double* y, long ystride, double variance, int part, int nparts, cudaStream_t st);

int wide_bwd_sweep(Context* ctx, int kind, const double* xp, long n, int d, const double* wcol, const double* ucol, double* rsum,
                   long ystride, double* gout, long gstride, int part, int nparts, cudaStream_t st);

bool dsweep_supported(const Context* ctx, int d, long n, int nparts);
bool dbwd_supported(const Context* ctx, int d, long n, int nparts);

// developer option "dsweep" (cglb_set_option; read from CGLB_DSWEEP once at cglb_create): 0 routes every d <= 32 symmetric sweep
// through the register-resident kernel, 2 forces the DMMA sweep wherever the packed width allows it (tests of small shapes)
static int dsweep_mode(const Context* ctx) { return ctx->opt_dsweep; }

static int dispatch(Context* ctx, int kind, int d, int mode, const SweepArgs& a, cudaStream_t st) {
    if (mode == 0 && d <= CGLB_MAX_REGISTER_D) {
        const int dm = dsweep_mode(ctx);
        if (dm != 0 && dsweep_supported(ctx, d, a.nrows, dm == 2 ? 0 : a.nparts)) mode = 3;
    }
    if (mode == 2 && d <= CGLB_MAX_REGISTER_D) {
        const int dm = dsweep_mode(ctx);
        if (dm != 0 && dbwd_supported(ctx, d, a.nrows, dm == 2 ? 0 : a.nparts)) mode = 4;
    }
    if (d > CGLB_MAX_REGISTER_D && mode == 2)
        return wide_bwd_sweep(ctx, kind, a.xp_rows, a.nrows, d, a.vcol, a.ucol, a.y, a.ystride, a.gout, a.gstride, a.part, a.nparts, st);
    if (d > CGLB_MAX_REGISTER_D)
        return wide_sweep(ctx, kind, mode == 0, a.xp_rows, a.nrows, a.xp_cols, a.ncols, d, a.vcol, a.y, a.ystride, a.variance, a.part,
                          a.nparts, st);
    sweep_fn f = get_sweep_fn(d);
    if (!f) {
        set_error("kernel sweep: d=%d has no register-resident instantiation in this build", d);
        return CGLB_ERR_UNSUPPORTED;
    }
    return f(ctx, kind, mode, a, st);
}

}  // namespace cglb

using namespace cglb;

extern "C" int cglb_packed_width(int d) { return packed_width(d); }
extern "C" long cglb_padded_rows(long n) { return padded_rows(n); }

extern "C" int cglb_pack_inputs(cglb_context* c, int kind, const double* x, long n, int d, const double* lengthscale,
                                const double* shift, double* xp, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && x && lengthscale && xp, "null pointer");
    CGLB_CHECK_ARG(n >= 0 && d >= 1, "n >= 0, d >= 1");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    long n_pad = padded_rows(n);
    if (n_pad == 0) return CGLB_OK;
    pack_inputs_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, x, n, n_pad, d, packed_width(d),
                                                                                         lengthscale, shift, xp);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_packed_width_f32(int d) { return packed_width_f32(d); }

extern "C" int cglb_pack_inputs_f32(cglb_context* c, int kind, const double* x, long n, int d, const double* lengthscale,
                                    const double* shift, float* xpf, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && x && lengthscale && xpf, "null pointer");
    CGLB_CHECK_ARG(n >= 0 && d >= 1, "n >= 0, d >= 1");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    const long n_pad = padded_rows(n);
    if (n_pad == 0) return CGLB_OK;
    pack_inputs_f32_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, x, n, n_pad, d, packed_width_f32(d),
                                                                                             lengthscale, shift, xpf);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_kmv_sym_f32(cglb_context* c, int kind, const float* xpf, long n, int d, const double* v, double* y,
                                double variance, double diag, int part, int nparts, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, "part/nparts");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    if (n == 0) return CGLB_OK;
    CGLB_CHECK_ARG(xpf && v && y, "null pointer");
    f32_fn f = d <= CGLB_MAX_REGISTER_D ? get_f32_fn(d) : nullptr;
    if (!f) {
        set_error("fp32-pair sweep: d=%d is not instantiated (d <= %d only; use cglb_kmv_sym)", d, CGLB_MAX_REGISTER_D);
        return CGLB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long v_pad = (n + 2047) / 2048 * 2048;          // rows are read in blocks of up to 2048
    int rc = ensure_vpad(ctx, v_pad);                     // the float copy of v lives in the (double) u workspace
    if (rc) return rc;
    float* vpad32 = reinterpret_cast<float*>(ctx->upad);
    const int nslots = ctx->num_sms;
    rc = ensure_ypart(ctx, (long)nslots * v_pad);
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * v_pad, st));
    kmv_prologue_f32_kernel<<<(unsigned)((v_pad + 255) / 256), 256, 0, st>>>(v, n, v_pad, vpad32, y, part == 0 ? diag : 0.0);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    SweepArgsF32 a{};
    a.xp = xpf; a.vcol = vpad32; a.y = ctx->ypart; a.ystride = v_pad; a.n = n; a.variance = variance; a.part = part; a.nparts = nparts;
    rc = f(ctx, kind, a, st);
    if (rc) return rc;
    reduce_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->ypart, v_pad, nslots, v, part == 0 ? diag : 0.0, y, n);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_kmv_bwd_sym_f32(cglb_context* c, int kind, const float* xpf, const double* xp, long n, int d, const double* u,
                                    const double* w, double variance, const double* lengthscale, double* out, int part,
                                    int nparts, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, "part/nparts");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    if (n == 0) return CGLB_OK;
    CGLB_CHECK_ARG(xpf && xp && u && w && out && lengthscale, "null pointer");
    f32_bwd_fn f = d <= CGLB_MAX_REGISTER_D ? get_f32_bwd_fn(d) : nullptr;
    if (!f) {
        set_error("fp32-pair backward sweep: d=%d is not instantiated (d <= %d only; use cglb_kmv_bwd_sym)", d, CGLB_MAX_REGISTER_D);
        return CGLB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long v_pad = (n + 1023) / 1024 * 1024;
    int rc = ensure_vpad(ctx, v_pad);
    if (rc) return rc;
    rc = ensure_scratch(ctx, kScratchScalars);
    if (rc) return rc;
    // float copies of u and w live in the (double) u / v workspaces
    float* upad32 = reinterpret_cast<float*>(ctx->upad);
    float* wpad32 = reinterpret_cast<float*>(ctx->vpad);
    bwd_prologue_f32_kernel<<<(unsigned)((v_pad + 255) / 256), 256, 0, st>>>(u, w, n, v_pad, upad32, wpad32, ctx->rsum);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    const int nslots = ctx->num_sms;
    rc = ensure_ypart(ctx, (long)nslots * (v_pad + kScratchScalars));
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * (v_pad + kScratchScalars), st));
    double* gpart = ctx->ypart + (long)nslots * v_pad;
    BwdArgsF32 a{};
    a.xp = xpf; a.wcol = wpad32; a.ucol = upad32; a.rsum = ctx->ypart; a.ystride = v_pad; a.gout = gpart; a.gstride = kScratchScalars;
    a.n = n; a.part = part; a.nparts = nparts;
    rc = f(ctx, kind, a, st);
    if (rc) return rc;
    reduce_bwd_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->ypart, v_pad, nslots, ctx->rsum, n, gpart, kScratchScalars, d,
                                                                          ctx->scratch);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    const double cfac = (kind == CGLB_MATERN32) ? 1.0 : 2.0;
    bwd_epilogue_kernel<<<d, 256, 0, st>>>(xp, n, d, packed_width(d), ctx->rsum, ctx->scratch, lengthscale, variance, cfac, out, 1);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_kmv_sym_variant(const cglb_context* c, int d, long n, int nparts) {
    const Context* ctx = reinterpret_cast<const Context*>(c);
    if (!ctx || d < 1 || n < 0 || nparts < 1) return CGLB_ERR_ARG;
    if (d > CGLB_MAX_REGISTER_D) return 2;
    const int dm = dsweep_mode(ctx);
    return (dm != 0 && dsweep_supported(ctx, d, n, dm == 2 ? 0 : nparts)) ? 1 : 0;
}

extern "C" int cglb_kmv_sym(cglb_context* c, int kind, const double* xp, long n, int d, const double* v, double* y,
                            double variance, double diag, int part, int nparts, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, "part/nparts");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    if (n == 0) return CGLB_OK;      // empty input: nothing to do (pointers may be null)
    CGLB_CHECK_ARG(xp && v && y, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    long n_pad = padded_rows(n);
    // the symmetric sweep reads rows in blocks of up to 1024: pad the vector accordingly
    long v_pad = (n + 1023) / 1024 * 1024;
    int rc = ensure_vpad(ctx, v_pad);
    if (rc) return rc;
    (void)n_pad;
    // fixed summation order: one copy of y per CTA slot (zeroed here), summed in slot order after the sweep
    const int nslots = ctx->num_sms;
    rc = ensure_ypart(ctx, (long)nslots * v_pad);
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * v_pad, st));
    kmv_prologue_kernel<<<(unsigned)((v_pad + 255) / 256), 256, 0, st>>>(v, n, v_pad, ctx->vpad, nullptr, 0, 0.0);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    SweepArgs a{};
    a.xp_rows = xp; a.xp_cols = xp; a.vcol = ctx->vpad; a.ucol = nullptr; a.y = ctx->ypart; a.ystride = v_pad; a.gout = nullptr;
    a.nrows = n; a.ncols = n; a.exp_tab = ctx->exp_table; a.variance = variance; a.part = part; a.nparts = nparts;
    rc = dispatch(ctx, kind, d, 0, a, st);
    if (rc) return rc;
    reduce_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->ypart, v_pad, nslots, v, part == 0 ? diag : 0.0, y, n);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_kmv_sym_multi(cglb_context* c, int kind, const double* xp, long n, int d, const double* v, int t, double* y,
                                  double variance, double diag, int part, int nparts, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, "part/nparts");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    CGLB_CHECK_ARG(t >= 1, "t >= 1 right-hand sides");
    if (n == 0) return CGLB_OK;
    CGLB_CHECK_ARG(xp && v && y, "null pointer");
    if (t == 1) return cglb_kmv_sym(c, kind, xp, n, d, v, y, variance, diag, part, nparts, stream);
    if (d > CGLB_MAX_REGISTER_D) {
        set_error("cglb_kmv_sym_multi: d=%d > %d (wide inputs): call cglb_kmv_sym once per right-hand side", d, CGLB_MAX_REGISTER_D);
        return CGLB_ERR_UNSUPPORTED;
    }
    sweep_fn f = get_sweep_fn(d);
    if (!f) {
        set_error("kernel sweep: d=%d has no register-resident instantiation in this build", d);
        return CGLB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long v_pad = (n + 1023) / 1024 * 1024;
    const int nslots = ctx->num_sms;
    // columns in groups of 4 (or 2 for a last pair): one sweep per group, every kernel pair evaluated once per group
    for (int t0 = 0; t0 < t;) {
        const int T = (t - t0 > 2) ? 4 : 2;
        int rc = ensure_vpad(ctx, v_pad * T);
        if (rc) return rc;
        rc = ensure_ypart(ctx, (long)nslots * v_pad * T);
        if (rc) return rc;
        CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * v_pad * T, st));
        multi_prologue_kernel<<<(unsigned)((v_pad * T + 255) / 256), 256, 0, st>>>(v, n, v_pad, t, t0, T, ctx->vpad);
        ctx->launches++;
        CGLB_LAUNCH_OK();
        SweepArgs a{};
        a.xp_rows = xp; a.xp_cols = xp; a.vcol = ctx->vpad; a.y = ctx->ypart; a.ystride = v_pad * T;
        a.nrows = n; a.ncols = n; a.exp_tab = ctx->exp_table; a.variance = variance; a.part = part; a.nparts = nparts;
        rc = f(ctx, kind, T == 2 ? 5 : 6, a, st);
        if (rc) return rc;
        multi_reduce_kernel<<<(unsigned)((n * T + 255) / 256), 256, 0, st>>>(ctx->ypart, v_pad * T, nslots, v, part == 0 ? diag : 0.0, y, n, t,
                                                                               t0, T);
        ctx->launches++;
        CGLB_LAUNCH_OK();
        t0 += T;
    }
    return CGLB_OK;
}

extern "C" int cglb_kmv_rect(cglb_context* c, int kind, const double* xp_rows, long nrows, const double* xp_cols,
                             long ncols, int d, const double* v, double* y, double variance, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    if (nrows == 0) return CGLB_OK;
    CGLB_CHECK_ARG(xp_rows && y && (ncols == 0 || (xp_cols && v)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    long v_pad = (ncols + 1023) / 1024 * 1024;
    int rc = ensure_vpad(ctx, v_pad);
    if (rc) return rc;
    kmv_prologue_kernel<<<(unsigned)((v_pad + 255) / 256), 256, 0, st>>>(v, ncols, v_pad, ctx->vpad, nullptr, 0, 0.0);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    if (ncols == 0) {
        CGLB_CUDA_OK(cudaMemsetAsync(y, 0, sizeof(double) * nrows, st));
        return CGLB_OK;
    }
    const int nslots = ctx->num_sms;
    const long r_pad = (nrows + 7) / 8 * 8;
    rc = ensure_ypart(ctx, (long)nslots * r_pad);
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * r_pad, st));
    SweepArgs a{};
    a.xp_rows = xp_rows; a.xp_cols = xp_cols; a.vcol = ctx->vpad; a.y = ctx->ypart; a.ystride = r_pad;
    a.nrows = nrows; a.ncols = ncols; a.exp_tab = ctx->exp_table; a.variance = variance; a.part = 0; a.nparts = 1;
    rc = dispatch(ctx, kind, d, 1, a, st);
    if (rc) return rc;
    reduce_parts_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(ctx->ypart, r_pad, nslots, nullptr, 0.0, y, nrows);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_kmv_bwd_sym(cglb_context* c, int kind, const double* xp, long n, int d, const double* u,
                                const double* w, double variance, const double* lengthscale, double* out, int part,
                                int nparts, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx != nullptr, "null context");
    CGLB_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, "part/nparts");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    if (n == 0) return CGLB_OK;
    CGLB_CHECK_ARG(xp && u && w && out && lengthscale, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    long v_pad = (n + 1023) / 1024 * 1024;
    int rc = ensure_vpad(ctx, v_pad);
    if (rc) return rc;
    CGLB_CHECK_ARG(d + 1 <= kScratchScalars, "d too large");
    rc = ensure_scratch(ctx, kScratchScalars);
    if (rc) return rc;
    bwd_prologue_kernel<<<(unsigned)((v_pad + 255) / 256), 256, 0, st>>>(u, w, n, v_pad, ctx->upad, ctx->vpad, ctx->rsum);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    const int nslots = ctx->num_sms;
    rc = ensure_ypart(ctx, (long)nslots * (v_pad + kScratchScalars));
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(ctx->ypart, 0, sizeof(double) * nslots * (v_pad + kScratchScalars), st));
    double* gpart = ctx->ypart + (long)nslots * v_pad;
    SweepArgs a{};
    a.xp_rows = xp; a.xp_cols = xp; a.vcol = ctx->vpad; a.ucol = ctx->upad; a.y = ctx->ypart; a.ystride = v_pad;
    a.gout = gpart; a.gstride = kScratchScalars;
    a.nrows = n; a.ncols = n; a.exp_tab = ctx->exp_table; a.variance = variance; a.part = part; a.nparts = nparts;
    rc = dispatch(ctx, kind, d, 2, a, st);
    if (rc) return rc;
    reduce_bwd_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->ypart, v_pad, nslots, ctx->rsum, n, gpart, kScratchScalars, d,
                                                                          ctx->scratch);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    // dk/dl_q = variance * e' * delta_q^2 / l_q with e' = e^-s (Matern32) or 2 e^-q (RBF)
    const double cfac = (kind == CGLB_MATERN32) ? 1.0 : 2.0;
    bwd_epilogue_kernel<<<d, 256, 0, st>>>(xp, n, d, packed_width(d), ctx->rsum, ctx->scratch, lengthscale, variance, cfac, out, 1);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_knm_build(cglb_context* c, int kind, const double* zp, long m, const double* xp, long n, int d,
                              double variance, double* out, long ld, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && zp && xp && out, "null pointer");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    CGLB_CHECK_ARG(ld >= n, "ld >= n");
    if (d > CGLB_MAX_REGISTER_D) return knm_build_wide(ctx, kind, zp, m, xp, n, d, variance, out, ld, (cudaStream_t)stream);
    knm_fn f = get_knm_fn(d);
    if (!f) {
        set_error("knm_build: d=%d has no instantiation in this build", d);
        return CGLB_ERR_UNSUPPORTED;
    }
    KnmArgs a{};
    a.zp = zp; a.m = m; a.xp = xp; a.ncols = n; a.out = out; a.ld = ld; a.exp_tab = ctx->exp_table; a.variance = variance;
    return f(ctx, kind, 0, a, (cudaStream_t)stream);
}

extern "C" int cglb_knm_backward(cglb_context* c, int kind, const double* zp, long m, const double* xp, long ncols, int d,
                                 double variance, const double* lengthscale, const double* t, long ldt, const double* wt,
                                 const double* zvec, double* out_ls, double* out_var, double* out_z, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && zp && xp && lengthscale && out_ls && out_var, "null pointer");
    CGLB_CHECK_ARG(kind == CGLB_MATERN32 || kind == CGLB_RBF, "kernel kind");
    CGLB_CHECK_ARG((wt == nullptr) == (zvec == nullptr), "wt and zvec go together");
    if (d > CGLB_MAX_REGISTER_D)
        return knm_backward_wide(ctx, kind, zp, m, xp, ncols, d, variance, lengthscale, const_cast<double*>(t), ldt, wt, zvec, out_ls,
                                 out_var, out_z, (cudaStream_t)stream);
    knm_fn f = get_knm_fn(d);
    if (!f) {
        set_error("knm_backward: d=%d has no instantiation in this build", d);
        return CGLB_ERR_UNSUPPORTED;
    }
    KnmArgs a{};
    a.zp = zp; a.m = m; a.xp = xp; a.ncols = ncols; a.exp_tab = ctx->exp_table; a.variance = variance;
    a.t = t; a.ldt = ldt; a.wt = wt; a.zvec = zvec; a.lengthscale = lengthscale;
    a.out_ls = out_ls; a.out_var = out_var; a.out_z = out_z;
    a.cscale = (kind == CGLB_MATERN32) ? 1.7320508075688772935 : 0.70710678118654752440;
    a.cfac = (kind == CGLB_MATERN32) ? 1.0 : 2.0;
    return f(ctx, kind, 1, a, (cudaStream_t)stream);
}
