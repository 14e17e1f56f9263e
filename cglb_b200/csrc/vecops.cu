// K3 / K4 / K8: HBM-streaming GEMV pair of the Nystrom preconditioner, triangular applies, and the fused
// CG vector updates / dot products.  Replaces reference conjugate_gradient.py:58,67-75,105-113 (cuBLAS
// dgemv/dtrsv + ~10 torch elementwise launches per iteration).  All scalars stay in device memory.
#include "common.cuh"

namespace cglb {

constexpr int kRedBlocks = 592;      // 4 x 148: fixed grid of the deterministic reductions
constexpr int kRedThreads = 256;

// ---------------------------------------------------------------------------------------------
// deterministic reduction: every block writes its partial, the last block to finish sums the
// partials in index order (so the result does not depend on scheduling).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (warp == 0) {
        s = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
        s = warp_sum(s);
    }
    __syncthreads();
    return s;   // valid in warp 0
}

__device__ __forceinline__ void finish_reduction(double partial, double* partials, int* counter, double* out,
                                                 double scale, double* sh) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = partial;
        __threadfence();
        int done = atomicAdd(counter, 1);
        s_last = (done == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += partials[i];   // fixed assignment
        s = block_sum(s, sh);
        if (threadIdx.x == 0) {
            *out = scale * s;
            *counter = 0;
        }
    }
}

__global__ void __launch_bounds__(kRedThreads) dot_kernel(const double* __restrict__ x, const double* __restrict__ y, long n,
                                                          double* partials, int* counter, double* out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s = fma(x[i], y[i], s);
    s = block_sum(s, sh);
    finish_reduction(s, partials, counter, out, 1.0, sh);
}

// r = err - Kv ; out = sum v (r + 0.5 Kv)
__global__ void __launch_bounds__(kRedThreads) quad_terms_kernel(const double* __restrict__ err, const double* __restrict__ kv,
                                                                 const double* __restrict__ v, double* __restrict__ r, long n,
                                                                 double* partials, int* counter, double* out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double k = kv[i];
        const double ri = err[i] - k;
        r[i] = ri;
        s = fma(v[i], fma(0.5, k, ri), s);
    }
    s = block_sum(s, sh);
    finish_reduction(s, partials, counter, out, 1.0, sh);
}

__global__ void cg_step_kernel(long n, const double* __restrict__ rz, const double* __restrict__ pAp,
                               const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ v,
                               double* __restrict__ r, int restart) {
    const double gamma = *rz / *pAp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        v[i] = fma(gamma, p[i], v[i]);
        if (!restart) r[i] = fma(-gamma, Ap[i], r[i]);
    }
}

__global__ void residual_kernel(long n, const double* __restrict__ b, const double* __restrict__ av, double* __restrict__ r) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) r[i] = b[i] - av[i];
}

__global__ void cg_direction_kernel(long n, const double* __restrict__ z, double* __restrict__ p,
                                    const double* __restrict__ rz_new, const double* __restrict__ rz_old, int restart) {
    const double beta = restart ? 0.0 : (*rz_new / *rz_old);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        p[i] = restart ? z[i] : fma(p[i], beta, z[i]);
}

// ---------------------------------------------------------------------------------------------
// q[m] += sum_{c in chunk} A[m][c] r[c]        grid (col chunks, row groups), 8 warps, one row per warp pass
// ---------------------------------------------------------------------------------------------
constexpr int kGvCols = 4096;      // columns per CTA (r chunk staged in shared memory: 32 KB)
constexpr int kGvRows = 64;        // rows per CTA

__global__ void __launch_bounds__(256) gemv_rows_kernel(const double* __restrict__ a, long m, long ncols, long lda,
                                                        const double* __restrict__ r, double* __restrict__ q) {
    __shared__ __align__(16) double s_r[kGvCols];
    const long c0 = (long)blockIdx.x * kGvCols;
    const int cn = (int)((ncols - c0) < kGvCols ? (ncols - c0) : kGvCols);
    for (int i = threadIdx.x; i < kGvCols; i += blockDim.x) s_r[i] = (i < cn) ? r[c0 + i] : 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row0 = (long)blockIdx.y * kGvRows;
    const bool vec_ok = ((lda & 1) == 0) && ((c0 & 1) == 0) && (((uintptr_t)a & 15) == 0);
    for (int rr = warp; rr < kGvRows; rr += 8) {
        const long row = row0 + rr;
        if (row >= m) break;
        const double* ap = a + row * lda + c0;
        double acc0 = 0.0, acc1 = 0.0;
        if (vec_ok) {
            const int nvec = cn >> 1;     // double2 elements
            const double2* ap2 = reinterpret_cast<const double2*>(ap);
            const double2* sr2 = reinterpret_cast<const double2*>(s_r);
            int j = lane;
            for (; j + 7 * 32 < nvec; j += 8 * 32) {
                double2 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = __ldcs(ap2 + j + u * 32);     // streaming: A is read once per apply
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double2 rv = sr2[j + u * 32];
                    acc0 = fma(x[u].x, rv.x, acc0);
                    acc1 = fma(x[u].y, rv.y, acc1);
                }
            }
            for (; j < nvec; j += 32) {
                const double2 x = __ldcs(ap2 + j);
                const double2 rv = sr2[j];
                acc0 = fma(x.x, rv.x, acc0);
                acc1 = fma(x.y, rv.y, acc1);
            }
            if ((cn & 1) && lane == 0) acc0 = fma(ap[cn - 1], s_r[cn - 1], acc0);
        } else {
            for (int j = lane; j < cn; j += 32) acc0 = fma(ap[j], s_r[j], acc0);
        }
        const double s = warp_sum(acc0 + acc1);
        if (lane == 0) q[(long)blockIdx.x * m + row] = s;      // partial of this column chunk; summed in chunk order below
    }
}

// q[row] = sum over the column chunks, in chunk order (fixed summation order)
__global__ void __launch_bounds__(256) gemv_rows_reduce_kernel(const double* __restrict__ qpart, long nchunks, long m, double* __restrict__ q) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    double s = 0.0;
#pragma unroll 8
    for (long c = 0; c < nchunks; ++c) s += qpart[c * m + row];
    q[row] = s;
}

// t = Linv q  (Linv lower triangular, m x m): one warp per row
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ linv, long m, const double* __restrict__ q,
                                                         double* __restrict__ t, double* __restrict__ zero_a, long zero_a_n,
                                                         double* __restrict__ zero_b) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + warp;
    // side job: clear the accumulators of the following kernels
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (zero_a && gid < zero_a_n) zero_a[gid] = 0.0;
    if (zero_b && gid == 0) *zero_b = 0.0;
    if (row >= m) return;
    const double* lp = linv + row * m;
    double acc = 0.0;
    for (long j = lane; j <= row; j += 32) acc = fma(lp[j], q[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) t[row] = acc;
}

// wpart[rb][c] = sum_{r in row block rb} Linv[r][c] t[r]   for r >= c      grid (col blocks of 128, row blocks of 128);
// the row-block partials are summed in block order by the consumer (gemv_cols_finish_kernel)
__global__ void __launch_bounds__(128) trmv_lower_t_kernel(const double* __restrict__ linv, long m, const double* __restrict__ t,
                                                           double* __restrict__ wpart) {
    const long c = (long)blockIdx.x * 128 + threadIdx.x;
    const long r0 = (long)blockIdx.y * 128;
    if (blockIdx.y < blockIdx.x) return;      // block strictly above the diagonal: all zeros
    __shared__ double s_t[128];
    s_t[threadIdx.x] = (r0 + threadIdx.x < m) ? t[r0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (c >= m) return;
    const long rend = (r0 + 128 < m) ? r0 + 128 : m;
    double acc = 0.0;
    for (long r = r0; r < rend; ++r) acc = fma(linv[r * m + c], s_t[r - r0], acc);
    wpart[(long)blockIdx.y * m + c] = acc;
}

// z[i] = (r[i] - sum_m A[m][i] w[m]) / sigma_sq ; rz += sum z r        thread = 2 columns
__global__ void __launch_bounds__(128) gemv_cols_finish_kernel(const double* __restrict__ a, long m, long ncols, long lda,
                                                               const double* __restrict__ wpart, double* __restrict__ w_out,
                                                               const double* __restrict__ r, double inv_sigma_sq,
                                                               double* __restrict__ z, double* rz, double* partials, int* counter) {
    extern __shared__ __align__(16) double s_w[];      // [m]
    __shared__ double sh[32];
    // w = LBinv^T t: row-block partials of trmv_lower_t_kernel (blocks rb >= c / 128 exist), summed in block order
    const long nrb = (m + 127) / 128;
    for (long i = threadIdx.x; i < m; i += blockDim.x) {
        double s = 0.0;
        for (long rb = i / 128; rb < nrb; ++rb) s += wpart[rb * m + i];
        s_w[i] = s;
        if (blockIdx.x == 0) w_out[i] = s;
    }
    __syncthreads();
    const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const bool vec_ok = ((lda & 1) == 0) && (((uintptr_t)a & 15) == 0);
    double acc0 = 0.0, acc1 = 0.0;
    if (c < ncols) {
        if (vec_ok && c + 1 < ncols) {
            const double* ap = a + c;
            long mm = 0;
            for (; mm + 8 <= m; mm += 8) {
                double2 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = __ldcs(reinterpret_cast<const double2*>(ap + (mm + u) * lda));
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc0 = fma(x[u].x, s_w[mm + u], acc0);
                    acc1 = fma(x[u].y, s_w[mm + u], acc1);
                }
            }
            for (; mm < m; ++mm) {
                const double2 x = __ldcs(reinterpret_cast<const double2*>(ap + mm * lda));
                acc0 = fma(x.x, s_w[mm], acc0);
                acc1 = fma(x.y, s_w[mm], acc1);
            }
        } else {
            for (long mm = 0; mm < m; ++mm) {
                acc0 = fma(a[mm * lda + c], s_w[mm], acc0);
                if (c + 1 < ncols) acc1 = fma(a[mm * lda + c + 1], s_w[mm], acc1);
            }
        }
    }
    double part = 0.0;
    if (c < ncols) {
        const double r0 = r[c];
        const double z0 = (r0 - acc0) * inv_sigma_sq;
        z[c] = z0;
        part = z0 * r0;
        if (c + 1 < ncols) {
            const double r1 = r[c + 1];
            const double z1 = (r1 - acc1) * inv_sigma_sq;
            z[c + 1] = z1;
            part = fma(z1, r1, part);
        }
    }
    part = block_sum(part, sh);
    finish_reduction(part, partials, counter, rz, 1.0, sh);      // block partials summed in block order by the last block
}

static inline int red_blocks(long n) {
    long b = (n + kRedThreads - 1) / kRedThreads;
    if (b < 1) b = 1;
    return (int)(b < kRedBlocks ? b : kRedBlocks);
}

}  // namespace cglb

using namespace cglb;

extern "C" int cglb_dot(cglb_context* c, const double* x, const double* y, long n, double* out_dev, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && x && y && out_dev, "null pointer");
    int rc = ensure_scratch(ctx, kScratchScalars + kRedBlocks);
    if (rc) return rc;
    // partials live in the first kRedBlocks doubles after the 64 reserved scalars... but the dense
    // workspaces reuse that region, so dot products must not be interleaved *inside* a dense call (they are not).
    dot_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(x, y, n, ctx->scratch + kScratchScalars, ctx->counters + 0, out_dev);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_quad_terms(cglb_context* c, long n, const double* err, const double* Kv, const double* v, double* r,
                               double* out_dev, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && err && Kv && v && r && out_dev, "null pointer");
    int rc = ensure_scratch(ctx, kScratchScalars + kRedBlocks);
    if (rc) return rc;
    quad_terms_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(err, Kv, v, r, n, ctx->scratch + kScratchScalars,
                                                                              ctx->counters + 0, out_dev);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_cg_step(cglb_context* c, long n, const double* rz_dev, const double* pAp_dev, const double* p,
                            const double* Ap, double* v, double* r, int restart, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && rz_dev && pAp_dev && p && Ap && v && r, "null pointer");
    cg_step_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(n, rz_dev, pAp_dev, p, Ap, v, r, restart);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_residual(cglb_context* c, long n, const double* b, const double* Av, double* r, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && b && Av && r, "null pointer");
    residual_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(n, b, Av, r);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_cg_direction(cglb_context* c, long n, const double* z, double* p, const double* rz_new_dev,
                                 const double* rz_old_dev, int restart, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && z && p && rz_new_dev && rz_old_dev, "null pointer");
    cg_direction_kernel<<<red_blocks(n), kRedThreads, 0, (cudaStream_t)stream>>>(n, z, p, rz_new_dev, rz_old_dev, restart);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_precond_project(cglb_context* c, const double* a, long m, long ncols, long lda, const double* r,
                                    double* q, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && a && r && q, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return CGLB_OK;
    if (ncols == 0) {
        CGLB_CUDA_OK(cudaMemsetAsync(q, 0, sizeof(double) * m, st));
        return CGLB_OK;
    }
    const long nchunks = (ncols + kGvCols - 1) / kGvCols;
    int rc = ensure_ypart(ctx, nchunks * m);
    if (rc) return rc;
    dim3 grid((unsigned)nchunks, (unsigned)((m + kGvRows - 1) / kGvRows));
    gemv_rows_kernel<<<grid, 256, 0, st>>>(a, m, ncols, lda, r, ctx->ypart);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    gemv_rows_reduce_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ctx->ypart, nchunks, m, q);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_precond_finish(cglb_context* c, const double* a, long m, long ncols, long lda, const double* lbinv,
                                   const double* q, const double* r, double sigma_sq, double* z, double* w_out,
                                   double* rz_dev, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && a && lbinv && q && r && z && w_out && rz_dev, "null pointer");
    CGLB_CHECK_ARG(m * sizeof(double) <= 200 * 1024, "M too large for the shared-memory staged w");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_scratch(ctx, kScratchScalars + kRedBlocks + m);
    if (rc) return rc;
    double* t = ctx->scratch + kScratchScalars + kRedBlocks;
    // fixed-order partials: [row blocks][m] of w = LBinv^T t, then one r^T z partial per CTA of the finish kernel
    const long nrb = (m + 127) / 128;
    const long nfin = ncols > 0 ? (ncols + 255) / 256 : 1;      // a rank without columns still needs w = B^-1 A r (and r^T z = 0)
    rc = ensure_ypart(ctx, nrb * m + nfin + 8);
    if (rc) return rc;
    double* wpart = ctx->ypart;
    double* rzpart = ctx->ypart + nrb * m;
    if (m > 0) {
        // t = LBinv q
        trmv_lower_kernel<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(lbinv, m, q, t, nullptr, 0, nullptr);
        ctx->launches++;
        CGLB_LAUNCH_OK();
        dim3 g2((unsigned)nrb, (unsigned)nrb);
        trmv_lower_t_kernel<<<g2, 128, 0, st>>>(lbinv, m, t, wpart);
        ctx->launches++;
        CGLB_LAUNCH_OK();
    }
    size_t smem = sizeof(double) * (size_t)(m > 0 ? m : 1);
    CGLB_CUDA_OK(cudaFuncSetAttribute(gemv_cols_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    gemv_cols_finish_kernel<<<(unsigned)nfin, 128, smem, st>>>(a, m, ncols, lda, wpart, w_out, r, 1.0 / sigma_sq, z, rz_dev, rzpart,
                                                                ctx->counters + 1);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}
