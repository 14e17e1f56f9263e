/* cglb_b200 -- C-ABI of the B200-native (sm_100a) CGLB hot path.
 *
 * The reference (awav/CGLB) is pure Python and has no FFI of its own: on its hot path every device
 * operation is a third-party call (KeOps JIT reductions, cuBLAS, cuSOLVER/MAGMA through torch; SURVEY.md
 * section 2.2).  Each entry point below replaces one of those call sites and cites it as
 * "replaces: <file>:<line>" relative to the reference root.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative cglb_status otherwise; cglb_last_error() gives text.
 *     Nothing throws, nothing calls exit().
 *   - all buffers are caller-owned DEVICE pointers, fp64, row-major, contiguous unless a leading
 *     dimension is given.  `stream` is a cudaStream_t passed as void* (torch: current_stream().cuda_stream).
 *   - every call is asynchronous on `stream`; scalar results are written to device memory.
 *   - a cglb_context owns small workspaces for one device; it is not thread-safe.
 *   - "packed" inputs: see cglb_pack_inputs.
 */
#ifndef CGLB_B200_H
#define CGLB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CGLB_API __attribute__((visibility("default")))
#else
#define CGLB_API
#endif

#define CGLB_ABI_VERSION 1
#define CGLB_ROW_PAD 128 /* packed arrays are padded with zero rows to a multiple of this */

enum cglb_status {
    CGLB_OK = 0,
    CGLB_ERR_ARG = -1,
    CGLB_ERR_CUDA = -2,
    CGLB_ERR_NOMEM = -3,
    CGLB_ERR_UNSUPPORTED = -4,
    CGLB_ERR_NOT_POSDEF = -5
};

/* gpytorch.kernels[.keops].MaternKernel(nu=1.5) / RBFKernel under ScaleKernel, ARD lengthscales;
 * built at cglb/backend/pytorch/interface.py:207-230 */
enum cglb_kernel_kind { CGLB_MATERN32 = 0, CGLB_RBF = 1 };

typedef struct cglb_context cglb_context;

CGLB_API int cglb_abi_version(void);
CGLB_API const char* cglb_last_error(void);
CGLB_API int cglb_create(cglb_context** ctx, int device);
CGLB_API int cglb_destroy(cglb_context* ctx);
/* Developer options of a context (not part of the reference's interface: they select between kernels that compute the
 * same thing).  "dsweep": 0 register-resident sweeps only, 1 the size / dimension policy (default), 2 DMMA sweeps wherever the
 * packed width allows; "gemm_staging": 1 cp.async operand ring (default), 2 TMA bulk-copy ring for 16-byte aligned operands;
 * "superrow": column chunks per super-row of the DMMA sweeps' L2-blocked item order, 0 = sized for the L2 (default).
 * Initial values come from the environment variables CGLB_DSWEEP / CGLB_GEMM_STAGING, read once in cglb_create. */
CGLB_API int cglb_set_option(cglb_context* ctx, const char* name, long value);

/* number of kernels launched through this context so far (bench.py "gpu_launches") */
CGLB_API unsigned long long cglb_launch_count(const cglb_context* ctx);
CGLB_API int cglb_num_sms(const cglb_context* ctx);

/* ---- packed inputs ------------------------------------------------------------------------------
 * xp[i] = { c (x[i][q]-shift[q]) / lengthscale[q]  (q<d), 0 padding, |.|^2 in the last slot },
 * width cglb_packed_width(d) doubles, cglb_padded_rows(n) rows (rows >= n are zero).
 * c = sqrt(3) for Matern32, 1/sqrt(2) for RBF so that the kernels see s^2 resp. r^2/2 directly.
 * replaces: the x/lengthscale scaling inside gpytorch kernels (interface.py:207-230), done once per
 * hyper-parameter value instead of once per K*v. */
CGLB_API int cglb_packed_width(int d);
CGLB_API long cglb_padded_rows(long n);
CGLB_API int cglb_pack_inputs(cglb_context* ctx, int kind, const double* x, long n, int d, const double* lengthscale,
                     const double* shift, double* xp, void* stream);

/* ---- K1: matrix-free kernel matvec -------------------------------------------------------------
 * y = variance * K(X,X) v + diag * v   (symmetric sweep: each unordered tile pair is evaluated once).
 * With nparts > 1 only work items {t : t % nparts == part} are processed and y receives the partial
 * sum of this part (the diag term is added by part 0); the caller all-reduces y.
 * replaces: `A @ p`, `A @ v` at cglb/backend/pytorch/conjugate_gradient.py:57,66,72 and `cov @ v` at
 * cglb/backend/pytorch/models.py:280 (KeOps Genred sum-reduction behind gpytorch KeOpsLazyTensor). */
CGLB_API int cglb_kmv_sym(cglb_context* ctx, int kind, const double* xp, long n, int d, const double* v, double* y,
                 double variance, double diag, int part, int nparts, void* stream);

/* The same product against a block of t right-hand sides: Y = variance * K(X,X) V + diag * V with V, Y [n][t] row-major
 * (the layout of the reference's [N, t] tensors).  Columns are processed in groups of 4 (2 for a last pair); within a
 * group every kernel pair is evaluated ONCE and used for all its accumulations (3d + 5 + 2t FLOPs per pair instead of
 * t (3d + 7)).  d <= 32; t = 1 forwards to cglb_kmv_sym.  part / nparts as in cglb_kmv_sym.
 * replaces: `A @ x` for x of shape [N, t] -- the operator protocol of cglb/backend/pytorch/conjugate_gradient.py:57,66,72
 * ("b: [N, t]" in its docstring) and cglb/backend/pytorch/models.py:280. */
CGLB_API int cglb_kmv_sym_multi(cglb_context* ctx, int kind, const double* xp, long n, int d, const double* v, int t, double* y,
                       double variance, double diag, int part, int nparts, void* stream);

/* Which kernel cglb_kmv_sym launches for this shape: 0 = register-resident DFMA sweep (kmv_impl.cuh),
 * 1 = DMMA-distance sweep for 10 <= d <= 32 (dsweep_impl.cuh), 2 = wide DMMA sweep for d > 32 (widek.cu).
 * Pure query (no launch); benchmarks use it to name the kernel they time. */
CGLB_API int cglb_kmv_sym_variant(const cglb_context* ctx, int d, long n, int nparts);

/* ---- K1 in fp32-pair mode (the reference's fp32 switch, cglb/backend/pytorch/interface.py:96-110) ----------
 * Same product as cglb_kmv_sym with the n^2 kernel-pair evaluations in FP32 (FP32 FMA pipe + MUFU rsqrt/ex2);
 * v, y and all accumulation across tiles stay FP64.  Per-entry error ~1e-6 (expanded-form distances in
 * fp32), d <= 32.  xpf comes from cglb_pack_inputs_f32: width cglb_packed_width_f32(d) floats,
 * cglb_padded_rows(n) rows (coordinates additionally scaled by log2(e) resp. sqrt(log2(e)): opaque to callers).
 * replaces: the same call sites as cglb_kmv_sym when the model was created under set_default_float("fp32"). */
CGLB_API int cglb_packed_width_f32(int d);
CGLB_API int cglb_pack_inputs_f32(cglb_context* ctx, int kind, const double* x, long n, int d, const double* lengthscale,
                         const double* shift, float* xpf, void* stream);
CGLB_API int cglb_kmv_sym_f32(cglb_context* ctx, int kind, const float* xpf, long n, int d, const double* v, double* y,
                     double variance, double diag, int part, int nparts, void* stream);
/* K2 in fp32-pair mode: same contract as cglb_kmv_bwd_sym; the sweep reads xpf, the O(n d) epilogue reads the
 * fp64 packed array xp of the same inputs. */
CGLB_API int cglb_kmv_bwd_sym_f32(cglb_context* ctx, int kind, const float* xpf, const double* xp, long n, int d,
                         const double* u, const double* w, double variance, const double* lengthscale, double* out,
                         int part, int nparts, void* stream);

/* y[0:nrows] = variance * K(rows, cols) v[0:ncols]   (rectangular, e.g. K_sf v of PredictCG)
 * replaces: `ksf @ new_v` at cglb/backend/pytorch/models.py:334 */
CGLB_API int cglb_kmv_rect(cglb_context* ctx, int kind, const double* xp_rows, long nrows, const double* xp_cols,
                  long ncols, int d, const double* v, double* y, double variance, void* stream);

/* ---- K2: fused backward sweep -------------------------------------------------------------------
 * out[q] (q<d) += sum_ij u_i w_j dK_ij/dlengthscale_q ; out[d] += sum_ij u_i w_j dK_ij/dvariance,
 * one matrix-free sweep over unordered tile pairs.  out is accumulated into (caller zeroes it);
 * lengthscale is the same device array given to cglb_pack_inputs.
 * replaces: KeOps autograd of `cov @ v` (graph built at models.py:280, differentiated at
 * cglb/backend/pytorch/optimizer.py:97). */
CGLB_API int cglb_kmv_bwd_sym(cglb_context* ctx, int kind, const double* xp, long n, int d, const double* u,
                     const double* w, double variance, const double* lengthscale, double* out, int part,
                     int nparts, void* stream);

/* ---- K7: dense cross-covariance ------------------------------------------------------------------
 * out[m][i] = variance * k(z_m, x_i), out is M x ld row-major (ld >= n).
 * replaces: delazify(kernel(Z, X)), kernel(Z, Z) at cglb/backend/pytorch/models.py:196-201 */
CGLB_API int cglb_knm_build(cglb_context* ctx, int kind, const double* zp, long m, const double* xp, long n, int d,
                   double variance, double* out, long ld, void* stream);

/* ---- K5/K6: dense M-sized linear algebra ---------------------------------------------------------
 * cglb_potrf: in-place lower Cholesky of the m x m matrix a (upper triangle is zeroed);
 *   *info_dev = 0 on success, else 1+index of the first non-positive pivot block (device int).
 *   replaces: torch.cholesky at models.py:202,210 (cuSOLVER/MAGMA dpotrf)
 * cglb_tri_inverse: linv = l^-1 (lower triangular, m x m, out of place).
 * cglb_trsm_left_lower: b <- alpha * l^-1 b, b is m x n (ldb), blocked with inverted diagonal blocks.
 *   replaces: trisolve(kuf, kuu_chol, upper=False) at models.py:206 (cuBLAS dtrsm)
 * cglb_syrk: c (m x m, full symmetric) = [c +] a a^T, a is m x n (lda)   (split-K, atomics)
 *   replaces: A @ A.transpose(-1,-2) at models.py:207 (cuBLAS dgemm)
 * cglb_gemm: c = alpha * a * op(b) + beta * c ; a is m x k (lda); b is k x n (transb=0) or n x k (transb=1)
 */
CGLB_API int cglb_potrf(cglb_context* ctx, double* a, long m, long lda, int* info_dev, void* stream);
CGLB_API int cglb_tri_inverse(cglb_context* ctx, const double* l, long m, long ldl, double* linv, long ldi, void* stream);
CGLB_API int cglb_trsm_left_lower(cglb_context* ctx, const double* l, long m, long ldl, double* b, long n, long ldb,
                         double alpha, void* stream);
CGLB_API int cglb_syrk(cglb_context* ctx, const double* a, long m, long n, long lda, double* c, long ldc, int accumulate,
              void* stream);
CGLB_API int cglb_gemm(cglb_context* ctx, int transb, long m, long n, long k, double alpha, const double* a, long lda,
              const double* b, long ldb, double beta, double* c, long ldc, void* stream);

/* ---- K3/K4: preconditioner apply ------------------------------------------------------------------
 * NystromPreconditioner (cglb/backend/pytorch/conjugate_gradient.py:89-113) in two halves so that a
 * row-sharded caller can all-reduce q in between:
 *   cglb_precond_project:  q[M] = A[:, cols] r[cols]          (HBM-streaming GEMV, :105)
 *   cglb_precond_finish:   w = LB^-T LB^-1 q                   (:106-107, through lbinv = LB^-1)
 *                          z[cols] = (r[cols] - A[:, cols]^T w) / sigma_sq   (:110-113)
 *                          *rz_dev (+)= sum z[cols] r[cols]
 * A is M x ncols (lda), lbinv is M x M lower triangular (from cglb_tri_inverse). */
CGLB_API int cglb_precond_project(cglb_context* ctx, const double* a, long m, long ncols, long lda, const double* r,
                         double* q, void* stream);
CGLB_API int cglb_precond_finish(cglb_context* ctx, const double* a, long m, long ncols, long lda, const double* lbinv,
                        const double* q, const double* r, double sigma_sq, double* z, double* w_out,
                        double* rz_dev, void* stream);

/* ---- K8: fused CG vector updates and dot products --------------------------------------------------
 * replaces the ~10 torch elementwise/reduce launches per iteration at conjugate_gradient.py:58,67-75.
 * All scalars live in device memory (no host sync inside).
 *   cglb_dot:          *out = sum x_i y_i   (deterministic two-stage reduction)
 *   cglb_cg_step:      gamma = *rz / *pAp ; v += gamma p ; if (!restart) r -= gamma Ap     (:67-72)
 *   cglb_residual:     r = b - Av                                                          (:58,72)
 *   cglb_cg_direction: p = restart ? z : z + p * (*rz_new / *rz_old)                      (:75)
 *   cglb_quad_terms:   out[0] = sum v (r + 0.5 Kv)   with r = err - Kv written to r      (models.py:281-283)
 */
CGLB_API int cglb_dot(cglb_context* ctx, const double* x, const double* y, long n, double* out_dev, void* stream);
CGLB_API int cglb_cg_step(cglb_context* ctx, long n, const double* rz_dev, const double* pAp_dev, const double* p,
                 const double* Ap, double* v, double* r, int restart, void* stream);
CGLB_API int cglb_residual(cglb_context* ctx, long n, const double* b, const double* Av, double* r, void* stream);
CGLB_API int cglb_cg_direction(cglb_context* ctx, long n, const double* z, double* p, const double* rz_new_dev,
                      const double* rz_old_dev, int restart, void* stream);
CGLB_API int cglb_quad_terms(cglb_context* ctx, long n, const double* err, const double* Kv, const double* v, double* r,
                    double* out_dev, void* stream);

/* ---- backward of the M x n cross-covariance -----------------------------------------
 * Given G = dS/dK_uf (M x ncols, ld), accumulate
 *   out_ls[q]  += sum_mi G_mi dk(z_m,x_i)/dlengthscale_q,  out_var += sum_mi G_mi k/variance,
 *   out_z[m][q] += sum_i G_mi dk(z_m,x_i)/dz_mq
 * with G_mi = t[m][i] + wt[m] * zvec[i] (t may be NULL => only the rank-one part).
 * replaces: torch autograd through kernel(Z, X) (models.py:196-197). */
CGLB_API int cglb_knm_backward(cglb_context* ctx, int kind, const double* zp, long m, const double* xp, long ncols, int d,
                      double variance, const double* lengthscale, const double* t, long ldt, const double* wt,
                      const double* zvec, double* out_ls, double* out_var, double* out_z, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGLB_B200_H */
