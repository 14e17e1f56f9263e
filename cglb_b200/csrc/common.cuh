// Shared device/host helpers for the cglb_b200 sm_100a library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cglb_b200.h"

namespace cglb {

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns int, never throws
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define CGLB_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cglb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CGLB_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

#define CGLB_CHECK_ARG(cond, msg)                                                           \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            cglb::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);       \
            return CGLB_ERR_ARG;                                                            \
        }                                                                                   \
    } while (0)

#define CGLB_LAUNCH_OK()                                                                    \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            cglb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CGLB_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

// Context: one per (device, stream user).  Owns small workspaces so that the entry points can stay
// allocation-free on the hot path.
struct Context {
    int device;
    int num_sms;
    // arrival counter of the two-stage deterministic reductions (vecops.cu: the last block sums in fixed order)
    int* counters;        // [16]
    // padded copies of the vector operands of the sweeps
    double* vpad;         // [vpad_cap]
    double* upad;         // [vpad_cap]
    double* rsum;         // [vpad_cap]  row-sum scratch of the backward sweep
    long vpad_cap;
    double* scratch;      // generic scratch (partials of reductions)
    long scratch_cap;
    // deterministic accumulation (DESIGN.md "fixed-order reductions"): every CTA of a sweep adds into its OWN copy of the
    // output vector (ypart + blockIdx.x * stride), a second kernel sums the copies in CTA order
    double* ypart;        // [ypart_cap]
    long ypart_cap;
    double* exp_table;    // [64] 2^(j/64) followed by [1024] 2^(j/1024) with j << 10 subtracted from the high word
                          // (fast_exp_neg<10>); staged into shared memory by the kernels
    unsigned long long launches;   // number of kernels this context launched (bench.py gpu_launches)
    // developer options (cglb_set_option; initialised from the environment once, at cglb_create)
    int opt_dsweep;        // 0: register-resident sweeps only, 1: size/dimension policy (default), 2: DMMA sweeps wherever possible
    int opt_gemm_staging;  // 1: cp.async ring (default), 2: TMA bulk-copy ring for aligned operands
    long opt_superrow;     // DMMA sweeps: column chunks per super-row of the item order, 0 = sized for the L2 (default)
};

// the first kScratchScalars doubles of Context::scratch hold the scalar accumulators of the sweeps
// (d + 1 gradient sums, d <= 128); everything else starts after them
constexpr int kScratchScalars = 256;
int ensure_vpad(Context* ctx, long n_pad);
int ensure_scratch(Context* ctx, long n_doubles);
int ensure_ypart(Context* ctx, long n_doubles);

// ---------------------------------------------------------------------------------------------
// packed input layout (see DESIGN.md "data layout")
//   row i of a packed array = DP doubles: scaled+centred coordinates a_0..a_{d-1}, zero padding,
//   and the squared norm |a|^2 in the LAST slot.  DP = d+1 rounded up to even so that a row is a
//   multiple of 16 bytes (TMA bulk copies and LDS.128 broadcasts need that).
//   Rows are padded to a multiple of CGLB_ROW_PAD with zeros.
// ---------------------------------------------------------------------------------------------
//   d > CGLB_MAX_REGISTER_D ("wide" layout for the DMMA sweeps): coordinates zero-padded to KP = d rounded up
//   to a multiple of 4, |a|^2 at index KP, row width W = KP + 4 or KP + 8 chosen so that W mod 16 is 4 or 12
//   (conflict-free DMMA fragment loads straight from the TMA-landed tile).
#define CGLB_MAX_REGISTER_D 32
__host__ __device__ inline int wide_kp(int d) { return (d + 3) & ~3; }
__host__ __device__ inline int packed_width(int d) {
    if (d <= CGLB_MAX_REGISTER_D) return (d + 2) & ~1;
    const int kp = wide_kp(d);
    return kp + ((kp % 8 == 0) ? 4 : 8);
}
__host__ __device__ inline int norm_index(int d) { return d <= CGLB_MAX_REGISTER_D ? packed_width(d) - 1 : wide_kp(d); }
__host__ __device__ inline long padded_rows(long n) { return (n + CGLB_ROW_PAD - 1) / CGLB_ROW_PAD * CGLB_ROW_PAD; }

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy wrappers (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// FP64 math tuned for the sweeps (DESIGN.md "kernel arithmetic"): the libm sqrt+exp pair costs
// ~29 FP64 issue slots; these cost 5 + 9.
// ---------------------------------------------------------------------------------------------

// sqrt(q) for q in [2^-1000, 2^60] (the callers clamp q with integer min/max on its high word, kmv_impl.cuh): MUFU.RSQ64H seed (rel err 2^-20, measured) + one third-order
// correction: s = g(1 + e/2 + 3e^2/8), e = 1 - q y^2  -> rel err ~ (5/16) e^3 < 1e-18.  5 FP64 slots.
__device__ __forceinline__ double fast_sqrt(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    double g = q * y;
    double e = fma(-g, y, 1.0);
    double p = fma(e, 0.375, 0.5);
    double t = e * p;
    return fma(g, t, g);
}

// exp(-s) for 0 <= s <= 693: n = rint(-s*2^TB/ln2) by the magic-number add, r = -s - n ln2/2^TB, e^r by a short
// polynomial, 2^(n/2^TB) from a 2^TB-entry table in shared memory and an exponent-field add.
//   TB = 6  : 64-entry table, degree-5 polynomial, 9 FP64 slots  (small kernels: 512 B of shared memory)
//   TB = 10 : 1024-entry table (8 KB), degree-3 polynomial with the r^4/24 term folded into the quadratic
//             coefficient (Chebyshev), 7 FP64 slots, max rel. error 9.4e-17 + rounding    (the sweeps;
//             checked in 50-digit arithmetic by tests/test_kernel_arithmetic.py)
// ln2/2^TB is a single double: its representation error (3.3e-17 relative) adds s * 3.3e-17 to the relative error of
// e^-s (1e-15 at s = 30) -- next to the s * 1.1e-16 .. 2.2e-16 that the rounding of s itself costs any fp64
// evaluation (the whole map in emulated fp64 against 50-digit arithmetic: tests/test_kernel_arithmetic.py).
// The caller clamps s to [0, 693] (kappa() does it on the squared distance with two integer min/max), so
// e^-s >= 2^-1000 and the exponent-field add cannot wrap: no separate exponent clamp.
// Instruction diet (profiles/dsweep_ncu_r01.md, profiles/fp64_issue_model_r01.txt): every non-FP64 instruction costs the FP64 pipe ~1 issue cycle in
// these kernels, and a DFMA reading three distinct vector registers issues at 2/3 rate (tools/fp64_issue_model.cu);
// hence e^r = T * (1 + r p) as DFMA(2 regs + imm) + DMUL and the exponent insert as LOP3 + IMAD.
constexpr int kExpTabSmall = 64;
constexpr int kExpTabBig = 1024;
template <int TB>
__device__ __forceinline__ double fast_exp_neg(double s, const double* __restrict__ tab /* shared */) {
    const double MAGIC = 6755399441055744.0;              // 1.5 * 2^52
    constexpr double C = (TB == 6) ? 92.332482616893656820 : 1477.3197218702985091;     // 2^TB / ln 2
    constexpr double L = (TB == 6) ? 1.0830424696249145255e-02 : 6.7690154351557157843e-04;   // ln 2 / 2^TB
    double t = fma(s, -C, MAGIC);
    int n = __double2loint(t);
    double nf = t - MAGIC;
    double r = fma(nf, -L, -s);
    double p;
    if (TB == 6) {
        p = fma(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
        p = fma(p, r, 1.6666666666666665741e-01);
        p = fma(p, r, 0.5);
        p = fma(p, r, 1.0);
    } else {
        p = fma(r, 1.6666666666666665741e-01, 0.5 + 3.9540e-09);      // 1/2 + (sqrt2 - 1) h^2 / 12, h = ln2/2048: the
                                                                       // r^4/24 term equioscillates, max rel. err 9.4e-17
        p = fma(p, r, 1.0);
    }
    double T = tab[n & ((1 << TB) - 1)];
    if (TB == 10) {
        // The big table stores 2^(j/1024) with j << 10 SUBTRACTED from the high word (context.cu), so that one IMAD
        // n * 2^10 + hi(T') = hi(2^(j/1024)) + ((n >> 10) << 20) forms 2^(n/1024) without masking n first (n = (n >> 10) 2^10
        // + j): one LOP3 less per kernel pair.  Scaling by a power of two commutes with the rounding of the product, so the
        // result has the same bits as "multiply, then insert the exponent".
        int hi;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(n), "n"(1 << (20 - TB)), "r"(__double2hiint(T)));
        return __hiloint2double(hi, __double2loint(T)) * fma(r, p, 1.0);
    }
    double res = T * fma(r, p, 1.0);
    int hi;      // hi(res) + ((n >> TB) << 20)
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(n & ~((1 << TB) - 1)), "n"(1 << (20 - TB)), "r"(__double2hiint(res)));
    return __hiloint2double(hi, __double2loint(res));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cglb
