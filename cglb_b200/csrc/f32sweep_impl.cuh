// K1 in "fp32 pair" mode (the reference's fp32 switch, interface.py:96-110).
//
//   y (fp64) = variance * K(X,X) v + diag * v   with the n^2 kernel-pair evaluations in FP32:
//   squared distances in the expanded form on the FP32 FMA pipe (128 lanes/clk/SM), sqrt and exp from the
//   MUFU unit (rsqrt.approx / ex2.approx, 16 lanes/clk/SM) -- a different roofline from the FP64 sweeps: 2 MUFU
//   + ~18 FP32 + ~3 other instructions per pair, i.e. issue-bound at ~1 instruction per cycle per scheduler.
//   Row sums are kept in FP32 only within one 64-column tile and accumulated in FP64 registers across tiles;
//   column sums are reduced in FP32 inside a warp tile and leave the CTA as FP64 atomics.  Everything outside
//   the two n^2 sweeps (K_nm, Cholesky, CG vectors, preconditioner) stays FP64.
//
// Packed FP32 layout (cglb_pack_inputs_f32): row i = { c' (x_iq - shift_q) / l_q rounded to float (q < d), zero
// padding, |.|^2 of the ROUNDED coordinates in the last slot }, width DPF = d + 1 rounded up to a multiple of 4.
// c' folds the base-2 conversion of the exponential into the inputs: Matern32 c' = sqrt(3) log2(e), so that
// sqrt(q) = s log2(e) feeds MUFU.EX2 directly and kappa = e + ln2 (s' e); RBF c' = sqrt(log2(e) / 2), so that
// e^{-r^2/2} = 2^{-q}.  One FMUL less per pair than scaling inside the loop; MUFU.SQRT replaces RSQ + FMUL.
// Same work decomposition, TMA ring and symmetric-pair trick as kmv_sweep_kernel (kmv_impl.cuh).
#pragma once
#include <stdlib.h>

#include "kmv_impl.cuh"

namespace cglb {

__host__ __device__ inline int packed_width_f32(int d) { return (d + 4) & ~3; }

struct SweepArgsF32 {
    const float* xp;         // packed fp32 rows/cols [n_pad][DPF]
    const float* vcol;       // v rounded to float, padded
    double* y;               // output: CTA b accumulates into y + b * ystride (fixed order, see kmv_impl.cuh)
    long ystride;
    long n;
    long nb;                 // number of BI blocks
    long nitems;
    double variance;
    int part, nparts;
};

// scale of the packed fp32 coordinates relative to the fp64 packing (see the header comment)
__host__ __device__ inline double f32_input_scale(int kind) {
    return kind == CGLB_MATERN32 ? 1.4426950408889634074 : 1.2011224087864498;      // log2(e), sqrt(log2(e))
}

template <int KIND>
__device__ __forceinline__ float kappa_f32(float q) {
    q = fmaxf(q, 0.0f);                          // cancellation of the expanded form / the diagonal
    if (KIND == CGLB_MATERN32) {
        float s, e;                              // s = sqrt(3) r log2(e)
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(q));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-s));
        return fmaf(s * e, 0.69314718055994531f, e);
    } else {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q));
        return e;
    }
}

template <int CG>
__device__ __forceinline__ void col_reduce_f32(float (&c)[CG], int lane) {
    int cnt = CG;
    int off = 16;
#pragma unroll
    for (; cnt > 1; cnt >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int h = 0; h < cnt / 2; ++h) {
            float send = up ? c[h] : c[h + cnt / 2];
            float keep = up ? c[h + cnt / 2] : c[h];
            c[h] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) c[0] += __shfl_xor_sync(0xffffffffu, c[0], off);
}

template <int KIND, int D, int TI, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) f32_sweep_kernel(const SweepArgsF32 args) {
    constexpr int DPF = (D + 4) & ~3;
    constexpr int kThreads = WARPS * 32;
    constexpr int BI = kThreads * TI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);                    // [kStages][kBJ*DPF]
    float* s_v = s_x + kStages * kBJ * DPF;                             // [kStages][kBJ]
    float* s_col = s_v + kStages * kBJ;                                 // [2][WARPS][kBJ]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_col + 2 * WARPS * kBJ);
    uint64_t* s_empty = s_full + kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    // item cursor: t = tau * nparts + part over the unordered block pairs (I <= C)
    struct Cur {
        long tau, I, C, c0; int tile, ntiles; bool valid;
    };
    auto load_item = [&](Cur& c) {
        const long t = c.tau * args.nparts + args.part;
        c.valid = t < args.nitems;
        if (!c.valid) return;
        item_to_blocks_sym(t, c.I, c.C);
        c.c0 = c.C * BI;
        long cend = c.c0 + BI;
        if (cend > args.n) cend = args.n;
        c.ntiles = (int)((cend - c.c0 + kBJ - 1) / kBJ);
        c.tile = 0;
    };
    int pstage = 0, stage = 0;
    uint32_t pphase = 0, phase = 0;
    auto produce = [&](Cur& pc) {
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DPF + kBJ) * sizeof(float)));
        tma_load_1d(s_x + pstage * kBJ * DPF, args.xp + j0 * DPF, kBJ * DPF * sizeof(float), &s_full[pstage]);
        tma_load_1d(s_v + pstage * kBJ, args.vcol + j0, kBJ * sizeof(float), &s_full[pstage]);
        if (++pstage == kStages) { pstage = 0; pphase ^= 1; }
        if (++pc.tile == pc.ntiles) { pc.tau += gridDim.x; load_item(pc); }
    };

    Cur cc;
    cc.tau = blockIdx.x;
    load_item(cc);
    Cur pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kPrefetch; ++i) produce(pc);
    }
    int colbuf = 0;
    const double var = args.variance;

    while (cc.valid) {
        const bool offdiag = cc.I != cc.C;
        const long r0 = cc.I * BI;
        float a2[TI][D], na[TI], vi[TI];
        double racc[TI];
        bool live[TI];
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            const long row = r0 + ti * kThreads + tid;
            live[ti] = row < args.n;
            const float4* src = reinterpret_cast<const float4*>(args.xp + (live[ti] ? row : 0) * DPF);
            float tmp[DPF];
#pragma unroll
            for (int h = 0; h < DPF / 4; ++h) {
                const float4 p = __ldg(src + h);
                tmp[4 * h] = p.x; tmp[4 * h + 1] = p.y; tmp[4 * h + 2] = p.z; tmp[4 * h + 3] = p.w;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) a2[ti][k] = live[ti] ? -2.0f * tmp[k] : 0.0f;
            na[ti] = live[ti] ? tmp[DPF - 1] : 0.0f;
            vi[ti] = live[ti] ? __ldg(args.vcol + row) : 0.0f;
            racc[ti] = 0.0;
        }
        const int ntiles = cc.ntiles;
        const long c0 = cc.c0;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) produce(pc);
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const float* sx = s_x + stage * kBJ * DPF;
            const float* sv = s_v + stage * kBJ;
            float* scol = s_col + (colbuf * WARPS + warp) * kBJ;
            float rt[TI];
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) rt[ti] = 0.0f;
#pragma unroll 1
            for (int jg = 0; jg < kBJ; jg += kCG) {
                float c[kCG];
#pragma unroll
                for (int jj = 0; jj < kCG; ++jj) {
                    const float4* bp = reinterpret_cast<const float4*>(sx + (jg + jj) * DPF);
                    float b[DPF];
#pragma unroll
                    for (int h = 0; h < DPF / 4; ++h) {
                        const float4 p = bp[h];
                        b[4 * h] = p.x; b[4 * h + 1] = p.y; b[4 * h + 2] = p.z; b[4 * h + 3] = p.w;
                    }
                    const float vj = sv[jg + jj];
                    float cs = 0.0f;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) {
                        float q = na[ti] + b[DPF - 1];
#pragma unroll
                        for (int k = 0; k < D; ++k) q = fmaf(a2[ti][k], b[k], q);
                        const float kk = kappa_f32<KIND>(q);
                        rt[ti] = fmaf(kk, vj, rt[ti]);
                        cs = fmaf(kk, vi[ti], cs);
                    }
                    c[jj] = cs;
                }
                if (offdiag) {
                    col_reduce_f32<kCG>(c, lane);
                    if ((lane & (32 / kCG - 1)) == 0) scol[jg + reduced_col<kCG>(lane)] = c[0];
                }
            }
            // FP32 row partials of this tile (64 columns) -> FP64 accumulators
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) racc[ti] += (double)rt[ti];
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }

            if (offdiag) {
                __syncthreads();
                if (tid < kBJ) {
                    const long j = c0 + (long)tile * kBJ + tid;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) s += (double)s_col[(colbuf * WARPS + w) * kBJ + tid];
                    if (j < args.n) atomicAdd(args.y + (long)blockIdx.x * args.ystride + j, var * s);
                }
                colbuf ^= 1;
            }
        }
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
            if (live[ti]) atomicAdd(args.y + (long)blockIdx.x * args.ystride + r0 + ti * kThreads + tid, var * racc[ti]);
        __syncthreads();
        cc.tau += gridDim.x;
        load_item(cc);
    }
}

// ---------------------------------------------------------------------------------------------
// K2 in fp32-pair mode: the fused backward sweep of kmv_bwd_kernel (kmv_impl.cuh) with the pair
// arithmetic in FP32.  Thread-private FP32 partials live for one 64-column tile only and are flushed into FP64
// accumulators (row sums R, the d cross terms -2 X_q, the variance sum) at every tile.
// ---------------------------------------------------------------------------------------------
struct BwdArgsF32 {
    const float* xp;                       // packed fp32 [n_pad][DPF]
    const float* wcol; const float* ucol;  // padded float copies of w and u
    double* rsum;                          // R: CTA b accumulates into rsum + b * ystride
    double* gout;                          // [D+1]: -2 X_q ..., variance sum of CTA b at gout + b * gstride
    long ystride, gstride;
    long n, nb, nitems;
    int part, nparts;
};

template <int KIND>
__device__ __forceinline__ void kappa_dweight_f32(float q, float& kap, float& ew) {
    q = fmaxf(q, 0.0f);
    if (KIND == CGLB_MATERN32) {
        float s;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(q));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ew) : "f"(-s));
        kap = fmaf(s * ew, 0.69314718055994531f, ew);
    } else {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ew) : "f"(-q));
        kap = ew;
    }
}

template <int KIND, int D, int TI, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) f32_bwd_kernel(const BwdArgsF32 args) {
    constexpr int DPF = (D + 4) & ~3;
    constexpr int kThreads = WARPS * 32;
    constexpr int BI = kThreads * TI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);                    // [kStages][kBJ*DPF]
    float* s_v = s_x + kStages * kBJ * DPF;                             // [kStages][2][kBJ]  (w, u)
    float* s_col = s_v + kStages * 2 * kBJ;                             // [2][WARPS][kBJ]
    double* s_red = reinterpret_cast<double*>(s_col + 2 * WARPS * kBJ); // [WARPS][D+2]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_red + WARPS * (D + 2));
    uint64_t* s_empty = s_full + kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    struct Cur {
        long tau, I, C, c0; int tile, ntiles; bool valid;
    };
    auto load_item = [&](Cur& c) {
        const long t = c.tau * args.nparts + args.part;
        c.valid = t < args.nitems;
        if (!c.valid) return;
        item_to_blocks_sym(t, c.I, c.C);
        c.c0 = c.C * BI;
        long cend = c.c0 + BI;
        if (cend > args.n) cend = args.n;
        c.ntiles = (int)((cend - c.c0 + kBJ - 1) / kBJ);
        c.tile = 0;
    };
    int pstage = 0, stage = 0;
    uint32_t pphase = 0, phase = 0;
    auto produce = [&](Cur& pc) {
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DPF + 2 * kBJ) * sizeof(float)));
        tma_load_1d(s_x + pstage * kBJ * DPF, args.xp + j0 * DPF, kBJ * DPF * sizeof(float), &s_full[pstage]);
        tma_load_1d(s_v + pstage * 2 * kBJ, args.wcol + j0, kBJ * sizeof(float), &s_full[pstage]);
        tma_load_1d(s_v + pstage * 2 * kBJ + kBJ, args.ucol + j0, kBJ * sizeof(float), &s_full[pstage]);
        if (++pstage == kStages) { pstage = 0; pphase ^= 1; }
        if (++pc.tile == pc.ntiles) { pc.tau += gridDim.x; load_item(pc); }
    };

    Cur cc;
    cc.tau = blockIdx.x;
    load_item(cc);
    Cur pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kPrefetch; ++i) produce(pc);
    }
    int colbuf = 0;
    double gq[D], gvar = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) gq[k] = 0.0;
    const double inv_scale_sq = 1.0 / (f32_input_scale(KIND) * f32_input_scale(KIND));

    while (cc.valid) {
        const bool offdiag = cc.I != cc.C;
        const float half = offdiag ? 1.0f : 0.5f;
        const long r0 = cc.I * BI;
        float a2[TI][D], na[TI], ui[TI], wi[TI];
        double racc[TI];
        bool live[TI];
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            const long row = r0 + ti * kThreads + tid;
            live[ti] = row < args.n;
            const float4* src = reinterpret_cast<const float4*>(args.xp + (live[ti] ? row : 0) * DPF);
            float tmp[DPF];
#pragma unroll
            for (int h = 0; h < DPF / 4; ++h) {
                const float4 p = __ldg(src + h);
                tmp[4 * h] = p.x; tmp[4 * h + 1] = p.y; tmp[4 * h + 2] = p.z; tmp[4 * h + 3] = p.w;
            }
#pragma unroll
            for (int k = 0; k < D; ++k) a2[ti][k] = live[ti] ? -2.0f * tmp[k] : 0.0f;
            na[ti] = live[ti] ? tmp[DPF - 1] : 0.0f;
            ui[ti] = live[ti] ? half * __ldg(args.ucol + row) : 0.0f;
            wi[ti] = live[ti] ? half * __ldg(args.wcol + row) : 0.0f;
            racc[ti] = 0.0;
        }
        const int ntiles = cc.ntiles;
        const long c0 = cc.c0;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) produce(pc);
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const float* sx = s_x + stage * kBJ * DPF;
            const float* sw = s_v + stage * 2 * kBJ;
            const float* su = sw + kBJ;
            float* scol = s_col + (colbuf * WARPS + warp) * kBJ;
            float rt[TI], gqt[D], gvt = 0.0f;
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) rt[ti] = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) gqt[k] = 0.0f;
#pragma unroll 1
            for (int jg = 0; jg < kBJ; jg += kCG) {
                float c[kCG];
#pragma unroll
                for (int jj = 0; jj < kCG; ++jj) {
                    const float4* bp = reinterpret_cast<const float4*>(sx + (jg + jj) * DPF);
                    float b[DPF];
#pragma unroll
                    for (int h = 0; h < DPF / 4; ++h) {
                        const float4 p = bp[h];
                        b[4 * h] = p.x; b[4 * h + 1] = p.y; b[4 * h + 2] = p.z; b[4 * h + 3] = p.w;
                    }
                    const float wj = sw[jg + jj], uj = su[jg + jj];
                    float cs = 0.0f, wq[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) wq[k] = 0.0f;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) {
                        float q = na[ti] + b[DPF - 1];
#pragma unroll
                        for (int k = 0; k < D; ++k) q = fmaf(a2[ti][k], b[k], q);
                        float kap, ew;
                        kappa_dweight_f32<KIND>(q, kap, ew);
                        const float om = fmaf(wi[ti], uj, ui[ti] * wj);
                        const float cw = ew * om;
                        gvt = fmaf(kap, om, gvt);
                        rt[ti] += cw;
                        cs += cw;
#pragma unroll
                        for (int k = 0; k < D; ++k) wq[k] = fmaf(cw, a2[ti][k], wq[k]);
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) gqt[k] = fmaf(b[k], wq[k], gqt[k]);
                    c[jj] = cs;
                }
                if (offdiag) {
                    col_reduce_f32<kCG>(c, lane);
                    if ((lane & (32 / kCG - 1)) == 0) scol[jg + reduced_col<kCG>(lane)] = c[0];
                }
            }
            // FP32 partials of this tile -> FP64 accumulators
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) racc[ti] += (double)rt[ti];
#pragma unroll
            for (int k = 0; k < D; ++k) gq[k] = fma((double)gqt[k], inv_scale_sq, gq[k]);    // a_i a_j carries scale^2
            gvar += (double)gvt;
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }

            if (offdiag) {
                __syncthreads();
                if (tid < kBJ) {
                    const long j = c0 + (long)tile * kBJ + tid;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) s += (double)s_col[(colbuf * WARPS + w) * kBJ + tid];
                    if (j < args.n) atomicAdd(args.rsum + (long)blockIdx.x * args.ystride + j, s);
                }
                colbuf ^= 1;
            }
        }
        const double rscale = offdiag ? 1.0 : 2.0;
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
            if (live[ti]) atomicAdd(args.rsum + (long)blockIdx.x * args.ystride + r0 + ti * kThreads + tid, rscale * racc[ti]);
        __syncthreads();
        cc.tau += gridDim.x;
        load_item(cc);
    }

    // block reduction of the thread-private accumulators -> one atomic per CTA per component
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double s = warp_sum(gq[k]);
        if (lane == 0) s_red[warp * (D + 2) + k] = s;
    }
    {
        const double s = warp_sum(gvar);
        if (lane == 0) s_red[warp * (D + 2) + D] = s;
    }
    __syncthreads();
    if (tid <= D) {
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) s += s_red[w * (D + 2) + tid];
        args.gout[(long)blockIdx.x * args.gstride + tid] = s;
    }
}

template <int KIND, int D, int TI, int WARPS>
static int launch_f32_bwd(Context* ctx, BwdArgsF32 a, cudaStream_t st) {
    constexpr int DPF = (D + 4) & ~3;
    constexpr long BI = WARPS * 32 * TI;
    a.nb = (a.n + BI - 1) / BI;
    a.nitems = a.nb * (a.nb + 1) / 2;
    auto kern = f32_bwd_kernel<KIND, D, TI, WARPS>;
    const size_t smem = (size_t)(kStages * kBJ * DPF + kStages * 2 * kBJ + 2 * WARPS * kBJ) * sizeof(float) +
                        (size_t)WARPS * (D + 2) * sizeof(double) + 2 * kStages * sizeof(uint64_t);
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    const int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

template <int KIND, int D>
static int run_f32_bwd(Context* ctx, const BwdArgsF32& a, cudaStream_t st) {
    if constexpr (D <= 16) {
        if (count_items(a.n, a.n, true, a.nparts, 8 * 32 * 4) >= 12L * ctx->num_sms) return launch_f32_bwd<KIND, D, 4, 8>(ctx, a, st);
    }
    if (count_items(a.n, a.n, true, a.nparts, 8 * 32 * 2) >= 12L * ctx->num_sms) return launch_f32_bwd<KIND, D, 2, 8>(ctx, a, st);
    return launch_f32_bwd<KIND, D, 1, 8>(ctx, a, st);
}

template <int KIND, int D, int TI, int WARPS>
static int launch_f32(Context* ctx, SweepArgsF32 a, cudaStream_t st) {
    constexpr int DPF = (D + 4) & ~3;
    constexpr long BI = WARPS * 32 * TI;
    a.nb = (a.n + BI - 1) / BI;
    a.nitems = a.nb * (a.nb + 1) / 2;
    auto kern = f32_sweep_kernel<KIND, D, TI, WARPS>;
    const size_t smem = (size_t)(kStages * kBJ * DPF + kStages * kBJ + 2 * WARPS * kBJ) * sizeof(float) + 2 * kStages * sizeof(uint64_t);
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    const int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

// rows per CTA: big blocks amortise the column loads best; small problems need more items than SMs
template <int KIND, int D>
static int run_f32(Context* ctx, const SweepArgsF32& a, cudaStream_t st) {
#ifdef CGLB_KMV_EXPERIMENT
    if (const char* e = getenv("CGLB_F32_VARIANT")) {
        switch (atoi(e)) {
            case 88: return launch_f32<KIND, D, 8, 8>(ctx, a, st);
            case 48: return launch_f32<KIND, D, 4, 8>(ctx, a, st);
            case 412: return launch_f32<KIND, D, 4, 12>(ctx, a, st);
            case 416: return launch_f32<KIND, D, 4, 16>(ctx, a, st);
            case 216: return launch_f32<KIND, D, 2, 16>(ctx, a, st);
            case 612: return launch_f32<KIND, D, 6, 10>(ctx, a, st);
            default: break;
        }
    }
#endif
    // measured on B200 (tools/dev_f32_variants.py): 8 warps x 8 rows and 16 warps x 4 rows tie at d = 11 (2.24 Tpairs/s),
    // 16 x 4 wins at d = 3 (3.64 vs 3.16)
    if (count_items(a.n, a.n, true, a.nparts, 2048) >= 12L * ctx->num_sms) {
        if constexpr (D <= 4) return launch_f32<KIND, D, 4, 16>(ctx, a, st);
        else if constexpr (D <= 16) return launch_f32<KIND, D, 8, 8>(ctx, a, st);
    }
    if (count_items(a.n, a.n, true, a.nparts, 8 * 32 * 4) >= 12L * ctx->num_sms) return launch_f32<KIND, D, 4, 8>(ctx, a, st);
    return launch_f32<KIND, D, 1, 8>(ctx, a, st);
}

typedef int (*f32_fn)(Context*, int kind, const SweepArgsF32&, cudaStream_t);
typedef int (*f32_bwd_fn)(Context*, int kind, const BwdArgsF32&, cudaStream_t);
f32_fn get_f32_fn(int d);
f32_bwd_fn get_f32_bwd_fn(int d);

}  // namespace cglb
