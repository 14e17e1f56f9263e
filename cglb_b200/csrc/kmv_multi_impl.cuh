// K1 against a block of right-hand sides:  Y = variance * K(X,X) V + diag * V,  V = [n][T] row-major, T = 2 or 4.
//
// Replaces `A @ x` for x of shape [n, t] (reference conjugate_gradient.py:57,66,72 accept [N, t]; SURVEY.md 8b "b3").
// Same structure as the register-resident symmetric sweep of kmv_impl.cuh -- persistent CTAs, (row block, column chunk)
// items on or above the diagonal, TMA ring of 64-column tiles, expanded-form distances, one kernel-pair evaluation feeding
// y_i AND y_j -- but every evaluated pair is used for T accumulations on each side: the kernel map (sqrt, exp: 13 of the
// 27 FP64 slots at d = 11) and the distance are paid once per pair, not once per right-hand side.
// Algorithmic FLOPs per pair: 3d + 5 + 2T (Matern32), 3d + 2 + 2T (RBF).
// The per-CTA copies / fixed summation order are those of kmv_impl.cuh (row pitch T).
#pragma once
#include "kmv_impl.cuh"

namespace cglb {

template <int KIND, int D, int TI, int WARPS, int CB, int T>
__global__ void __launch_bounds__(WARPS * 32, 1) kmv_multi_kernel(const SweepArgs args) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int kWarps = WARPS, kThreads = WARPS * 32;
    constexpr int BI = kThreads * TI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_x = reinterpret_cast<double*>(smem_raw);                    // [kStages][kBJ*DP]
    double* s_v = s_x + kStages * kBJ * DP;                               // [kStages][kBJ*T]
    double* s_col = s_v + kStages * kBJ * T;                              // [2][kWarps][T][kBJ]
    double* s_tab = s_col + 2 * kWarps * T * kBJ;                         // [1024]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_tab + kExpTabBig);
    uint64_t* s_empty = s_full + kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kExpTabBig; i += kThreads) s_tab[i] = args.exp_tab[kExpTabSmall + i];
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    int pstage = 0, stage = 0;
    uint32_t pphase = 0, phase = 0;
    Cursor<BI, true> cc, pc;
    cc.start(args);
    pc = cc;
    auto produce = [&]() {
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DP + kBJ * T) * sizeof(double)));
        tma_load_1d(s_x + pstage * kBJ * DP, args.xp_cols + j0 * DP, kBJ * DP * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_v + pstage * kBJ * T, args.vcol + j0 * T, kBJ * T * sizeof(double), &s_full[pstage]);
        if (++pstage == kStages) { pstage = 0; pphase ^= 1; }
        pc.next_tile(args);
    };
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kPrefetch; ++i) produce();
    }
    int colbuf = 0;
    const double var = args.variance;
    double* const yb = args.y + (long)blockIdx.x * args.ystride;

    while (cc.valid) {
        const bool offdiag = (cc.I != cc.C);
        const long r0 = cc.I * BI;
        double a2[TI][D], na[TI], vi[TI][T], racc[TI][T];
        bool live[TI];
        load_rows<D, DP, TI, kThreads>(args.xp_rows, r0, args.nrows, tid, a2, na, live);
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
#pragma unroll
            for (int r = 0; r < T; ++r) {
                vi[ti][r] = live[ti] ? __ldg(args.vcol + (r0 + ti * kThreads + tid) * T + r) : 0.0;
                racc[ti][r] = 0.0;
            }
        const int ntiles = cc.ntiles;
        const long c0 = cc.c0;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) produce();
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const double* sx = s_x + stage * kBJ * DP;
            const double* sv = s_v + stage * kBJ * T;
            double* scol = s_col + (colbuf * kWarps + warp) * T * kBJ;
#pragma unroll 1
            for (int jg = 0; jg < kBJ; jg += kCG) {
                double c[T][kCG];
#pragma unroll
                for (int jb = 0; jb < kCG; jb += CB) {
                    double q[CB][TI];
                    {
                        double b[CB][DP];
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) {
                            const double2* bp = reinterpret_cast<const double2*>(sx + (jg + jb + cb) * DP);
#pragma unroll
                            for (int h = 0; h < DP / 2; ++h) {
                                double2 p = bp[h];
                                b[cb][2 * h] = p.x;
                                b[cb][2 * h + 1] = p.y;
                            }
#pragma unroll
                            for (int ti = 0; ti < TI; ++ti) q[cb][ti] = na[ti] + b[cb][DP - 1];
                        }
#pragma unroll
                        for (int k = 0; k < D; ++k)
#pragma unroll
                            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                                for (int ti = 0; ti < TI; ++ti) q[cb][ti] = fma(a2[ti][k], b[cb][k], q[cb][ti]);
                    }
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                        for (int ti = 0; ti < TI; ++ti) q[cb][ti] = kappa<KIND, 10>(q[cb][ti], s_tab);
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        double vj[T];
                        if (T == 2) {
                            const double2 p = *reinterpret_cast<const double2*>(sv + (jg + jb + cb) * 2);
                            vj[0] = p.x; vj[1] = p.y;
                        } else {
#pragma unroll
                            for (int h = 0; h < T / 2; ++h) {
                                const double2 p = *reinterpret_cast<const double2*>(sv + (jg + jb + cb) * T + 2 * h);
                                vj[2 * h] = p.x; vj[2 * h + 1] = p.y;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < T; ++r) {
                            double cs = 0.0;
#pragma unroll
                            for (int ti = 0; ti < TI; ++ti) {
                                racc[ti][r] = fma(q[cb][ti], vj[r], racc[ti][r]);
                                cs = fma(q[cb][ti], vi[ti][r], cs);
                            }
                            c[r][jb + cb] = cs;
                        }
                    }
                }
                if (offdiag) {
#pragma unroll
                    for (int r = 0; r < T; ++r) {
                        col_reduce<kCG>(c[r], lane);
                        if ((lane & (32 / kCG - 1)) == 0) scol[r * kBJ + jg + reduced_col<kCG>(lane)] = c[r][0];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }

            if (offdiag) {
                __syncthreads();
                for (int idx = tid; idx < kBJ * T; idx += kThreads) {
                    const int col = idx / T, r = idx % T;          // consecutive threads -> consecutive addresses of Y
                    const long j = c0 + (long)tile * kBJ + col;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) s += s_col[((colbuf * kWarps + w) * T + r) * kBJ + col];
                    if (j < args.ncols) atomicAdd(yb + j * T + r, var * s);
                }
                colbuf ^= 1;
            }
        }
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
            if (live[ti]) {
#pragma unroll
                for (int r = 0; r < T; ++r) atomicAdd(yb + (r0 + ti * kThreads + tid) * T + r, var * racc[ti][r]);
            }
        __syncthreads();
        cc.next_item(args);
    }
}

template <int KIND, int D, int T>
static int run_multi(Context* ctx, SweepArgs a, cudaStream_t st) {
    // rows per thread: registers hold TI x D coordinates + 2 TI T accumulators + T x 8 column partials
    constexpr int TI = (D <= 12 && T == 2) ? 2 : 1;
    constexpr int CB = (TI == 2) ? 2 : 4;
    constexpr int WARPS = 8;
    constexpr int DP = SmemLayout<D>::DP;
    constexpr long BI = WARPS * 32 * TI;
    a.nb_rows = (a.nrows + BI - 1) / BI;
    a.nb_cols = a.nb_rows;
    a.nitems = a.nb_rows * (a.nb_rows + 1) / 2;
    auto kern = kmv_multi_kernel<KIND, D, TI, WARPS, CB, T>;
    const size_t smem = (size_t)(kStages * kBJ * DP + kStages * kBJ * T + 2 * WARPS * T * kBJ + kExpTabBig) * sizeof(double) +
                        2 * kStages * sizeof(uint64_t);
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    const int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

}  // namespace cglb
