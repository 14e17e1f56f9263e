// K1 / K2: matrix-free kernel-evaluate-and-matvec sweeps for sm_100a.
//
//   y = variance * K(X,X) v + diag * v          (cglb_kmv_sym, cglb_kmv_rect)
//   d(u^T K w)/d{lengthscale, variance}          (cglb_kmv_bwd_sym)
//
// This file holds the register-resident DFMA sweeps (every d <= 32; the K*v sweeps of d = 10, 11, 13..32 and the
// backward sweeps of d = 6..8, 10..32 take the DMMA kernels of dsweep_impl.cuh when the problem is large enough,
// fp32 models the FP32-pair kernels of f32sweep_impl.cuh).
// Replaces the KeOps Genred reductions behind `A @ p` (reference conjugate_gradient.py:57,66,72,
// models.py:280) and their autograd (optimizer.py:97).  Design:
//   * persistent CTAs (one per SM), 8 warps; lane 0 of warp 0 also drives the TMA ring two tiles ahead;
//   * a work item is a (row block, column chunk) square of BI x BI kernel pairs; the symmetric sweep
//     visits only chunks on or above the diagonal and uses every evaluated k_ij for y_i AND y_j;
//   * each consumer thread keeps TI rows (scaled coordinates, -2a and |a|^2) in registers; column tiles
//     of BJ packed rows + the matching slice of v arrive in shared memory through cp.async.bulk (TMA)
//     on a 4-stage mbarrier ring and are read as warp-wide broadcasts;
//   * squared distances in the expanded form |a|^2+|b|^2-2ab (d+1 DFMA-class slots instead of 2d),
//     clamped in the integer pipe, sqrt/exp from common.cuh (5+7 slots);
//   * column sums: transposing butterfly over groups of CG columns (9 DADD per 8 columns), then a
//     cross-warp shared-memory reduction and one RED.ADD.F64 per column per tile.
#pragma once
#include "common.cuh"

namespace cglb {

// every warp computes; lane 0 of warp 0 also drives the TMA ring.  WARPS is a template parameter: 8 warps x 4
// rows/thread, 12 x 2 or 16 x 2 trade registers (ILP) against resident warps (TLP); see DESIGN.md section 3.
constexpr int kBJ = 64;                           // columns per pipeline stage
constexpr int kStages = 4;
constexpr int kPrefetch = 2;                      // tiles in flight ahead of the one being consumed
constexpr int kCG = 8;                            // column group of the transposing reduction

struct SweepArgs {
    const double* xp_rows;   // packed rows   [rows_pad][DP]
    const double* xp_cols;   // packed cols   [cols_pad][DP]
    const double* vcol;      // padded column vector (v for fwd; w for bwd)   [cols_pad]
    const double* ucol;      // bwd only: u on the column side                [cols_pad]
    double* y;               // fwd: output; bwd: R row sums.  CTA b accumulates into y + b * ystride (its own copy when
                             // ystride > 0: deterministic, summed in CTA order afterwards; ystride = 0: one shared vector)
    double* gout;            // bwd: [D+1] sums (-2 X_q ..., variance sum) of CTA b at gout + b * gstride (plain stores)
    long ystride, gstride;
    long nrows, ncols;       // valid counts
    long nb_rows, nb_cols;   // number of BI blocks
    long nitems;             // total work items (global, before the part split)
    long sr_chunks;          // DMMA sweeps: column chunks per super-row of the item order (dsweep_impl.cuh)
    const double* exp_tab;   // 64 doubles 2^(j/64) in global memory
    double variance;
    int part, nparts;
};

template <int CLAMP_HI>
__device__ __forceinline__ double clamp_sq_t(double q) {
    int hi = __double2hiint(q);
    hi = max(hi, 0x01700000);   // 2^-1000 (negative q -> hi < 0 as a signed int)
    hi = min(hi, CLAMP_HI);
    return __hiloint2double(hi, __double2loint(q));
}

// kappa(q): Matern32 -> (1+s) e^-s with s = sqrt(q) (inputs pre-scaled by sqrt3/l);  RBF -> e^-q
// TB selects the exp table (6: 64 entries / 9 slots, 10: 1024 entries / 7 slots).  The squared distance is clamped
// to [2^-1000, 693^2] resp. [2^-1000, 693] with two integer min/max on its high word (no FP64 slot): the lower
// bound absorbs the tiny negative values of the expanded form and the diagonal, the upper bound keeps
// e^-s >= 2^-1000 (so fast_exp_neg needs no exponent clamp; kernel values below 1e-301 are flushed to ~1e-301).
constexpr int kClampHiMatern = 0x411d4fe4;   // high word of 693^2
constexpr int kClampHiRbf = 0x4085a800;      // high word of 693
template <int KIND, int TB = 6>
__device__ __forceinline__ double kappa(double q, const double* tab) {
    if (KIND == CGLB_MATERN32) {
        q = clamp_sq_t<kClampHiMatern>(q);
        double s = fast_sqrt(q);
        double e = fast_exp_neg<TB>(s, tab);
        return fma(s, e, e);
    } else {
        q = clamp_sq_t<kClampHiRbf>(q);
        return fast_exp_neg<TB>(q, tab);
    }
}

// kappa and the lengthscale-derivative weight e' (Matern32: e^-s; RBF: e^-q, factor 2 applied by host)
template <int KIND, int TB = 6>
__device__ __forceinline__ void kappa_and_dweight(double q, const double* tab, double& kap, double& ew) {
    if (KIND == CGLB_MATERN32) {
        q = clamp_sq_t<kClampHiMatern>(q);
        double s = fast_sqrt(q);
        ew = fast_exp_neg<TB>(s, tab);
        kap = fma(s, ew, ew);
    } else {
        q = clamp_sq_t<kClampHiRbf>(q);
        ew = fast_exp_neg<TB>(q, tab);
        kap = ew;
    }
}

// Transposing butterfly: every lane holds CG partial column sums; on return lanes with
// (lane & (32/CG - 1)) == 0 hold in c[0] the warp-wide sum of column `reduced_col(lane)`.
template <int CG>
__device__ __forceinline__ void col_reduce(double (&c)[CG], int lane) {
    int cnt = CG;
    int off = 16;
#pragma unroll
    for (; cnt > 1; cnt >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int h = 0; h < cnt / 2; ++h) {
            double send = up ? c[h] : c[h + cnt / 2];
            double keep = up ? c[h + cnt / 2] : c[h];
            c[h] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) c[0] += __shfl_xor_sync(0xffffffffu, c[0], off);
}
template <int CG>
__device__ __forceinline__ int reduced_col(int lane) {
    int col = 0, cnt = CG, off = 16;
    for (; cnt > 1; cnt >>= 1, off >>= 1) col = col * 2 + ((lane & off) ? 1 : 0);
    return col;
}

__device__ __forceinline__ void item_to_blocks_sym(long t, long& I, long& C) {
    // t = C(C+1)/2 + I, I <= C
    long c = (long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (c * (c + 1) / 2 > t) --c;
    while ((c + 1) * (c + 2) / 2 <= t) ++c;
    C = c;
    I = t - c * (c + 1) / 2;
}

template <int D>
struct SmemLayout {
    static constexpr int DP = (D + 2) & ~1;
};

// Position in the static (round-robin over CTAs) stream of work items and of their column tiles.
// The TMA-issuing lane and the consumers each walk their own cursor over the same sequence.
template <int BI, bool SYM>
struct Cursor {
    long tau;       // local item counter: global item t = tau * nparts + part
    long I, C;      // row block, column chunk
    long c0;        // first column of the chunk
    int tile, ntiles;
    bool valid;
    __device__ __forceinline__ void load_item(const SweepArgs& a) {
        const long t = tau * a.nparts + a.part;
        valid = t < a.nitems;
        if (!valid) return;
        if (SYM) item_to_blocks_sym(t, I, C);
        else { I = t / a.nb_cols; C = t % a.nb_cols; }
        c0 = C * BI;
        long cend = c0 + BI;
        if (cend > a.ncols) cend = a.ncols;
        ntiles = (int)((cend - c0 + kBJ - 1) / kBJ);
        tile = 0;
    }
    __device__ __forceinline__ void start(const SweepArgs& a) { tau = blockIdx.x; load_item(a); }
    __device__ __forceinline__ void next_item(const SweepArgs& a) { tau += gridDim.x; load_item(a); }
    __device__ __forceinline__ void next_tile(const SweepArgs& a) { if (++tile == ntiles) next_item(a); }
};

// mbarrier ring shared by the forward and backward sweeps.  NVEC = vectors staged per tile (1: v; 2: w,u)
template <int D, int NVEC, int WARPS>
struct TileRing {
    static constexpr int DP = SmemLayout<D>::DP;
    double* s_x;          // [kStages][kBJ*DP]
    double* s_v;          // [kStages][NVEC][kBJ]
    uint64_t* s_full;     // [kStages]
    uint64_t* s_empty;    // [kStages]
    int pstage; uint32_t pphase;     // producer side
    int stage;  uint32_t phase;      // consumer side

    __device__ __forceinline__ void init_barriers() {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], WARPS);
        }
        mbar_fence_init();
    }
    template <class CursorT>
    __device__ __forceinline__ void produce(CursorT& pc, const SweepArgs& a) {   // one lane only
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DP + NVEC * kBJ) * sizeof(double)));
        tma_load_1d(s_x + pstage * kBJ * DP, a.xp_cols + j0 * DP, kBJ * DP * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_v + pstage * NVEC * kBJ, a.vcol + j0, kBJ * sizeof(double), &s_full[pstage]);
        if (NVEC == 2) tma_load_1d(s_v + pstage * NVEC * kBJ + kBJ, a.ucol + j0, kBJ * sizeof(double), &s_full[pstage]);
        if (++pstage == kStages) { pstage = 0; pphase ^= 1; }
        pc.next_tile(a);
    }
    __device__ __forceinline__ void consumer_wait() { mbar_wait(&s_full[stage], phase); }
    __device__ __forceinline__ void consumer_release(int lane) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
};

template <int D, int DP, int TI, int kThreads>
__device__ __forceinline__ void load_rows(const double* __restrict__ xp, long r0, long nrows, int tid,
                                          double (&a2)[TI][D], double (&na)[TI], bool (&live)[TI]) {
#pragma unroll
    for (int ti = 0; ti < TI; ++ti) {
        const long row = r0 + ti * kThreads + tid;
        live[ti] = row < nrows;
        const double2* src = reinterpret_cast<const double2*>(xp + (live[ti] ? row : 0) * DP);
        double tmp[DP];
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            double2 p = __ldg(src + h);
            tmp[2 * h] = live[ti] ? p.x : 0.0;
            tmp[2 * h + 1] = live[ti] ? p.y : 0.0;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) a2[ti][k] = -2.0 * tmp[k];
        na[ti] = tmp[DP - 1];
    }
}

// ---------------------------------------------------------------------------------------------
// forward sweep
// ---------------------------------------------------------------------------------------------
// CB = columns whose pair evaluations are interleaved: CB*TI independent dependency chains per warp (the
// FP64 pipe needs ~12+ chains per scheduler to hide the 8-cycle DFMA latency through the sqrt/exp chains)
template <int KIND, int D, int TI, bool SYM, int WARPS, int CB>
__global__ void __launch_bounds__(WARPS * 32, 1) kmv_sweep_kernel(const SweepArgs args) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int kWarps = WARPS, kThreads = WARPS * 32;
    constexpr int BI = kThreads * TI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileRing<D, 1, WARPS> ring;
    ring.s_x = reinterpret_cast<double*>(smem_raw);                       // [kStages][kBJ*DP]
    ring.s_v = ring.s_x + kStages * kBJ * DP;                             // [kStages][kBJ]
    double* s_col = ring.s_v + kStages * kBJ;                             // [2][kWarps][kBJ]
    double* s_tab = s_col + 2 * kWarps * kBJ;                             // [1024] 2^(j/1024)
    ring.s_full = reinterpret_cast<uint64_t*>(s_tab + kExpTabBig);        // [kStages]
    ring.s_empty = ring.s_full + kStages;
    ring.pstage = ring.stage = 0;
    ring.pphase = ring.phase = 0;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kExpTabBig; i += kThreads) s_tab[i] = args.exp_tab[kExpTabSmall + i];
    if (tid == 0) ring.init_barriers();
    __syncthreads();

    Cursor<BI, SYM> cc, pc;     // consumer / producer cursors
    cc.start(args);
    pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kPrefetch; ++i) ring.produce(pc, args);
    }
    int colbuf = 0;
    const double var = args.variance;
    // Fixed summation order: all adds of this CTA go to its own copy of y.  Two adds to the same address are either issued
    // by the same thread (column j -> thread j % 64, row i -> thread i % kThreads) or separated by a CTA barrier (the
    // per-tile barrier below, the barrier at the end of every item), and the items of a CTA are a static sequence.
    double* const yb = args.y + (long)blockIdx.x * args.ystride;

    while (cc.valid) {
        const bool offdiag = SYM && (cc.I != cc.C);
        const long r0 = cc.I * BI;
        double a2[TI][D], na[TI], vi[TI], racc[TI];
        bool live[TI];
        load_rows<D, DP, TI, kThreads>(args.xp_rows, r0, args.nrows, tid, a2, na, live);
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            // SYM: rows and columns share the padded vector
            vi[ti] = (SYM && live[ti]) ? __ldg(args.vcol + r0 + ti * kThreads + tid) : 0.0;
            racc[ti] = 0.0;
        }
        const int ntiles = cc.ntiles;
        const long c0 = cc.c0;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) ring.produce(pc, args);
            __syncwarp();
            ring.consumer_wait();
            const double* sx = ring.s_x + ring.stage * kBJ * DP;
            const double* sv = ring.s_v + ring.stage * kBJ;
            double* scol = s_col + (colbuf * kWarps + warp) * kBJ;
#pragma unroll 1
            for (int jg = 0; jg < kBJ; jg += kCG) {
                double c[kCG];
#pragma unroll
                for (int jb = 0; jb < kCG; jb += CB) {
                    double q[CB][TI], vj[CB];
                    // phase 1: CB*TI squared distances, dimension-major so that the chains interleave
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
                            vj[cb] = sv[jg + jb + cb];
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
                    // phase 2: kernel map on all chains
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                        for (int ti = 0; ti < TI; ++ti) q[cb][ti] = kappa<KIND, 10>(q[cb][ti], s_tab);
                    // phase 3: row / column accumulation
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        double cs = 0.0;
#pragma unroll
                        for (int ti = 0; ti < TI; ++ti) {
                            racc[ti] = fma(q[cb][ti], vj[cb], racc[ti]);
                            if (SYM) cs = fma(q[cb][ti], vi[ti], cs);
                        }
                        c[jb + cb] = cs;
                    }
                }
                if (offdiag) {
                    col_reduce<kCG>(c, lane);
                    if ((lane & (32 / kCG - 1)) == 0) scol[jg + reduced_col<kCG>(lane)] = c[0];
                }
            }
            ring.consumer_release(lane);

            if (offdiag) {
                // cross-warp reduction of the column sums of this tile, one RED per column
                __syncthreads();
                if (tid < kBJ) {
                    const long j = c0 + (long)tile * kBJ + tid;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) s += s_col[(colbuf * kWarps + w) * kBJ + tid];
                    if (j < args.ncols) atomicAdd(yb + j, var * s);
                }
                colbuf ^= 1;
            }
        }
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
            if (live[ti]) atomicAdd(yb + r0 + ti * kThreads + tid, var * racc[ti]);
        if (SYM) __syncthreads();       // row adds of this item before the column adds of the next (fixed order)
        cc.next_item(args);
    }
}

// ---------------------------------------------------------------------------------------------
// backward sweep (symmetric only):  see DESIGN.md section 3.3
//   per unordered pair: omega = u_i w_j + w_i u_j, c = e' * omega
//   R_i += c, R_j += c                      (args.y = R, atomics)
//   gq[k] += b_jk * sum_ti c * (-2 a_ik)    (= -2 X_k, thread-private, reduced at kernel end)
//   gvar  += kappa * omega
// diagonal items visit ordered pairs with u_i,w_i halved and double their row sums at the end.
// ---------------------------------------------------------------------------------------------
template <int KIND, int D, int TI, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kmv_bwd_kernel(const SweepArgs args) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int kWarps = WARPS, kThreads = WARPS * 32;
    constexpr int BI = kThreads * TI;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileRing<D, 2, WARPS> ring;
    ring.s_x = reinterpret_cast<double*>(smem_raw);                       // [kStages][kBJ*DP]
    ring.s_v = ring.s_x + kStages * kBJ * DP;                             // [kStages][2][kBJ]  (w, u)
    double* s_col = ring.s_v + kStages * 2 * kBJ;                         // [2][kWarps][kBJ]
    double* s_tab = s_col + 2 * kWarps * kBJ;                             // [1024] 2^(j/1024)
    double* s_red = s_tab + kExpTabBig;                                   // [kWarps][D+2]
    ring.s_full = reinterpret_cast<uint64_t*>(s_red + kWarps * (D + 2));
    ring.s_empty = ring.s_full + kStages;
    ring.pstage = ring.stage = 0;
    ring.pphase = ring.phase = 0;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kExpTabBig; i += kThreads) s_tab[i] = args.exp_tab[kExpTabSmall + i];
    if (tid == 0) ring.init_barriers();
    __syncthreads();

    Cursor<BI, true> cc, pc;
    cc.start(args);
    pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < kPrefetch; ++i) ring.produce(pc, args);
    }
    int colbuf = 0;
    double gq[D], gvar = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) gq[k] = 0.0;
    double* const yb = args.y + (long)blockIdx.x * args.ystride;      // this CTA's copy of R (fixed order, see the forward sweep)

    while (cc.valid) {
        const bool offdiag = (cc.I != cc.C);
        const double half = offdiag ? 1.0 : 0.5;
        const long r0 = cc.I * BI;
        double a2[TI][D], na[TI], ui[TI], wi[TI], racc[TI];
        bool live[TI];
        load_rows<D, DP, TI, kThreads>(args.xp_rows, r0, args.nrows, tid, a2, na, live);
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            const long row = r0 + ti * kThreads + tid;
            ui[ti] = live[ti] ? half * __ldg(args.ucol + row) : 0.0;
            wi[ti] = live[ti] ? half * __ldg(args.vcol + row) : 0.0;
            racc[ti] = 0.0;
        }
        const int ntiles = cc.ntiles;
        const long c0 = cc.c0;
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) ring.produce(pc, args);
            __syncwarp();
            ring.consumer_wait();
            const double* sx = ring.s_x + ring.stage * kBJ * DP;
            const double* sw = ring.s_v + ring.stage * 2 * kBJ;
            const double* su = sw + kBJ;
            double* scol = s_col + (colbuf * kWarps + warp) * kBJ;
#pragma unroll 1
            for (int jg = 0; jg < kBJ; jg += kCG) {
                double c[kCG];
#pragma unroll
                for (int jj = 0; jj < kCG; ++jj) {
                    const double2* bp = reinterpret_cast<const double2*>(sx + (jg + jj) * DP);
                    double b[DP];
#pragma unroll
                    for (int h = 0; h < DP / 2; ++h) {
                        double2 p = bp[h];
                        b[2 * h] = p.x;
                        b[2 * h + 1] = p.y;
                    }
                    const double nb = b[DP - 1];
                    const double wj = sw[jg + jj], uj = su[jg + jj];
                    double cs = 0.0;
                    double wq[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) wq[k] = 0.0;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) {
                        double q = na[ti] + nb;
#pragma unroll
                        for (int k = 0; k < D; ++k) q = fma(a2[ti][k], b[k], q);
                        double kap, ew;
                        kappa_and_dweight<KIND, 10>(q, s_tab, kap, ew);
                        const double om = fma(wi[ti], uj, ui[ti] * wj);
                        const double cw = ew * om;
                        gvar = fma(kap, om, gvar);
                        racc[ti] += cw;
                        cs += cw;
#pragma unroll
                        for (int k = 0; k < D; ++k) wq[k] = fma(cw, a2[ti][k], wq[k]);
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) gq[k] = fma(b[k], wq[k], gq[k]);
                    c[jj] = cs;
                }
                if (offdiag) {
                    col_reduce<kCG>(c, lane);
                    if ((lane & (32 / kCG - 1)) == 0) scol[jg + reduced_col<kCG>(lane)] = c[0];
                }
            }
            ring.consumer_release(lane);

            if (offdiag) {
                __syncthreads();
                if (tid < kBJ) {
                    const long j = c0 + (long)tile * kBJ + tid;
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) s += s_col[(colbuf * kWarps + w) * kBJ + tid];
                    if (j < args.ncols) atomicAdd(yb + j, s);
                }
                colbuf ^= 1;
            }
        }
        const double rscale = offdiag ? 1.0 : 2.0;
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
            if (live[ti]) atomicAdd(yb + r0 + ti * kThreads + tid, rscale * racc[ti]);
        __syncthreads();
        cc.next_item(args);
    }

    // block reduction of the thread-private accumulators -> one atomic per CTA per component
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double s = warp_sum(gq[k]);
        if (lane == 0) s_red[warp * (D + 2) + k] = s;
    }
    {
        double s = warp_sum(gvar);
        if (lane == 0) s_red[warp * (D + 2) + D] = s;
    }
    __syncthreads();
    if (tid <= D) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += s_red[w * (D + 2) + tid];
        args.gout[(long)blockIdx.x * args.gstride + tid] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// K7: dense cross-covariance  out[m][i] = variance * kappa(z_m, x_i)      (thread = column i)
// ---------------------------------------------------------------------------------------------
struct KnmArgs {
    const double* zp; long m;          // packed inducing points [m_pad][DP]
    const double* xp; long ncols;      // packed inputs          [n_pad][DP]
    double* out; long ld;              // build: M x ld output
    const double* exp_tab;
    double variance;
    // backward only
    const double* t; long ldt;         // dS/dKuf dense part (may be null)
    const double* wt;                  // [m]   rank-one part: G_mi = t_mi + wt_m * zvec_i
    const double* zvec;                // [ncols]
    const double* lengthscale;         // [D]
    double* out_ls;                    // [D]   accumulated
    double* out_var;                   // [1]   accumulated
    double* out_z;                     // [m][D] accumulated
    // backward, fixed summation order: CTA (bx, by) stores its sums (plain stores), knm_bwd_reduce_kernel adds them in bx order
    double* zpart;                     // [gridDim.x][m][D]
    double* lspart;                    // [gridDim.y][gridDim.x][D + 1]
    double cscale;                     // sqrt3 (Matern32) or 1/sqrt2 (RBF): packed = cscale * (x - shift) / l
    double cfac;                       // 1 (Matern32) or 2 (RBF)
};

constexpr int kKnmRows = 64;           // rows of Z per CTA (build)
// rows of Z per CTA in the backward: the per-warp partial slabs [8][rows][CGR] must fit in 48 KB of static smem
__host__ __device__ constexpr int knm_bwd_rows(int d) { return d <= 16 ? 32 : 16; }

template <int KIND, int D>
__global__ void __launch_bounds__(256) knm_build_kernel(const KnmArgs args) {
    constexpr int DP = SmemLayout<D>::DP;
    __shared__ __align__(16) double s_z[kKnmRows * DP];
    __shared__ double s_tab[64];
    const int tid = threadIdx.x;
    if (tid < 64) s_tab[tid] = args.exp_tab[tid];
    const long m0 = (long)blockIdx.y * kKnmRows;
    const int mr = (int)((args.m - m0) < kKnmRows ? (args.m - m0) : kKnmRows);
    for (int i = tid; i < mr * DP; i += blockDim.x) s_z[i] = args.zp[m0 * DP + i];
    __syncthreads();
    const long col = (long)blockIdx.x * blockDim.x + tid;
    if (col >= args.ncols) return;
    double a2[D], na;
    {
        const double2* src = reinterpret_cast<const double2*>(args.xp + col * DP);
        double tmp[DP];
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            double2 p = __ldg(src + h);
            tmp[2 * h] = p.x;
            tmp[2 * h + 1] = p.y;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) a2[k] = -2.0 * tmp[k];
        na = tmp[DP - 1];
    }
#pragma unroll 4
    for (int mm = 0; mm < mr; ++mm) {
        const double* b = s_z + mm * DP;
        double q = na + b[DP - 1];
#pragma unroll
        for (int k = 0; k < D; ++k) q = fma(a2[k], b[k], q);
        args.out[(m0 + mm) * args.ld + col] = args.variance * kappa<KIND>(q, s_tab);
    }
}

// backward of the cross-covariance: thread = column i, loops over a chunk of inducing points.
template <int KIND, int D>
__global__ void __launch_bounds__(256) knm_bwd_kernel(const KnmArgs args) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int CGR = D <= 1 ? 1 : D <= 2 ? 2 : D <= 4 ? 4 : D <= 8 ? 8 : D <= 16 ? 16 : 32;
    constexpr int kKnmBwdRows = knm_bwd_rows(D);
    __shared__ __align__(16) double s_z[kKnmBwdRows * DP];
    __shared__ double s_tab[64];
    __shared__ double s_wt[kKnmBwdRows];
    __shared__ double s_part[8][kKnmBwdRows][CGR];     // per-warp partial sums of g' * delta_k for every row
    __shared__ double s_red[8][D + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 64) s_tab[tid] = args.exp_tab[tid];
    const long m0 = (long)blockIdx.y * kKnmBwdRows;
    const int mr = (int)((args.m - m0) < kKnmBwdRows ? (args.m - m0) : kKnmBwdRows);
    for (int i = tid; i < mr * DP; i += blockDim.x) s_z[i] = args.zp[m0 * DP + i];
    for (int i = tid; i < mr; i += blockDim.x) s_wt[i] = args.wt ? args.wt[m0 + i] : 0.0;
    __syncthreads();
    const long col = (long)blockIdx.x * blockDim.x + tid;
    const bool live = col < args.ncols;
    double x[D];
    {
        const double2* src = reinterpret_cast<const double2*>(args.xp + (live ? col : 0) * DP);
        double tmp[DP];
#pragma unroll
        for (int h = 0; h < DP / 2; ++h) {
            double2 p = __ldg(src + h);
            tmp[2 * h] = p.x;
            tmp[2 * h + 1] = p.y;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = tmp[k];
    }
    const double zv = (live && args.zvec) ? args.zvec[col] : 0.0;
    const double vc = args.variance * args.cfac;
    double bl[D], bvar = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) bl[k] = 0.0;

    for (int mm = 0; mm < mr; ++mm) {
        const double* b = s_z + mm * DP;
        double del[D], q = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            del[k] = b[k] - x[k];
            q = fma(del[k], del[k], q);
        }
        double kap, ew;
        kappa_and_dweight<KIND>(q, s_tab, kap, ew);
        double G = s_wt[mm] * zv;
        if (args.t && live) G += args.t[(m0 + mm) * args.ldt + col];
        if (!live) G = 0.0;
        bvar = fma(G, kap, bvar);
        const double gp = G * ew * vc;
        double c[CGR];
#pragma unroll
        for (int k = 0; k < CGR; ++k) c[k] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double tk = gp * del[k];
            c[k] = tk;
            bl[k] = fma(tk, del[k], bl[k]);
        }
        col_reduce<CGR>(c, lane);
        if ((lane & (32 / CGR - 1)) == 0) s_part[warp][mm][reduced_col<CGR>(lane)] = c[0];
    }
    __syncthreads();
    // out_z[m][k] += -(cscale / l_k) * sum_warps
    for (int idx = tid; idx < mr * D; idx += blockDim.x) {
        const int mm = idx / D, k = idx % D;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_part[w][mm][k];
        if (args.out_z) args.zpart[((long)blockIdx.x * args.m + m0 + mm) * D + k] = -args.cscale / args.lengthscale[k] * s;
    }
    // lengthscale / variance sums: block reduce, one atomic per CTA per component
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double s = warp_sum(bl[k]);
        if (lane == 0) s_red[warp][k] = s;
    }
    {
        double s = warp_sum(bvar);
        if (lane == 0) s_red[warp][D] = s;
    }
    __syncthreads();
    if (tid <= D) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_red[w][tid];
        args.lspart[((long)blockIdx.y * gridDim.x + blockIdx.x) * (D + 1) + tid] = (tid < D) ? s / args.lengthscale[tid] : s;
    }
}

// second stage of the K_nm backward: out_z[i] += sum_bx zpart[bx][i]; out_ls / out_var += sum over the CTAs in (by, bx) order
static __global__ void knm_bwd_reduce_kernel(const double* __restrict__ zpart, long gx, long md, double* __restrict__ out_z,
                                      const double* __restrict__ lspart, long nctas, int d, double* __restrict__ out_ls,
                                      double* __restrict__ out_var) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_z != nullptr && i < md) {
        double s = 0.0;
#pragma unroll 8
        for (long b = 0; b < gx; ++b) s += zpart[b * md + i];
        out_z[i] += s;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x <= d) {
        double s = 0.0;
        for (long b = 0; b < nctas; ++b) s += lspart[b * (d + 1) + threadIdx.x];
        if ((int)threadIdx.x < d) out_ls[threadIdx.x] += s;
        else *out_var += s;
    }
}

template <int KIND, int D>
static int run_knm(Context* ctx, int bwd, KnmArgs a, cudaStream_t st) {
    if (a.m <= 0 || a.ncols <= 0) return CGLB_OK;
    const int rows = bwd ? knm_bwd_rows(D) : kKnmRows;
    dim3 grid((unsigned)((a.ncols + 255) / 256), (unsigned)((a.m + rows - 1) / rows));
    if (!bwd) {
        knm_build_kernel<KIND, D><<<grid, 256, 0, st>>>(a);
        ctx->launches++;
        CGLB_LAUNCH_OK();
        return CGLB_OK;
    }
    const long md = a.m * D, nctas = (long)grid.x * grid.y;
    const long zwords = a.out_z ? (long)grid.x * md : 0;
    int rc = ensure_ypart(ctx, zwords + nctas * (D + 1));
    if (rc) return rc;
    a.zpart = ctx->ypart;
    a.lspart = ctx->ypart + zwords;
    knm_bwd_kernel<KIND, D><<<grid, 256, 0, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    const long work = a.out_z ? md : 1;
    knm_bwd_reduce_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(a.zpart, grid.x, md, a.out_z, a.lspart, nctas, D, a.out_ls, a.out_var);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

// ---------------------------------------------------------------------------------------------
// per-dimension launchers (one translation unit per D, see kmv_inst.cu)
// ---------------------------------------------------------------------------------------------
template <int D, int WARPS>
static size_t fwd_smem_bytes() {
    constexpr int DP = SmemLayout<D>::DP;
    return (size_t)(kStages * kBJ * DP + kStages * kBJ + 2 * WARPS * kBJ + kExpTabBig) * sizeof(double) + 2 * kStages * sizeof(uint64_t);
}
template <int D, int WARPS>
static size_t bwd_smem_bytes() {
    constexpr int DP = SmemLayout<D>::DP;
    return (size_t)(kStages * kBJ * DP + kStages * 2 * kBJ + 2 * WARPS * kBJ + kExpTabBig + WARPS * (D + 2)) * sizeof(double) +
           2 * kStages * sizeof(uint64_t);
}

template <int KIND, int D, int TI, bool SYM, int WARPS, int CB>
static int launch_fwd(Context* ctx, SweepArgs a, cudaStream_t st) {
    constexpr long BI = WARPS * 32 * TI;
    a.nb_rows = (a.nrows + BI - 1) / BI;
    a.nb_cols = (a.ncols + BI - 1) / BI;
    a.nitems = SYM ? a.nb_rows * (a.nb_rows + 1) / 2 : a.nb_rows * a.nb_cols;
    auto kern = kmv_sweep_kernel<KIND, D, TI, SYM, WARPS, CB>;
    size_t smem = fwd_smem_bytes<D, WARPS>();
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

template <int KIND, int D, int TI, int WARPS>
static int launch_bwd(Context* ctx, SweepArgs a, cudaStream_t st) {
    constexpr long BI = WARPS * 32 * TI;
    a.nb_rows = (a.nrows + BI - 1) / BI;
    a.nb_cols = a.nb_rows;
    a.nitems = a.nb_rows * (a.nb_rows + 1) / 2;
    auto kern = kmv_bwd_kernel<KIND, D, TI, WARPS>;
    size_t smem = bwd_smem_bytes<D, WARPS>();
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

// number of work items a (rows-per-CTA) choice yields for this launch
static inline long count_items(long nrows, long ncols, bool sym, int nparts, long bi) {
    long nbr = (nrows + bi - 1) / bi, nbc = (ncols + bi - 1) / bi;
    return (sym ? nbr * (nbr + 1) / 2 : nbr * nbc) / nparts;
}

#ifdef CGLB_KMV_EXPERIMENT
// developer build: every (WARPS, TI) variant, selected with the environment variable CGLB_KMV_VARIANT
template <int KIND, int D, bool SYM>
static int run_fwd(Context* ctx, const SweepArgs& a, cudaStream_t st) {
    const char* e = getenv("CGLB_KMV_VARIANT");
    int v = e ? atoi(e) : 0;      // WARPS*100 + TI*10 + CB
    switch (v) {
        case 841: return launch_fwd<KIND, D, 4, SYM, 8, 1>(ctx, a, st);
        case 842: return launch_fwd<KIND, D, 4, SYM, 8, 2>(ctx, a, st);
        case 832: return launch_fwd<KIND, D, 3, SYM, 8, 2>(ctx, a, st);
        case 824: return launch_fwd<KIND, D, 2, SYM, 8, 4>(ctx, a, st);
        case 1222: return launch_fwd<KIND, D, 2, SYM, 12, 2>(ctx, a, st);
        case 1224: return launch_fwd<KIND, D, 2, SYM, 12, 4>(ctx, a, st);
        case 1232: return launch_fwd<KIND, D, 3, SYM, 12, 2>(ctx, a, st);
        case 1622: return launch_fwd<KIND, D, 2, SYM, 16, 2>(ctx, a, st);
        case 1614: return launch_fwd<KIND, D, 1, SYM, 16, 4>(ctx, a, st);
        default: return launch_fwd<KIND, D, 4, SYM, 8, 1>(ctx, a, st);
    }
}
template <int KIND, int D>
static int run_bwd(Context* ctx, const SweepArgs& a, cudaStream_t st) {
    const char* e = getenv("CGLB_BWD_VARIANT");
    int v = e ? atoi(e) : 0;
    switch (v) {
        case 82: return launch_bwd<KIND, D, 2, 8>(ctx, a, st);
        case 81: return launch_bwd<KIND, D, 1, 8>(ctx, a, st);
        case 122: return launch_bwd<KIND, D, 2, 12>(ctx, a, st);
        case 121: return launch_bwd<KIND, D, 1, 12>(ctx, a, st);
        case 161: return launch_bwd<KIND, D, 1, 16>(ctx, a, st);
        default: return launch_bwd<KIND, D, 2, 8>(ctx, a, st);
    }
}
#else
// Rows per CTA: big blocks amortise the column loads best; small problems need more items than SMs.
template <int KIND, int D, bool SYM>
static int run_fwd(Context* ctx, const SweepArgs& a, cudaStream_t st) {
    constexpr bool BIG = SYM && (D <= 12);
    if constexpr (BIG) {
        if (count_items(a.nrows, a.ncols, SYM, a.nparts, 8 * 32 * 4) >= 12L * ctx->num_sms) return launch_fwd<KIND, D, 4, SYM, 8, 1>(ctx, a, st);
    }
    if (count_items(a.nrows, a.ncols, SYM, a.nparts, 8 * 32 * 2) >= 12L * ctx->num_sms) return launch_fwd<KIND, D, 2, SYM, 8, 2>(ctx, a, st);
    return launch_fwd<KIND, D, 1, SYM, 8, 4>(ctx, a, st);
}

template <int KIND, int D>
static int run_bwd(Context* ctx, const SweepArgs& a, cudaStream_t st) {
    if constexpr (D <= 12) {
        if (count_items(a.nrows, a.ncols, true, a.nparts, 8 * 32 * 2) >= 12L * ctx->num_sms) return launch_bwd<KIND, D, 2, 8>(ctx, a, st);
    }
    return launch_bwd<KIND, D, 1, 8>(ctx, a, st);
}
#endif

// dispatch table filled by the per-D translation units
typedef int (*sweep_fn)(Context*, int kind, int mode /*0 sym fwd, 1 rect fwd, 2 sym bwd, 3 sym fwd on DMMA, 4 sym bwd on DMMA, 5/6 sym fwd with 2/4 right-hand sides*/, const SweepArgs&, cudaStream_t);
typedef int (*knm_fn)(Context*, int kind, int bwd, const KnmArgs&, cudaStream_t);
knm_fn get_knm_fn(int d);
constexpr int kMaxRegisterD = 32;
sweep_fn get_sweep_fn(int d);

}  // namespace cglb
