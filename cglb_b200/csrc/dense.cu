// K5 / K6: dense FP64 linear algebra for the Nystrom preconditioner and log-det terms, sm_100a.
//
//   cglb_gemm             C = alpha A op(B) + beta C           DMMA.8x8x4 tiles, cp.async 3-stage ring
//   cglb_syrk             C = A A^T (split-K, lower tiles + mirror; the K slices are summed in slice order)
//   cglb_potrf            blocked right-looking Cholesky (128-wide panels)
//   cglb_tri_inverse      L^-1 through inverted diagonal blocks + GEMMs
//   cglb_trsm_left_lower  B <- alpha L^-1 B, one GEMM per 128-row block against [ -D^-1 L | alpha D^-1 ]
//
// Replaces torch.cholesky / torch.triangular_solve / A @ A^T at reference models.py:202-210
// (cuSOLVER dpotrf, cuBLAS dtrsm/dgemm).  FP64 has no tcgen05 form; mma.sync.m8n8k4.f64 (DMMA) measured
// 37.1 TFLOP/s on this B200 vs 34.2 for plain DFMA (profiles/fp64_peaks_r01.json), so the contraction
// runs on DMMA.
#include <stdlib.h>
#include <type_traits>

#include "kmv_impl.cuh"

namespace cglb {

constexpr int GM = 128, GN = 128, GK = 16;
constexpr int GPITCH_K = GK + 4;     // As[m][k], Bs(NT)[n][k]: pitch 20 doubles -> conflict-free fragment loads
constexpr int GPITCH_N = GN + 4;     // Bs(NN)[k][n]: pitch 132
constexpr int GSTAGES = 3;
constexpr int GTHREADS = 256;
constexpr int NB = 128;              // block size of the blocked factorisations

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_SYRK = 2, EPI_KMAP = 3, EPI_KBWD = 4 };

// extra operands of the kernel-map epilogues (wide-input K_nm build / backward): the GEMM computes
// S = Zp Xp^T, the epilogue turns it into variance*kappa(|z|^2+|x|^2-2S) (EPI_KMAP) or into
// GP = (T + wt zvec^T) * e' * variance * cfac written over T, with sum G*kappa, row sums and column sums (EPI_KBWD)
struct KEpiArgs {
    const double* nz; long nz_stride;     // |z_m|^2 at nz[m * nz_stride]
    const double* nx; long nx_stride;
    const double* exp_tab;
    const double* wt; const double* zvec;  // may be null
    double* rsum; double* csum; double* gk_sum;   // EPI_KBWD partial slots: [column tiles][m], [row tiles][n], [row tiles][column tiles]
    double variance, vc;
    int kind;
};

struct GemmArgs {
    const double* A; long lda;
    const double* B; long ldb;
    double* C; long ldc;
    long m, n, k;
    long k_chunk;         // K range handled per blockIdx.z
    double alpha, beta;
    int lower_only;       // skip tiles strictly above the block diagonal
    KEpiArgs ke;
    // split-K (EPI_ATOMIC / EPI_SYRK): slice z stores alpha * (its partial product) at part + z * part_stride, row pitch n;
    // splitk_reduce_kernel sums the slices in slice order (fixed summation order, no atomics)
    double* part; long part_stride;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
// 16-byte chunk = two doubles; falls back to two 8-byte copies when the operand is not 16-byte aligned
__device__ __forceinline__ void cp_chunk(bool vec, double* smem, const double* src, int bytes) {
    if (vec) {
        cp_async16(smem, src, bytes);
    } else {
        cp_async8(smem, src, bytes >= 8 ? 8 : 0);
        cp_async8(smem + 1, bytes == 16 ? src + 1 : src, bytes == 16 ? 8 : 0);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <bool TRANSB>
struct GemmSmem {
    static constexpr int kA = GM * GPITCH_K;
    static constexpr int kB = TRANSB ? GN * GPITCH_K : GK * GPITCH_N;
    static constexpr size_t bytes = (size_t)GSTAGES * (kA + kB) * sizeof(double);
};

// Epilogue shared by the two GEMM kernels: lane (g, t) of warp (wm, wn) holds rows wm*64 + 8i + g, columns wn*32 + 8j + 2t + e
// of the 128 x 128 tile (bm, bn).  `active` is false for threads that only take part in the CTA barriers (the TMA producer warp).
template <int EPI>
__device__ __forceinline__ void gemm_epilogue(const GemmArgs& p, double (&acc)[8][4][2], int bm, int bn, int tid, bool active) {
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) & 1, wn = warp & 3;
    const long m0 = (long)bm * GM, n0 = (long)bn * GN;
    if (EPI == EPI_KMAP || EPI == EPI_KBWD) {
        __shared__ double s_tab[64];
        __shared__ double s_rs[4][GM], s_cs[2][GN], s_gk[8];      // EPI_KBWD: per-warp row / column / G*kappa sums
        __syncthreads();
        if (tid < 64) s_tab[tid] = p.ke.exp_tab[tid];
        __syncthreads();
        double gk = 0.0, racc[8], cacc[4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) racc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) cacc[j][0] = cacc[j][1] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long row = m0 + wm * 64 + i * 8 + g;
            const bool rlive = row < p.m;
            const double nzr = rlive ? p.ke.nz[row * p.ke.nz_stride] : 0.0;
            const double wtr = (EPI == EPI_KBWD && rlive && p.ke.wt) ? p.ke.wt[row] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const long cc = n0 + wn * 32 + j * 8 + 2 * t + e;
                    if (!rlive || cc >= p.n) continue;
                    const double q = fma(-2.0, acc[i][j][e], nzr + p.ke.nx[cc * p.ke.nx_stride]);
                    double kap, ew;
                    if (p.ke.kind == CGLB_MATERN32) kappa_and_dweight<CGLB_MATERN32>(q, s_tab, kap, ew);
                    else kappa_and_dweight<CGLB_RBF>(q, s_tab, kap, ew);
                    double* dst = p.C + row * p.ldc + cc;
                    if (EPI == EPI_KMAP) {
                        *dst = p.ke.variance * kap;
                    } else {
                        double G = (p.beta != 0.0) ? *dst : 0.0;            // beta != 0: a dense T is present
                        if (p.ke.zvec) G = fma(wtr, p.ke.zvec[cc], G);
                        const double gp = G * ew * p.ke.vc;
                        *dst = gp;
                        gk = fma(G, kap, gk);
                        racc[i] += gp;
                        cacc[j][e] += gp;
                    }
                }
            }
        }
        if (EPI == EPI_KBWD) {
            // rows: reduce over the 4 lanes sharing g; columns: over the 8 lanes sharing t
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double s = racc[i];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (t == 0) s_rs[wn][wm * 64 + i * 8 + g] = s;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double s = cacc[j][e];
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (g == 0) s_cs[wm][wn * 32 + j * 8 + 2 * t + e] = s;
                }
            gk = warp_sum(gk);
            if (lane == 0) s_gk[warp] = gk;
        }
        if (EPI == EPI_KBWD) {
            // fixed summation order: the warps' sums are combined in warp order, every CTA stores its row / column / G*kappa
            // sums in its own slot (rsum: [column tile][m], csum: [row tile][n], gk: [row tile][column tile]);
            // kbwd_reduce_kernel adds the slots in tile order
            __syncthreads();
            if (tid < GM) {
                const long row = m0 + tid;
                if (row < p.m) p.ke.rsum[(long)bn * p.m + row] = s_rs[0][tid] + s_rs[1][tid] + s_rs[2][tid] + s_rs[3][tid];
            } else if (tid < GM + GN) {
                const long cc = n0 + (tid - GM);
                if (cc < p.n) p.ke.csum[(long)bm * p.n + cc] = s_cs[0][tid - GM] + s_cs[1][tid - GM];
            }
            if (tid == 0) {
                double s = 0.0;
                for (int w8 = 0; w8 < 8; ++w8) s += s_gk[w8];
                p.ke.gk_sum[(long)bm * gridDim.x + bn] = s;
            }
        }
        return;
    }
    if (!active) return;
    // epilogue: lane holds (row g, cols 2t, 2t+1) of every 8x8 tile.  Per row group the (optional) reads of C
    // are issued together before the dependent stores, as 16-byte accesses when C allows it.
    const bool c_vec = ((p.ldc & 1) == 0) && (((uintptr_t)p.C & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long row = m0 + wm * 64 + i * 8 + g;
        if (row >= p.m) continue;
        if (EPI == EPI_STORE) {
            double* crow = p.C + row * p.ldc;
            double2 old[4];
            bool full[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long col = n0 + wn * 32 + j * 8 + 2 * t;
                full[j] = c_vec && (col + 1 < p.n);
                old[j] = make_double2(0.0, 0.0);
                if (p.beta != 0.0) {
                    if (full[j]) old[j] = *reinterpret_cast<const double2*>(crow + col);
                    else {
                        if (col < p.n) old[j].x = crow[col];
                        if (col + 1 < p.n) old[j].y = crow[col + 1];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long col = n0 + wn * 32 + j * 8 + 2 * t;
                double2 v;
                v.x = fma(p.beta, old[j].x, p.alpha * acc[i][j][0]);
                v.y = fma(p.beta, old[j].y, p.alpha * acc[i][j][1]);
                if (full[j]) *reinterpret_cast<double2*>(crow + col) = v;
                else {
                    if (col < p.n) crow[col] = v.x;
                    if (col + 1 < p.n) crow[col + 1] = v.y;
                }
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long col = n0 + wn * 32 + j * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long cc = col + e;
                if (cc >= p.n) continue;
                p.part[(long)blockIdx.z * p.part_stride + row * p.n + cc] = p.alpha * acc[i][j][e];
            }
        }
    }
}

// C tile (bm, bn) of size 128x128; 8 warps as 2 (m) x 4 (n), warp tile 64x32 = 8x4 DMMA tiles.
template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_cpasync_kernel(const GemmArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + GSTAGES * GemmSmem<TRANSB>::kA;

    const int bm = blockIdx.y, bn = blockIdx.x;
    if (p.lower_only && bn > bm) return;
    const long m0 = (long)bm * GM, n0 = (long)bn * GN;
    const long kbeg = (long)blockIdx.z * p.k_chunk;
    long kend = kbeg + p.k_chunk;
    if (kend > p.k) kend = p.k;
    const int nkt = (int)((kend - kbeg + GK - 1) / GK);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;

    const bool a_vec = ((p.lda & 1) == 0) && (((uintptr_t)p.A & 15) == 0);
    const bool b_vec = ((p.ldb & 1) == 0) && (((uintptr_t)p.B & 15) == 0);
    auto load_tile = [&](int kt, int stage) {
        const long k0 = kbeg + (long)kt * GK;
        double* a = sA + stage * GemmSmem<TRANSB>::kA;
        double* b = sB + stage * GemmSmem<TRANSB>::kB;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GTHREADS;        // 1024 chunks of 16 B
            const int row = c >> 3, kc = (c & 7) * 2;
            const long gr = m0 + row, gk = k0 + kc;
            long rem = kend - gk;
            int bytes = (gr < p.m && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
            const double* src = bytes ? p.A + gr * p.lda + gk : p.A;
            cp_chunk(a_vec, a + row * GPITCH_K + kc, src, bytes);
        }
        if (TRANSB) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid + i * GTHREADS;
                const int row = c >> 3, kc = (c & 7) * 2;
                const long gr = n0 + row, gk = k0 + kc;
                long rem = kend - gk;
                int bytes = (gr < p.n && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
                const double* src = bytes ? p.B + gr * p.ldb + gk : p.B;
                cp_chunk(b_vec, b + row * GPITCH_K + kc, src, bytes);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid + i * GTHREADS;
                const int row = c >> 6, nc = (c & 63) * 2;
                const long gk = k0 + row, gc = n0 + nc;
                long rem = p.n - gc;
                int bytes = (gk < kend && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
                const double* src = bytes ? p.B + gk * p.ldb + gc : p.B;
                cp_chunk(b_vec, b + row * GPITCH_N + nc, src, bytes);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < nkt) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        if (kt + GSTAGES - 1 < nkt) load_tile(kt + GSTAGES - 1, (kt + GSTAGES - 1) % GSTAGES);
        cp_async_commit();
        const double* a = sA + (kt % GSTAGES) * GemmSmem<TRANSB>::kA + (wm * 64 + g) * GPITCH_K + t;
        const double* b = TRANSB ? sB + (kt % GSTAGES) * GemmSmem<TRANSB>::kB + (wn * 32 + g) * GPITCH_K + t
                                 : sB + (kt % GSTAGES) * GemmSmem<TRANSB>::kB + t * GPITCH_N + wn * 32 + g;
#pragma unroll
        for (int k4 = 0; k4 < GK / 4; ++k4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = a[i * 8 * GPITCH_K + k4 * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = TRANSB ? b[j * 8 * GPITCH_K + k4 * 4] : b[k4 * 4 * GPITCH_N + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    gemm_epilogue<EPI>(p, acc, bm, bn, tid, true);
}

// ---------------------------------------------------------------------------------------------
// The same GEMM with TMA operand staging (round 2): the operand tiles arrive through 1-D bulk copies (cp.async.bulk ->
// UBLKCP, one per tile row, straight into the padded bank-conflict-free layouts above) on a GT_STAGES-deep full/empty
// mbarrier ring.  Every warp stages 1/8 of the rows of a k-tile (one row per lane: lanes 0-15 a row of A, lanes 16-31 a row
// of B; one mbarrier arrival per warp) GT_STAGES - 1 tiles ahead, so the main loop has no CTA barrier, no cp.async address arithmetic (4 + 4 16-byte
// copies per thread and k-tile before) and no wait_group; warps only meet through the barriers of a stage S - 1 tiles back.
// (A ninth, dedicated producer warp would cap the kernel at 168 registers: three warps on one scheduler.)
// Rows past the matrix and the K tail are written by the lane itself (zeros / the few valid doubles).  Needs 16-byte
// aligned operand rows (lda, ldb even, aligned bases); anything else takes gemm_cpasync_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int GT_STAGES = 5;
constexpr int GT_AHEAD = GT_STAGES - 2;      // tiles in flight ahead of the one being multiplied: a stage is refilled two tiles after
                                            // its last use, so a warp never waits for the warps that are one tile behind it

template <bool TRANSB>
struct GemmTmaSmem {
    static constexpr int kA = GM * GPITCH_K;
    static constexpr int kB = TRANSB ? GN * GPITCH_K : GK * GPITCH_N;
    static constexpr size_t bytes = (size_t)GT_STAGES * (kA + kB) * sizeof(double) + 2 * GT_STAGES * sizeof(uint64_t);
};

template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_kernel(const GemmArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + GT_STAGES * GemmTmaSmem<TRANSB>::kA;
    uint64_t* s_full = reinterpret_cast<uint64_t*>(sB + GT_STAGES * GemmTmaSmem<TRANSB>::kB);
    uint64_t* s_empty = s_full + GT_STAGES;

    // tile order: the row tile runs fastest, so the CTAs in flight together share one column panel of B (L2-resident) and
    // all of A's row tiles; with bn fastest the (M x M)(M x n) product read B 16 times from DRAM (26.6 GB for 1.6 GB, ncu)
    const long lin = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const int bm = (int)(lin % gridDim.y), bn = (int)(lin / gridDim.y);
    if (p.lower_only && bn > bm) return;
    const long m0 = (long)bm * GM, n0 = (long)bn * GN;
    const long kbeg = (long)blockIdx.z * p.k_chunk;
    long kend = kbeg + p.k_chunk;
    if (kend > p.k) kend = p.k;
    const int nkt = (int)((kend - kbeg + GK - 1) / GK);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&s_full[s], GTHREADS / 32); mbar_init(&s_empty[s], GTHREADS / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    // this lane's row of every k-tile: lanes 0-15 -> row 16 warp + lane of A; lanes 16-31 -> a row of B
    long nval = p.n - n0;
    if (nval > GN) nval = GN;
    const bool is_a = lane < 16;
    const int arow = warp * 16 + (lane & 15);                           // A (and NT B) tile row of this lane
    const bool b_nn_lane = !TRANSB && !is_a && (lane & 15) < 2;         // NN: 16 rows of B per k-tile, 2 per warp
    const int brow_nn = warp * 2 + (lane & 15);
    int pstage = 0;
    uint32_t pphase = 0;
    auto stage_tile = [&](int kt) {
        // wait until every warp has released the tile that lived in this stage, then fill this lane's row
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long k0 = kbeg + (long)kt * GK;
        const int kval = (int)((kend - k0) < GK ? (kend - k0) : GK);
        double* dst = nullptr;
        const double* src = nullptr;
        int len = 0, valid = 0;
        if (is_a) {
            const long gr = m0 + arow;
            dst = sA + pstage * GemmTmaSmem<TRANSB>::kA + arow * GPITCH_K;
            src = p.A + (gr < p.m ? gr : 0) * p.lda + k0;
            len = GK; valid = gr < p.m ? kval : 0;
        } else if (TRANSB) {
            const long gr = n0 + arow;
            dst = sB + pstage * GemmTmaSmem<TRANSB>::kB + arow * GPITCH_K;
            src = p.B + (gr < p.n ? gr : 0) * p.ldb + k0;
            len = GK; valid = gr < p.n ? kval : 0;
        } else if (b_nn_lane) {
            const long gk = k0 + brow_nn;
            dst = sB + pstage * GemmTmaSmem<TRANSB>::kB + brow_nn * GPITCH_N;
            src = p.B + (gk < kend ? gk : kbeg) * p.ldb + n0;
            len = GN; valid = gk < kend ? (int)nval : 0;
        }
        // partial / out-of-range rows are written by the lane itself; full rows go to the TMA unit.  One arrival per warp:
        // lane 0 arrives with the bytes of the warp's full rows after the warp's plain stores (release), then the copies go out.
        const bool full = len > 0 && valid == len;
        if (!full)
            for (int i = 0; i < len; ++i) dst[i] = (i < valid) ? src[i] : 0.0;
        const unsigned fmask = __ballot_sync(0xffffffffu, full);
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(__popc(fmask & 0xffffu) * GK * sizeof(double)) +
                                   (uint32_t)(__popc(fmask >> 16) * (TRANSB ? GK : GN) * sizeof(double));
            if (bytes) mbar_expect_tx(&s_full[pstage], bytes);
            else mbar_arrive(&s_full[pstage]);
        }
        __syncwarp();
        if (full) tma_load_1d(dst, src, (uint32_t)(len * sizeof(double)), &s_full[pstage]);
        if (++pstage == GT_STAGES) { pstage = 0; pphase ^= 1; }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll 1
    for (int kt = 0; kt < GT_AHEAD && kt < nkt; ++kt) stage_tile(kt);
    int stage = 0;
    uint32_t phase = 0;
    // (loading the fragments one k4-step ahead across the k-tile boundary through a second register set was measured and is
    // slower: 28.6 vs 31.7 TFLOP/s at 2048 x 200k x 2048 -- 214-222 registers instead of 178-198)
#pragma unroll 1
    for (int kt = 0; kt < nkt; ++kt) {
        if (kt + GT_AHEAD < nkt) stage_tile(kt + GT_AHEAD);
        mbar_wait(&s_full[stage], phase);
        const double* a = sA + stage * GemmTmaSmem<TRANSB>::kA + (wm * 64 + g) * GPITCH_K + t;
        const double* b = TRANSB ? sB + stage * GemmTmaSmem<TRANSB>::kB + (wn * 32 + g) * GPITCH_K + t
                                 : sB + stage * GemmTmaSmem<TRANSB>::kB + t * GPITCH_N + wn * 32 + g;
#pragma unroll
        for (int k4 = 0; k4 < GK / 4; ++k4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = a[i * 8 * GPITCH_K + k4 * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = TRANSB ? b[j * 8 * GPITCH_K + k4 * 4] : b[k4 * 4 * GPITCH_N + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
        if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
    }
    gemm_epilogue<EPI>(p, acc, bm, bn, tid, true);
}

template <bool TRANSB, int EPI>
static int launch_gemm(Context* ctx, const GemmArgs& p, int ksplit, cudaStream_t st) {
    if (p.m <= 0 || p.n <= 0) return CGLB_OK;
    dim3 grid((unsigned)((p.n + GN - 1) / GN), (unsigned)((p.m + GM - 1) / GM), (unsigned)ksplit);
    if (grid.y > 65535 || grid.z > 65535) {
        set_error("gemm: m=%ld too large for grid.y", p.m);
        return CGLB_ERR_UNSUPPORTED;
    }
    // bulk copies need 16-byte aligned rows (and k-tiles that start on an even column: k_chunk is a multiple of GK)
    // Default: the cp.async ring.  The TMA-staged kernel is selected with cglb_set_option("gemm_staging", 2): it matches the
    // cp.async kernel on the one big (M x M)(M x n) product (32.0 vs 31.8 TFLOP/s) and loses everywhere else (TRSM block rows
    // 13.4 vs 26.4 TFLOP/s at 2048 x 54k, the kin40k-shaped step 23.4 vs 21.6 ms).
    const bool tma_ok = ctx->opt_gemm_staging == 2 && ((p.lda & 1) == 0) && ((p.ldb & 1) == 0) && (((uintptr_t)p.A & 15) == 0) &&
                        (((uintptr_t)p.B & 15) == 0);
    if (tma_ok) {
        auto kern = gemm_kernel<TRANSB, EPI>;
        size_t smem = GemmTmaSmem<TRANSB>::bytes;
        CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, GTHREADS, smem, st>>>(p);
    } else {
        auto kern = gemm_cpasync_kernel<TRANSB, EPI>;
        size_t smem = GemmSmem<TRANSB>::bytes;
        CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, GTHREADS, smem, st>>>(p);
    }
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

// C = beta C + sum_z part[z]  (slices in order);  sym: only col <= row is computed (lower tiles) and mirrored
__global__ void splitk_reduce_kernel(const double* __restrict__ part, long part_stride, int ksplit, long m, long n,
                                     double* __restrict__ c, long ldc, double beta, int sym) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * n) return;
    const long row = idx / n, col = idx % n;
    if (sym && col > row) return;
    double s = (beta == 0.0) ? 0.0 : beta * c[row * ldc + col];
    for (int z = 0; z < ksplit; ++z) s += part[(long)z * part_stride + idx];
    c[row * ldc + col] = s;
    if (sym && col != row) c[col * ldc + row] = s;
}

template <bool TRANSB, int EPI>
static int launch_gemm_splitk(Context* ctx, GemmArgs p, int ksplit, int sym, cudaStream_t st) {
    p.part_stride = p.m * p.n;
    int rc = ensure_ypart(ctx, (long)ksplit * p.part_stride);
    if (rc) return rc;
    p.part = ctx->ypart;
    rc = launch_gemm<TRANSB, EPI>(ctx, p, ksplit, st);
    if (rc) return rc;
    splitk_reduce_kernel<<<(unsigned)((p.m * p.n + 255) / 256), 256, 0, st>>>(p.part, p.part_stride, ksplit, p.m, p.n, p.C, p.ldc, p.beta, sym);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

// C = alpha A op(B) + beta C where C does not alias A or B: few output tiles and a long K are split over
// blockIdx.z (every slice stores its partial product, summed in slice order afterwards) so that all SMs work.
template <bool TRANSB>
static int launch_gemm_auto(Context* ctx, GemmArgs p, cudaStream_t st) {
    const long tiles = ((p.m + GM - 1) / GM) * ((p.n + GN - 1) / GN);
    long ksplit = 1;
    if (tiles > 0 && tiles * 2 <= ctx->num_sms && p.k >= 512) {
        ksplit = ctx->num_sms / tiles;
        const long max_split = p.k / 256;
        if (ksplit > max_split) ksplit = max_split;
    }
    if (ksplit <= 1) return launch_gemm<TRANSB, EPI_STORE>(ctx, p, 1, st);
    p.k_chunk = ((p.k + ksplit - 1) / ksplit + GK - 1) / GK * GK;
    ksplit = (p.k + p.k_chunk - 1) / p.k_chunk;
    return launch_gemm_splitk<TRANSB, EPI_ATOMIC>(ctx, p, (int)ksplit, 0, st);
}

static int gemm_checked(Context* ctx, int transb, long m, long n, long k, double alpha, const double* a, long lda,
                        const double* b, long ldb, double beta, double* c, long ldc, cudaStream_t st) {
    GemmArgs p{a, lda, b, ldb, c, ldc, m, n, k, (k + GK - 1) / GK * GK, alpha, beta, 0, {}};
    if (p.k_chunk == 0) p.k_chunk = GK;
    const bool alias = (c == a) || (c == b);
    if (alias) return transb ? launch_gemm<true, EPI_STORE>(ctx, p, 1, st) : launch_gemm<false, EPI_STORE>(ctx, p, 1, st);
    return transb ? launch_gemm_auto<true>(ctx, p, st) : launch_gemm_auto<false>(ctx, p, st);
}

// ---------------------------------------------------------------------------------------------
// diagonal blocks: Cholesky and/or inverse of one NB x NB block per CTA, all in shared memory.
// Layout: one (NB x (NB+1)) array; strictly-lower part = L, upper part incl. diagonal = (L^-1)^T,
// L's diagonal (and its reciprocal) in separate vectors.
// ---------------------------------------------------------------------------------------------
constexpr int DP1 = NB + 1;

// mode 0: factor the block (a <- L, upper zeroed) and write inv(L) to dinv; mode 1: block already
// holds L (lower), only invert.  nb_actual = rows in this block (<= NB).  info: 1+block on failure.
//
// Register-tiled: 256 threads form a 16 x 16 grid, thread (ty, tx) owns the 8 x 8 elements
// (i, k) = (ty + 16 a, tx + 16 b) of the working matrix in registers.  Each of the 128 column steps costs ONE
// CTA barrier: the owners of column j publish it (unscaled) in shared memory, everybody derives 1/pivot
// redundantly and updates its own registers (right-looking); the loops over the 16-column groups are unrolled
// so that finished register tiles drop out statically.  The inverse is a forward substitution on the identity
// with the same ownership (B in registers, row j of X published per step).  The first version kept the matrix
// in shared memory (5 barriers + dependent read-modify-write chains per column, 1024 threads: 212 us per
// block); with 8 warps the per-column instruction overhead is paid 8 times instead of 32.
constexpr int DG = 16;               // thread grid is DG x DG, register tile is (NB/DG) x (NB/DG)
constexpr int DT = NB / DG;          // 8

__device__ __forceinline__ double rsqrt_full(double d) {
    // MUFU.RSQ64H seed (2^-20) + one third-order and one Newton step: full fp64 accuracy
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d * y, y, 1.0);
    y = fma(y * fma(e, 0.375, 0.5), e, y);
    e = fma(-d * y, y, 1.0);
    return fma(0.5 * y, e, y);
}

// Cholesky steps j = 16 JB .. 16 JB + 15
template <int JB>
__device__ __forceinline__ void chol_group(double (&r)[DT][DT], double* s, double* sr, double* buf, int ty, int tx, int* s_fail) {
#pragma unroll 1
    for (int jx = 0; jx < DG; ++jx) {
        const int j = JB * DG + jx;
        double* col = buf + (j & 1) * NB;
        if (tx == jx) {      // owners of column j publish rows ty + 16 a (a >= JB holds the rows >= 16 JB)
#pragma unroll
            for (int aa = JB; aa < DT; ++aa) col[ty + DG * aa] = r[aa][JB];
        }
        __syncthreads();
        double d = col[j];
        if (!(d > 0.0)) {
            *s_fail = 1;
            d = 1.0;
        }
        const double rinv = rsqrt_full(d);
        const double dinv = rinv * rinv;
        double ci[DT], ck[DT];
#pragma unroll
        for (int aa = JB; aa < DT; ++aa) ci[aa] = col[ty + DG * aa] * dinv;      // a_ij / d
#pragma unroll
        for (int bb = JB; bb < DT; ++bb) ck[bb] = col[tx + DG * bb];             // a_kj
        // trailing update r[i][k] -= a_ij a_kj / d for j < k <= i
#pragma unroll
        for (int bb = JB; bb < DT; ++bb)
#pragma unroll
            for (int aa = bb; aa < DT; ++aa) {
                bool on = true;
                if (bb == JB) on = on && (tx > jx);
                if (aa == bb) on = on && (tx <= ty);
                if (on) r[aa][bb] = fma(-ci[aa], ck[bb], r[aa][bb]);
            }
        if (tx == jx) {      // column j of L is final: l_ij = a_ij / sqrt(d), l_jj = sqrt(d)
#pragma unroll
            for (int aa = JB; aa < DT; ++aa) {
                const int i = ty + DG * aa;
                if (i > j) s[i * DP1 + j] = col[i] * rinv;
                else if (i == j) {
                    double sq = d * rinv;
                    s[i * DP1 + j] = fma(fma(-sq, sq, d), 0.5 * rinv, sq);      // sqrt(d), one correction
                    sr[j] = rinv;
                }
            }
        }
    }
}

// forward-substitution steps j = 16 JB .. 16 JB + 15 of X = L^-1
template <int JB>
__device__ __forceinline__ void inv_group(double (&bm)[DT][DT], const double* s, const double* sr, double* buf, int ty, int tx) {
#pragma unroll 1
    for (int jx = 0; jx < DG; ++jx) {
        const int j = JB * DG + jx;
        double* row = buf + (j & 1) * NB;
        if (ty == jx) {      // owners of row j: X[j][c] = B[j][c] / L[j][j], final; columns c <= j live in b <= JB
            const double rinv = sr[j];
#pragma unroll
            for (int bb = 0; bb <= JB; ++bb) {
                bm[JB][bb] *= rinv;
                row[tx + DG * bb] = bm[JB][bb];
            }
        }
        __syncthreads();
        double lj[DT], xj[DT];
#pragma unroll
        for (int aa = JB; aa < DT; ++aa) lj[aa] = s[(ty + DG * aa) * DP1 + j];
#pragma unroll
        for (int bb = 0; bb <= JB; ++bb) xj[bb] = row[tx + DG * bb];
        // B[i][c] -= L[i][j] X[j][c] for i > j, c <= j
#pragma unroll
        for (int aa = JB; aa < DT; ++aa)
#pragma unroll
            for (int bb = 0; bb <= JB; ++bb) {
                bool on = true;
                if (aa == JB) on = on && (ty > jx);
                if (bb == JB) on = on && (tx <= jx);
                if (on) bm[aa][bb] = fma(-lj[aa], xj[bb], bm[aa][bb]);
            }
    }
}

__global__ void __launch_bounds__(DG * DG, 1) diag_block_kernel(double* a, long lda, long m, long first_block,
                                                               double* dinv_base, int mode, int* info) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s = reinterpret_cast<double*>(smem_raw);      // [NB][DP1]: L (lower incl. diagonal) once known
    double* sr = s + NB * DP1;                            // [NB] reciprocal diagonal of L
    double* buf = sr + NB;                                // [2][NB] published column (Cholesky) / row (inverse) of a step
    __shared__ int s_fail;
    const long blk = first_block + blockIdx.x;
    const long r0 = blk * NB;
    const int nb = (int)((m - r0) < NB ? (m - r0) : NB);
    double* ablk = a + r0 * lda + r0;
    double* dinv = dinv_base + blk * (long)NB * NB;
    const int tid = threadIdx.x;
    const int ty = tid / DG, tx = tid % DG;
    if (tid == 0) s_fail = 0;

    if (mode == 0) {
        // working matrix in registers: lower triangle of the block, identity padding beyond nb
        double r[DT][DT];
#pragma unroll
        for (int aa = 0; aa < DT; ++aa)
#pragma unroll
            for (int bb = 0; bb < DT; ++bb) {
                const int i = ty + DG * aa, k = tx + DG * bb;
                double v = 0.0;
                if (i < nb && k <= i) v = ablk[(long)i * lda + k];
                if (i >= nb && i == k) v = 1.0;
                r[aa][bb] = v;
            }
        __syncthreads();
        chol_group<0>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<1>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<2>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<3>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<4>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<5>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<6>(r, s, sr, buf, ty, tx, &s_fail);
        chol_group<7>(r, s, sr, buf, ty, tx, &s_fail);
        __syncthreads();
        // write L back, zero the upper triangle of the block
        for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
            const int i = idx / nb, j = idx % nb;
            ablk[(long)i * lda + j] = (j <= i) ? s[i * DP1 + j] : 0.0;
        }
        if (tid == 0 && s_fail) atomicCAS(info, 0, (int)(blk + 1));
    } else {
        for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
            const int i = idx / NB, j = idx % NB;
            double v = 0.0;
            if (i < nb && j <= i) v = ablk[(long)i * lda + j];
            if (i >= nb && i == j) v = 1.0;
            s[i * DP1 + j] = v;
            if (i == j) sr[i] = 1.0 / v;
        }
        __syncthreads();
    }
    // inverse X = L^-1 by forward substitution on the identity, B in registers (same ownership)
    double bm[DT][DT];
#pragma unroll
    for (int aa = 0; aa < DT; ++aa)
#pragma unroll
        for (int bb = 0; bb < DT; ++bb) bm[aa][bb] = (ty + DG * aa == tx + DG * bb) ? 1.0 : 0.0;
    inv_group<0>(bm, s, sr, buf, ty, tx);
    inv_group<1>(bm, s, sr, buf, ty, tx);
    inv_group<2>(bm, s, sr, buf, ty, tx);
    inv_group<3>(bm, s, sr, buf, ty, tx);
    inv_group<4>(bm, s, sr, buf, ty, tx);
    inv_group<5>(bm, s, sr, buf, ty, tx);
    inv_group<6>(bm, s, sr, buf, ty, tx);
    inv_group<7>(bm, s, sr, buf, ty, tx);
#pragma unroll
    for (int aa = 0; aa < DT; ++aa)
#pragma unroll
        for (int bb = 0; bb < DT; ++bb) {
            const int i = ty + DG * aa, c = tx + DG * bb;
            dinv[i * NB + c] = (c <= i) ? bm[aa][bb] : 0.0;
        }
}

static size_t diag_smem_bytes() { return (size_t)(NB * DP1 + 3 * NB) * sizeof(double); }

static int launch_diag(Context* ctx, double* a, long lda, long m, long first_block, long nblocks, double* dinv, int mode,
                       int* info, cudaStream_t st) {
    size_t smem = diag_smem_bytes();
    CGLB_CUDA_OK(cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    diag_block_kernel<<<(unsigned)nblocks, DG * DG, smem, st>>>(a, lda, m, first_block, dinv, mode, info);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

__global__ void zero_upper_kernel(double* a, long m, long lda) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    long i = idx / m, j = idx % m;
    if (j > i) a[i * lda + j] = 0.0;
}

__global__ void set_zero_kernel(double* a, long rows, long cols, long ld) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    a[(idx / cols) * ld + idx % cols] = 0.0;
}

// copies diag inverse blocks (scaled) into the block diagonal of dst
__global__ void place_diag_blocks_kernel(const double* dinv, double* dst, long m, long ldd, double scale) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long nblk = (m + NB - 1) / NB;
    if (idx >= nblk * NB * NB) return;
    long blk = idx / (NB * NB);
    int i = (int)((idx / NB) % NB), j = (int)(idx % NB);
    long r = blk * NB + i, c = blk * NB + j;
    if (r < m && c < m) dst[r * ldd + c] = scale * dinv[idx];
}

// workspace layout inside ctx->scratch for the blocked algorithms
struct DenseWs {
    double* dinv;   // [nblk][NB][NB]
    double* lhat;   // [m_pad][m_pad]
    long m_pad;
};
static int get_dense_ws(Context* ctx, long m, DenseWs& ws) {
    long nblk = (m + NB - 1) / NB;
    ws.m_pad = nblk * NB;
    long need = nblk * NB * NB + ws.m_pad * ws.m_pad + kScratchScalars;
    int rc = ensure_scratch(ctx, need + 64);
    if (rc) return rc;
    // first 64 doubles of scratch are reserved for the sweeps' scalar accumulators
    ws.dinv = ctx->scratch + kScratchScalars;
    ws.lhat = ws.dinv + nblk * NB * NB;
    return CGLB_OK;
}

// lhat[k, 0:k) = -dinv_k * l[k, 0:k)   for every block row k   (one small GEMM per block row)
static int build_lhat(Context* ctx, const double* l, long m, long ldl, const DenseWs& ws, cudaStream_t st) {
    long nblk = (m + NB - 1) / NB;
    for (long k = 1; k < nblk; ++k) {
        long rows = (m - k * NB) < NB ? (m - k * NB) : NB;
        // A-operand: dinv_k (NB x NB, lda NB) restricted to `rows` rows; B-operand: l[k*NB.., 0:k*NB) (K = rows)
        GemmArgs p{ws.dinv + k * NB * NB, NB, l + k * NB * ldl, ldl, ws.lhat + k * NB * ws.m_pad, ws.m_pad,
                   rows, k * NB, rows, NB, -1.0, 0.0, 0, {}};
        int rc = launch_gemm<false, EPI_STORE>(ctx, p, 1, st);      // K = 128: nothing to split
        if (rc) return rc;
    }
    return CGLB_OK;
}


// ---------------------------------------------------------------------------------------------
// wide-input (d > 32) K_nm build and backward through the DMMA GEMM + kernel-map epilogues
// ---------------------------------------------------------------------------------------------
__global__ void knm_wide_assemble_kernel(const double* __restrict__ zp, long m, const double* __restrict__ xp, long ncols, int d, int w, int kp,
                                         const double* __restrict__ rsum, const double* __restrict__ csum, const double* __restrict__ gx,
                                         const double* __restrict__ gk_sum, const double* __restrict__ ls, double cscale,
                                         double* __restrict__ out_ls, double* __restrict__ out_var, double* __restrict__ out_z) {
    // one block per input dimension q
    const int q = blockIdx.x;
    __shared__ double sh[256];
    double s = 0.0;
    for (long mm = threadIdx.x; mm < m; mm += blockDim.x) {
        const double z = zp[mm * w + q], r = rsum[mm], gxv = gx[mm * kp + q];
        s += z * z * r - 2.0 * z * gxv;
        if (out_z) atomicAdd(out_z + mm * d + q, -(cscale / ls[q]) * (z * r - gxv));
    }
    for (long i = threadIdx.x; i < ncols; i += blockDim.x) {
        const double x = xp[i * w + q];
        s = fma(x * x, csum[i], s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(out_ls + q, sh[0] / ls[q]);
        if (q == 0) atomicAdd(out_var, *gk_sum);
    }
}

// second stage of the EPI_KBWD sums: slots added in tile order (fixed summation order)
__global__ void kbwd_reduce_kernel(const double* __restrict__ rpart, const double* __restrict__ cpart, const double* __restrict__ gpart,
                                   long m, long n, long tiles_m, long tiles_n, double* __restrict__ rsum, double* __restrict__ csum,
                                   double* __restrict__ gk) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        double s = 0.0;
        for (long b = 0; b < tiles_n; ++b) s += rpart[b * m + i];
        rsum[i] = s;
    } else if (i < m + n) {
        const long c = i - m;
        double s = 0.0;
        for (long b = 0; b < tiles_m; ++b) s += cpart[b * n + c];
        csum[c] = s;
    }
    if (i == 0) {
        double s = 0.0;
        for (long b = 0; b < tiles_m * tiles_n; ++b) s += gpart[b];
        *gk = s;
    }
}

int knm_build_wide(Context* ctx, int kind, const double* zp, long m, const double* xp, long n, int d, double variance,
                   double* out, long ld, cudaStream_t st) {
    const int kp = wide_kp(d), w = packed_width(d);
    KEpiArgs ke{};
    ke.nz = zp + kp; ke.nz_stride = w; ke.nx = xp + kp; ke.nx_stride = w; ke.exp_tab = ctx->exp_table;
    ke.variance = variance; ke.kind = kind;
    GemmArgs p{zp, w, xp, w, out, ld, m, n, kp, (kp + GK - 1) / GK * GK, 1.0, 0.0, 0, ke};
    return launch_gemm<true, EPI_KMAP>(ctx, p, 1, st);
}

int knm_backward_wide(Context* ctx, int kind, const double* zp, long m, const double* xp, long ncols, int d, double variance,
                      const double* lengthscale, double* t, long ldt, const double* wt, const double* zvec, double* out_ls,
                      double* out_var, double* out_z, cudaStream_t st) {
    if (!t) {
        set_error("knm_backward: for d > %d the dense operand t is required (it is used as workspace and overwritten)", CGLB_MAX_REGISTER_D);
        return CGLB_ERR_UNSUPPORTED;
    }
    const int kp = wide_kp(d), w = packed_width(d);
    // workspace: [rsum m | csum ncols | gk 1 | gx m*kp] in scratch; the epilogue's per-CTA partial slots in ypart
    const long need = m + ncols + 1 + m * kp;
    int rc = ensure_scratch(ctx, kScratchScalars + need);
    if (rc) return rc;
    double* rsum = ctx->scratch + kScratchScalars;
    double* csum = rsum + m;
    double* gk = csum + ncols;
    double* gx = gk + 1;
    const long tiles_m = (m + GM - 1) / GM, tiles_n = (ncols + GN - 1) / GN;
    rc = ensure_ypart(ctx, tiles_n * m + tiles_m * ncols + tiles_m * tiles_n);
    if (rc) return rc;
    double* rpart = ctx->ypart;
    double* cpart = rpart + tiles_n * m;
    double* gpart = cpart + tiles_m * ncols;
    KEpiArgs ke{};
    ke.nz = zp + kp; ke.nz_stride = w; ke.nx = xp + kp; ke.nx_stride = w; ke.exp_tab = ctx->exp_table;
    ke.wt = wt; ke.zvec = zvec; ke.rsum = rpart; ke.csum = cpart; ke.gk_sum = gpart;
    ke.variance = variance; ke.vc = variance * ((kind == CGLB_MATERN32) ? 1.0 : 2.0); ke.kind = kind;
    // S = Zp Xp^T, epilogue writes GP over t (beta = 1 flags "t holds a dense G part")
    GemmArgs p{zp, w, xp, w, t, ldt, m, ncols, kp, (kp + GK - 1) / GK * GK, 1.0, 1.0, 0, ke};
    rc = launch_gemm<true, EPI_KBWD>(ctx, p, 1, st);
    if (rc) return rc;
    kbwd_reduce_kernel<<<(unsigned)((m + ncols + 255) / 256), 256, 0, st>>>(rpart, cpart, gpart, m, ncols, tiles_m, tiles_n, rsum, csum, gk);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    // GX = GP * Xp (coordinates): split-K over the columns
    long tiles = (m + GM - 1) / GM;
    long want = (2L * ctx->num_sms + tiles - 1) / tiles;
    long max_split = (ncols + 255) / 256;
    long ksplit = want < max_split ? want : max_split;
    if (ksplit < 1) ksplit = 1;
    long chunk = ((ncols + ksplit - 1) / ksplit + GK - 1) / GK * GK;
    ksplit = (ncols + chunk - 1) / chunk;
    GemmArgs pg{t, ldt, xp, w, gx, kp, m, kp, ncols, chunk, 1.0, 0.0, 0, {}};
    rc = launch_gemm_splitk<false, EPI_ATOMIC>(ctx, pg, (int)ksplit, 0, st);
    if (rc) return rc;
    const double cscale = (kind == CGLB_MATERN32) ? 1.7320508075688772935 : 0.70710678118654752440;
    knm_wide_assemble_kernel<<<d, 256, 0, st>>>(zp, m, xp, ncols, d, w, kp, rsum, csum, gx, gk, lengthscale, cscale, out_ls, out_var, out_z);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

}  // namespace cglb

using namespace cglb;

extern "C" int cglb_gemm(cglb_context* c, int transb, long m, long n, long k, double alpha, const double* a, long lda,
                         const double* b, long ldb, double beta, double* cc, long ldc, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && a && b && cc, "null pointer");
    CGLB_CHECK_ARG(m >= 0 && n >= 0 && k >= 0, "negative dimension");
    return gemm_checked(ctx, transb, m, n, k, alpha, a, lda, b, ldb, beta, cc, ldc, (cudaStream_t)stream);
}

extern "C" int cglb_syrk(cglb_context* c, const double* a, long m, long n, long lda, double* cm, long ldc, int accumulate,
                         void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && a && cm, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return CGLB_OK;
    if (n == 0) {
        if (!accumulate) {
            set_zero_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, st>>>(cm, m, m, ldc);
            ctx->launches++;
            CGLB_LAUNCH_OK();
        }
        return CGLB_OK;
    }
    long tiles_m = (m + GM - 1) / GM;
    long lower_tiles = tiles_m * (tiles_m + 1) / 2;
    // enough CTAs for ~4 waves, but at least 512 columns of K per chunk
    long want = (4L * ctx->num_sms + lower_tiles - 1) / lower_tiles;
    long max_split = (n + 511) / 512;
    long ksplit = want < max_split ? want : max_split;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 65535) ksplit = 65535;
    long chunk = ((n + ksplit - 1) / ksplit + GK - 1) / GK * GK;
    ksplit = (n + chunk - 1) / chunk;
    GemmArgs p{a, lda, a, lda, cm, ldc, m, m, n, chunk, 1.0, accumulate ? 1.0 : 0.0, 1, {}};
    return launch_gemm_splitk<true, EPI_SYRK>(ctx, p, (int)ksplit, 1, st);
}

extern "C" int cglb_potrf(cglb_context* c, double* a, long m, long lda, int* info_dev, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && a && info_dev, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return CGLB_OK;
    DenseWs ws;
    int rc = get_dense_ws(ctx, m, ws);
    if (rc) return rc;
    CGLB_CUDA_OK(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
    long nblk = (m + NB - 1) / NB;
    for (long k = 0; k < nblk; ++k) {
        rc = launch_diag(ctx, a, lda, m, k, 1, ws.dinv, 0, info_dev, st);
        if (rc) return rc;
        long r1 = (k + 1) * NB;
        if (r1 >= m) break;
        long rows = m - r1;
        // panel: a[r1:, kNB:r1) <- a[r1:, kNB:r1) * dinv_k^T      (NT, in place: a CTA owns its rows)
        GemmArgs pp{a + r1 * lda + k * NB, lda, ws.dinv + k * NB * NB, NB, a + r1 * lda + k * NB, lda,
                    rows, NB, NB, NB, 1.0, 0.0, 0, {}};
        rc = launch_gemm<true, EPI_STORE>(ctx, pp, 1, st);
        if (rc) return rc;
        // trailing: a[r1:, r1:) -= P P^T  (lower tiles only)
        GemmArgs pt{a + r1 * lda + k * NB, lda, a + r1 * lda + k * NB, lda, a + r1 * lda + r1, lda,
                    rows, rows, NB, NB, -1.0, 1.0, 1, {}};
        rc = launch_gemm<true, EPI_STORE>(ctx, pt, 1, st);
        if (rc) return rc;
    }
    zero_upper_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, st>>>(a, m, lda);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

extern "C" int cglb_tri_inverse(cglb_context* c, const double* l, long m, long ldl, double* linv, long ldi, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && l && linv, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return CGLB_OK;
    DenseWs ws;
    int rc = get_dense_ws(ctx, m, ws);
    if (rc) return rc;
    long nblk = (m + NB - 1) / NB;
    rc = launch_diag(ctx, const_cast<double*>(l), ldl, m, 0, nblk, ws.dinv, 1, nullptr, st);
    if (rc) return rc;
    rc = build_lhat(ctx, l, m, ldl, ws, st);
    if (rc) return rc;
    set_zero_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, st>>>(linv, m, m, ldi);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    place_diag_blocks_kernel<<<(unsigned)((nblk * NB * NB + 255) / 256), 256, 0, st>>>(ws.dinv, linv, m, ldi, 1.0);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    // linv[k, 0:k) = lhat[k, 0:k) * linv[0:k, 0:k)
    for (long k = 1; k < nblk; ++k) {
        long rows = (m - k * NB) < NB ? (m - k * NB) : NB;
        GemmArgs p{ws.lhat + k * NB * ws.m_pad, ws.m_pad, linv, ldi, linv + k * NB * ldi, ldi,
                   rows, k * NB, k * NB, k * NB, 1.0, 0.0, 0, {}};
        rc = launch_gemm_auto<false>(ctx, p, st);                   // skinny output, long K: split over the SMs
        if (rc) return rc;
    }
    return CGLB_OK;
}

extern "C" int cglb_trsm_left_lower(cglb_context* c, const double* l, long m, long ldl, double* b, long n, long ldb,
                                    double alpha, void* stream) {
    Context* ctx = reinterpret_cast<Context*>(c);
    CGLB_CHECK_ARG(ctx && l && b, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0 || n == 0) return CGLB_OK;
    DenseWs ws;
    int rc = get_dense_ws(ctx, m, ws);
    if (rc) return rc;
    long nblk = (m + NB - 1) / NB;
    rc = launch_diag(ctx, const_cast<double*>(l), ldl, m, 0, nblk, ws.dinv, 1, nullptr, st);
    if (rc) return rc;
    rc = build_lhat(ctx, l, m, ldl, ws, st);
    if (rc) return rc;
    // block diagonal of lhat <- alpha * dinv_k :   X_k = [lhat_k | alpha dinv_k] [X_0..X_{k-1}; B_k]
    place_diag_blocks_kernel<<<(unsigned)((nblk * NB * NB + 255) / 256), 256, 0, st>>>(ws.dinv, ws.lhat, ws.m_pad, ws.m_pad, alpha);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    for (long k = 0; k < nblk; ++k) {
        long rows = (m - k * NB) < NB ? (m - k * NB) : NB;
        long kk = k * NB + rows;
        GemmArgs p{ws.lhat + k * NB * ws.m_pad, ws.m_pad, b, ldb, b + k * NB * ldb, ldb,
                   rows, n, kk, (kk + GK - 1) / GK * GK, 1.0, 0.0, 0, {}};
        rc = launch_gemm<false, EPI_STORE>(ctx, p, 1, st);
        if (rc) return rc;
    }
    return CGLB_OK;
}
