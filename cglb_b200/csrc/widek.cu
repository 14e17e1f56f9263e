// Wide-input (d > 32, e.g. the song-shaped d = 90 config) matrix-free sweeps: the distance contraction
// a_i . b_j runs on DMMA.8x8x4 with both operand tiles in shared memory; the kernel map, the products with
// v and the row/column reductions are applied to the accumulator fragments in registers.
//
//   CTA = 8 warps as 4 (rows) x 2 (cols), warp tile 32 x 32 pairs = 4 x 4 DMMA tiles
//   work item = (row block of 128 rows, chunk of 1024 columns) ; the row tile stays resident, 64-column
//   tiles stream through a 2-stage TMA (cp.async.bulk) ring; symmetric sweep: tiles on/above the diagonal
//   only, off-diagonal tiles feed y_i and y_j.
// Algorithmic FLOPs per pair (expanded form): 2d + 10 (Matern32).
#include "kmv_impl.cuh"

namespace cglb {

constexpr int WT_ROWS = 128;         // rows per item
constexpr int WT_COLS = 64;          // columns per streamed tile
constexpr int WT_CHUNK = 1024;       // columns per item
constexpr int WT_THREADS = 256;

struct WideArgs {
    const double* xp_rows; const double* xp_cols;   // wide packed [.][W]
    const double* vcol;                              // padded column vector
    double* y;                                       // CTA b accumulates into y + b * ystride (fixed order, kmv_impl.cuh)
    long ystride;
    const double* exp_tab;
    long nrows, ncols;
    long nb_rows;            // row blocks of 128
    long n_chunks;           // column chunks of 1024
    long nitems;
    double variance;
    int d, kp, w;
    int part, nparts;
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// item t -> (row block I, column chunk c).  SYM: chunk-major enumeration of {(I, c) : I < 8 (c + 1)} (the
// row block must start at or before the end of the chunk); entries with I >= nb_rows are skipped.
__device__ __forceinline__ bool wide_item(long t, const WideArgs& a, bool sym, long& I, long& c) {
    if (!sym) { I = t / a.n_chunks; c = t % a.n_chunks; return true; }
    long cc = (long)((sqrt(1.0 + (double)t) - 1.0) * 0.5);
    while (4 * cc * (cc + 1) > t) --cc;
    while (4 * (cc + 1) * (cc + 2) <= t) ++cc;
    c = cc;
    I = t - 4 * cc * (cc + 1);
    return I < a.nb_rows;
}

template <int KIND, bool SYM>
__global__ void __launch_bounds__(WT_THREADS, 1) wide_sweep_kernel(const WideArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = args.w, KP = args.kp;
    double* s_rows = reinterpret_cast<double*>(smem_raw);                 // [128][W]
    double* s_cols = s_rows + WT_ROWS * W;                                // [2][64][W]
    double* s_v = s_cols + 2 * WT_COLS * W;                               // [2][64]
    double* s_col = s_v + 2 * WT_COLS;                                    // [2][4][64] per row-warp column sums, by tile parity
    double* s_row = s_col + 8 * WT_COLS;                                  // [2][128]  per col-warp row sums
    double* s_tab = s_row + 2 * WT_ROWS;                                  // [64] (DMMA-dominated: the small exp table keeps d <= 104 in 227 KB)
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_tab + kExpTabSmall); // [2] column tiles
    uint64_t* s_empty = s_full + 2;                                       // [2]
    uint64_t* s_rowbar = s_empty + 2;                                     // [1] row tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    if (tid < kExpTabSmall) s_tab[tid] = args.exp_tab[tid];
    if (tid == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], 8); mbar_init(&s_empty[1], 8);
        mbar_init(s_rowbar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    int stage = 0; uint32_t phase = 0;          // consumer side of the column ring
    int pstage = 0; uint32_t pphase = 0;        // producer side
    uint32_t rowphase = 0;
    const double var = args.variance;
    const uint32_t tile_bytes = (uint32_t)(WT_COLS * W * sizeof(double));
    // this CTA's copy of y: every add below follows a CTA barrier and column j / row i always belong to the same thread
    double* const yb = args.y + (long)blockIdx.x * args.ystride;

    for (long tau = blockIdx.x;; tau += gridDim.x) {
        const long t = tau * args.nparts + args.part;
        if (t >= args.nitems) break;
        long I, C;
        if (!wide_item(t, args, SYM, I, C)) continue;
        const long r0 = I * WT_ROWS;
        const long cbeg0 = C * WT_CHUNK;
        long cend = cbeg0 + WT_CHUNK;
        if (cend > args.ncols) cend = args.ncols;
        // SYM: first tile at or after the start of the row block (tiles are 64 wide, the row block 128)
        long cbeg = cbeg0;
        if (SYM && cbeg < r0) cbeg = r0;
        if (cbeg >= cend) continue;
        const int ntiles = (int)((cend - cbeg + WT_COLS - 1) / WT_COLS);

        // all warps are done with the previous item's row tile before it is overwritten
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_rowbar, (uint32_t)(WT_ROWS * W * sizeof(double)));
            tma_load_1d(s_rows, args.xp_rows + r0 * W, (uint32_t)(WT_ROWS * W * sizeof(double)), s_rowbar);
            // first column tile
            mbar_wait(&s_empty[pstage], pphase ^ 1);
            mbar_expect_tx(&s_full[pstage], tile_bytes + WT_COLS * sizeof(double));
            tma_load_1d(s_cols + pstage * WT_COLS * W, args.xp_cols + cbeg * W, tile_bytes, &s_full[pstage]);
            tma_load_1d(s_v + pstage * WT_COLS, args.vcol + cbeg, WT_COLS * sizeof(double), &s_full[pstage]);
            if (++pstage == 2) { pstage = 0; pphase ^= 1; }
        }
        mbar_wait(s_rowbar, rowphase);
        rowphase ^= 1;

        double na[4], vrow[4], racc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = wm * 32 + i * 8 + g;
            na[i] = s_rows[r * W + KP];
            const long row = r0 + r;
            vrow[i] = (SYM && row < args.nrows) ? __ldg(args.vcol + row) : 0.0;
            racc[i] = 0.0;
        }

        for (int tile = 0; tile < ntiles; ++tile) {
            const long j0 = cbeg + (long)tile * WT_COLS;
            if (tid == 0 && tile + 1 < ntiles) {                 // prefetch the next column tile
                mbar_wait(&s_empty[pstage], pphase ^ 1);
                mbar_expect_tx(&s_full[pstage], tile_bytes + WT_COLS * sizeof(double));
                tma_load_1d(s_cols + pstage * WT_COLS * W, args.xp_cols + (j0 + WT_COLS) * W, tile_bytes, &s_full[pstage]);
                tma_load_1d(s_v + pstage * WT_COLS, args.vcol + j0 + WT_COLS, WT_COLS * sizeof(double), &s_full[pstage]);
                if (++pstage == 2) { pstage = 0; pphase ^= 1; }
            }
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const double* sc = s_cols + stage * WT_COLS * W;
            const double* sv = s_v + stage * WT_COLS;
            // SYM: a tile that overlaps the diagonal block of this row block contributes to rows only
            const bool offdiag = SYM && (j0 >= r0 + WT_ROWS);

            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            const double* ap = s_rows + (wm * 32 + g) * W + t4;
            const double* bp = sc + (wn * 32 + g) * W + t4;
#pragma unroll 2
            for (int k4 = 0; k4 < KP; k4 += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = ap[i * 8 * W + k4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = bp[j * 8 * W + k4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j], af[i], bf[j]);
            }
            // epilogue on the fragments: lane holds (row g + 8i, cols 8j + 2 t4 + e)
            double cacc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double nb[2], vc[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = wn * 32 + j * 8 + 2 * t4 + e;
                    nb[e] = sc[c * W + KP];
                    vc[e] = sv[c];
                    cacc[j][e] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double q = fma(-2.0, acc[i][j][e], na[i] + nb[e]);
                        const double kk = kappa<KIND>(q, s_tab);
                        racc[i] = fma(kk, vc[e], racc[i]);
                        if (SYM) cacc[j][e] = fma(kk, vrow[i], cacc[j][e]);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == 2) { stage = 0; phase ^= 1; }

            if (offdiag) {
                // reduce the 8 column partials over the 8 lanes that share t4 (bits 2..4 of the lane id):
                // transposing butterfly, 7 adds instead of 24
                double c8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) { c8[2 * j] = cacc[j][0]; c8[2 * j + 1] = cacc[j][1]; }
                int cnt = 8;
#pragma unroll
                for (int off = 16; off >= 4; off >>= 1, cnt >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int h = 0; h < cnt / 2; ++h) {
                        const double send = up ? c8[h] : c8[h + cnt / 2];
                        const double keep = up ? c8[h + cnt / 2] : c8[h];
                        c8[h] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                // lane now holds entry idx = (bit4, bit3, bit2) of c8 order: idx = 2 j + e
                const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                const int col = wn * 32 + (idx >> 1) * 8 + 2 * t4 + (idx & 1);
                s_col[((tile & 1) * 4 + wm) * WT_COLS + col] = c8[0];
            }
            if (SYM) {
                __syncthreads();          // also orders s_col reuse between tiles
                if (offdiag && tid < WT_COLS) {
                    const long j = j0 + tid;
                    const double* sc4 = s_col + (tile & 1) * 4 * WT_COLS;
                    const double s = sc4[tid] + sc4[WT_COLS + tid] + sc4[2 * WT_COLS + tid] + sc4[3 * WT_COLS + tid];
                    if (j < args.ncols) atomicAdd(yb + j, var * s);
                }
            }
        }
        // row sums: reduce over the 4 lanes sharing g, then over the 2 column warps
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = racc[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (t4 == 0) s_row[wn * WT_ROWS + wm * 32 + i * 8 + g] = s;
        }
        __syncthreads();
        if (tid < WT_ROWS) {
            const long row = r0 + tid;
            if (row < args.nrows) atomicAdd(yb + row, var * (s_row[tid] + s_row[WT_ROWS + tid]));
        }
    }
}

static size_t wide_smem_bytes(int w) {
    return (size_t)(WT_ROWS * w + 2 * WT_COLS * w + 2 * WT_COLS + 8 * WT_COLS + 2 * WT_ROWS + kExpTabSmall) * sizeof(double) + 8 * sizeof(uint64_t);
}

int wide_sweep(Context* ctx, int kind, bool sym, const double* xp_rows, long nrows, const double* xp_cols, long ncols, int d,
               const double* vcol, double* y, long ystride, double variance, int part, int nparts, cudaStream_t st) {
    WideArgs a{};
    a.xp_rows = xp_rows; a.xp_cols = xp_cols; a.vcol = vcol; a.y = y; a.ystride = ystride; a.exp_tab = ctx->exp_table;
    a.nrows = nrows; a.ncols = ncols; a.variance = variance; a.d = d; a.kp = wide_kp(d); a.w = packed_width(d);
    a.part = part; a.nparts = nparts;
    a.nb_rows = (nrows + WT_ROWS - 1) / WT_ROWS;
    a.n_chunks = (ncols + WT_CHUNK - 1) / WT_CHUNK;
    a.nitems = sym ? 4 * a.n_chunks * (a.n_chunks + 1) : a.nb_rows * a.n_chunks;
    size_t smem = wide_smem_bytes(a.w);
    if (smem > 227 * 1024) {
        set_error("wide sweep: d=%d needs %zu bytes of shared memory (> 227 KB); d <= 104 is supported", d, smem);
        return CGLB_ERR_UNSUPPORTED;
    }
    long my_items = (a.nitems - part + nparts - 1) / nparts;
    if (my_items <= 0) return CGLB_OK;
    int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
#define LAUNCH(K, S)                                                                                             \
    do {                                                                                                         \
        CGLB_CUDA_OK(cudaFuncSetAttribute(wide_sweep_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        wide_sweep_kernel<K, S><<<grid, WT_THREADS, smem, st>>>(a);                                               \
    } while (0)
    if (kind == CGLB_MATERN32) { if (sym) LAUNCH(CGLB_MATERN32, true); else LAUNCH(CGLB_MATERN32, false); }
    else { if (sym) LAUNCH(CGLB_RBF, true); else LAUNCH(CGLB_RBF, false); }
#undef LAUNCH
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}


// ---------------------------------------------------------------------------------------------
// wide backward sweep (K2 for d > 32): per 64 x 64 tile
//   phase A  S = A_I A_J^T on DMMA -> (kappa, e') ; omega = u_i w_j + w_i u_j ; c = e' omega
//            gvar += kappa omega ; row/column sums of c (R) ; c -> shared memory
//   phase B  Y_I += C (64x64) * A_J (64 x KP) on DMMA         (cross term X_q = sum_i a_iq Y_iq at item end)
// item = (row block of 64 rows, chunk of 1024 columns); tiles on/above the diagonal; the diagonal tile is
// evaluated as a full square with halved weights and doubled row sums (as in the d <= 32 kernel).
// ---------------------------------------------------------------------------------------------
constexpr int WB_ROWS = 64;
constexpr int WB_CP = 68;            // pitch of the c tile

struct WideBwdArgs {
    const double* xp;                // wide packed [n_pad][W]
    const double* wcol; const double* ucol;   // padded vectors
    double* rsum;                    // R: CTA b accumulates into rsum + b * ystride
    double* gout;                    // [d+1]: -2 X_q, variance sum of CTA b at gout + b * gstride
    long ystride, gstride;
    const double* exp_tab;
    long n;
    long nb_rows, n_chunks, nitems;
    int d, kp, w;
    int part, nparts;
};

__device__ __forceinline__ bool wide_bwd_item(long t, const WideBwdArgs& a, long& I, long& c) {
    // chunk-major enumeration of {(I, c) : I < 16 (c + 1)}: prefix 8 c (c + 1)
    long cc = (long)((sqrt(1.0 + 0.5 * (double)t) - 1.0) * 0.5);
    while (8 * cc * (cc + 1) > t) --cc;
    while (8 * (cc + 1) * (cc + 2) <= t) ++cc;
    c = cc;
    I = t - 8 * cc * (cc + 1);
    return I < a.nb_rows;
}

template <int KIND, int NQ>
__global__ void __launch_bounds__(WT_THREADS, 1) wide_bwd_kernel(const WideBwdArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = args.w, KP = args.kp;
    double* s_rows = reinterpret_cast<double*>(smem_raw);                 // [64][W]
    double* s_cols = s_rows + WB_ROWS * W;                                // [2][64][W]
    double* s_c = s_cols + 2 * WT_COLS * W;                               // [64][68]
    double* s_wu = s_c + WB_ROWS * WB_CP;                                 // [2][2][64]  (w, u) per stage
    double* s_col = s_wu + 4 * WT_COLS;                                   // [2][2][64]  column sums (tile parity, row warp)
    double* s_row = s_col + 4 * WT_COLS;                                  // [4][64]     row sums per column warp
    double* s_xq = s_row + 4 * WB_ROWS;                                   // [8][NQ*8]
    double* s_tab = s_xq + 8 * NQ * 8;                                    // [64]
    double* s_red = s_tab + kExpTabSmall;                                 // [8]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_red + 8);
    uint64_t* s_empty = s_full + 2;
    uint64_t* s_rowbar = s_empty + 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    if (tid < kExpTabSmall) s_tab[tid] = args.exp_tab[tid];
    if (tid == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], 8); mbar_init(&s_empty[1], 8);
        mbar_init(s_rowbar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    int stage = 0; uint32_t phase = 0;
    int pstage = 0; uint32_t pphase = 0;
    uint32_t rowphase = 0;
    const uint32_t tile_bytes = (uint32_t)(WT_COLS * W * sizeof(double));
    double gvar = 0.0;
    double* const rb = args.rsum + (long)blockIdx.x * args.ystride;
    double* const gb = args.gout + (long)blockIdx.x * args.gstride;
    double gx = 0.0;                 // thread tid < d: running -2 X_q of this CTA (items in their static order)

    auto issue_tile = [&](long j0) {
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        mbar_expect_tx(&s_full[pstage], tile_bytes + 2 * WT_COLS * sizeof(double));
        tma_load_1d(s_cols + pstage * WT_COLS * W, args.xp + j0 * W, tile_bytes, &s_full[pstage]);
        tma_load_1d(s_wu + pstage * 2 * WT_COLS, args.wcol + j0, WT_COLS * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_wu + pstage * 2 * WT_COLS + WT_COLS, args.ucol + j0, WT_COLS * sizeof(double), &s_full[pstage]);
        if (++pstage == 2) { pstage = 0; pphase ^= 1; }
    };

    for (long tau = blockIdx.x;; tau += gridDim.x) {
        const long t = tau * args.nparts + args.part;
        if (t >= args.nitems) break;
        long I, C;
        if (!wide_bwd_item(t, args, I, C)) continue;
        const long r0 = I * WB_ROWS;
        long cend = C * WT_CHUNK + WT_CHUNK;
        if (cend > args.n) cend = args.n;
        long cbeg = C * WT_CHUNK;
        if (cbeg < r0) cbeg = r0;
        if (cbeg >= cend) continue;
        const int ntiles = (int)((cend - cbeg + WT_COLS - 1) / WT_COLS);

        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_rowbar, (uint32_t)(WB_ROWS * W * sizeof(double)));
            tma_load_1d(s_rows, args.xp + r0 * W, (uint32_t)(WB_ROWS * W * sizeof(double)), s_rowbar);
            issue_tile(cbeg);
        }
        mbar_wait(s_rowbar, rowphase);
        rowphase ^= 1;

        double na[4], ui[4], wi[4], racc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = wm * 32 + i * 8 + g;
            na[i] = s_rows[r * W + KP];
            const long row = r0 + r;
            const bool live = row < args.n;
            ui[i] = live ? __ldg(args.ucol + row) : 0.0;
            wi[i] = live ? __ldg(args.wcol + row) : 0.0;
            racc[i] = 0.0;
        }
        double yacc[NQ][2];
#pragma unroll
        for (int nq = 0; nq < NQ; ++nq) yacc[nq][0] = yacc[nq][1] = 0.0;

        for (int tile = 0; tile < ntiles; ++tile) {
            const long j0 = cbeg + (long)tile * WT_COLS;
            if (tid == 0 && tile + 1 < ntiles) issue_tile(j0 + WT_COLS);
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const double* sc = s_cols + stage * WT_COLS * W;
            const double* sw = s_wu + stage * 2 * WT_COLS;
            const double* su = sw + WT_COLS;
            const bool offdiag = (j0 >= r0 + WB_ROWS);
            const double half = offdiag ? 1.0 : 0.5;      // diagonal tile: ordered pairs with halved weights ...
            const double rmul = offdiag ? 1.0 : 2.0;      // ... whose row sums count double

            // ---- phase A
            double acc[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            {
                const double* ap = s_rows + (wm * 32 + g) * W + t4;
                const double* bp = sc + (wn * 16 + g) * W + t4;
#pragma unroll 2
                for (int k4 = 0; k4 < KP; k4 += 4) {
                    double af[4], bf[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[i] = ap[i * 8 * W + k4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) bf[j] = bp[j * 8 * W + k4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma884(acc[i][j], af[i], bf[j]);
                }
            }
            double cacc[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = wn * 16 + j * 8 + 2 * t4 + e;
                    const double nb = sc[c * W + KP], wj = sw[c], uj = su[c];
                    double cs = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double q = fma(-2.0, acc[i][j][e], na[i] + nb);
                        double kap, ew;
                        kappa_and_dweight<KIND>(q, s_tab, kap, ew);
                        const double om = half * fma(wi[i], uj, ui[i] * wj);
                        const double cw = ew * om;
                        gvar = fma(kap, om, gvar);
                        racc[i] = fma(rmul, cw, racc[i]);
                        cs += cw;
                        s_c[(wm * 32 + i * 8 + g) * WB_CP + c] = cw;
                    }
                    cacc[j][e] = cs;
                }
            }
            if (offdiag) {
                // 4 column partials over the 8 lanes sharing t4: 2 + 1 transposing steps, then one plain step
                double c4[4] = {cacc[0][0], cacc[0][1], cacc[1][0], cacc[1][1]};
                {
                    const bool up = (lane & 16) != 0;
                    double s0 = up ? c4[0] : c4[2], k0 = up ? c4[2] : c4[0];
                    double s1 = up ? c4[1] : c4[3], k1 = up ? c4[3] : c4[1];
                    c4[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                    c4[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                }
                {
                    const bool up = (lane & 8) != 0;
                    double s0 = up ? c4[0] : c4[1], k0 = up ? c4[1] : c4[0];
                    c4[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
                }
                c4[0] += __shfl_xor_sync(0xffffffffu, c4[0], 4);
                const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);          // = 2 j + e
                if ((lane & 4) == 0) s_col[((tile & 1) * 2 + wm) * WT_COLS + wn * 16 + (idx >> 1) * 8 + 2 * t4 + (idx & 1)] = c4[0];
            }
            __syncthreads();                       // c tile complete (and column sums visible)
            if (offdiag && tid < WT_COLS) {
                const long j = j0 + tid;
                const double* sc2 = s_col + (tile & 1) * 2 * WT_COLS;
                if (j < args.n) atomicAdd(rb + j, sc2[tid] + sc2[WT_COLS + tid]);
            }
            // ---- phase B: Y[rows of this warp (8)] += C[8 x 64] * A_J[64 x KP]
            {
                const double* cp = s_c + (warp * 8 + g) * WB_CP + t4;
                const double* bq = sc + t4 * W + g;
#pragma unroll 2
                for (int k4 = 0; k4 < WT_COLS; k4 += 4) {
                    const double af = cp[k4];
#pragma unroll
                    for (int nq = 0; nq < NQ; ++nq) dmma884(yacc[nq], af, bq[k4 * W + nq * 8]);
                }
            }
            __syncthreads();                       // all warps done with s_c and with this column stage
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == 2) { stage = 0; phase ^= 1; }
        }
        // ---- item end: row sums and the cross term
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = racc[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (t4 == 0) s_row[wn * WB_ROWS + wm * 32 + i * 8 + g] = s;
        }
        // X_q partial of this warp's 8 rows: lane holds Y[row w*8+g][nq*8 + 2 t4 + e]
#pragma unroll
        for (int nq = 0; nq < NQ; ++nq)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = nq * 8 + 2 * t4 + e;
                double s = yacc[nq][e] * s_rows[(warp * 8 + g) * W + q];
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (g == 0) s_xq[warp * NQ * 8 + q] = s;
            }
        __syncthreads();
        if (tid < WB_ROWS) {
            const long row = r0 + tid;
            if (row < args.n) atomicAdd(rb + row, s_row[tid] + s_row[WB_ROWS + tid] + s_row[2 * WB_ROWS + tid] + s_row[3 * WB_ROWS + tid]);
        }
        if (tid < args.d) {
            double s = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) s += s_xq[w8 * NQ * 8 + tid];
            gx = fma(-2.0, s, gx);
        }
    }
    if (tid < args.d) gb[tid] = gx;
    gvar = warp_sum(gvar);
    if (lane == 0) s_red[warp] = gvar;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) s += s_red[w8];
        gb[args.d] = s;
    }
}

template <int NQ>
static size_t wide_bwd_smem_bytes(int w) {
    return (size_t)(WB_ROWS * w + 2 * WT_COLS * w + WB_ROWS * WB_CP + 4 * WT_COLS + 4 * WT_COLS + 4 * WB_ROWS + 8 * NQ * 8 + kExpTabSmall + 8) * sizeof(double) +
           8 * sizeof(uint64_t);
}

template <int KIND, int NQ>
static int launch_wide_bwd(Context* ctx, const WideBwdArgs& a, int grid, cudaStream_t st) {
    size_t smem = wide_bwd_smem_bytes<NQ>(a.w);
    if (smem > 227 * 1024) {
        set_error("wide backward sweep: d=%d needs %zu bytes of shared memory", a.d, smem);
        return CGLB_ERR_UNSUPPORTED;
    }
    CGLB_CUDA_OK(cudaFuncSetAttribute(wide_bwd_kernel<KIND, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wide_bwd_kernel<KIND, NQ><<<grid, WT_THREADS, smem, st>>>(a);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

int wide_bwd_sweep(Context* ctx, int kind, const double* xp, long n, int d, const double* wcol, const double* ucol, double* rsum,
                   long ystride, double* gout, long gstride, int part, int nparts, cudaStream_t st) {
    WideBwdArgs a{};
    a.xp = xp; a.wcol = wcol; a.ucol = ucol; a.rsum = rsum; a.ystride = ystride; a.gout = gout; a.gstride = gstride;
    a.exp_tab = ctx->exp_table;
    a.n = n; a.d = d; a.kp = wide_kp(d); a.w = packed_width(d); a.part = part; a.nparts = nparts;
    a.nb_rows = (n + WB_ROWS - 1) / WB_ROWS;
    a.n_chunks = (n + WT_CHUNK - 1) / WT_CHUNK;
    a.nitems = 8 * a.n_chunks * (a.n_chunks + 1);
    long my_items = (a.nitems - part + nparts - 1) / nparts;
    if (my_items <= 0) return CGLB_OK;
    int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    const int nq = (a.kp + 7) / 8;
    if (nq > 16) {
        set_error("wide backward sweep: d=%d not supported (d <= 128)", d);
        return CGLB_ERR_UNSUPPORTED;
    }
#define WB(K)                                                     \
    (nq <= 8 ? launch_wide_bwd<K, 8>(ctx, a, grid, st)            \
             : nq <= 12 ? launch_wide_bwd<K, 12>(ctx, a, grid, st) : launch_wide_bwd<K, 16>(ctx, a, grid, st))
    return kind == CGLB_MATERN32 ? WB(CGLB_MATERN32) : WB(CGLB_RBF);
#undef WB
}

}  // namespace cglb
