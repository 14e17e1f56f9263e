// K1 for d <= 32 with the distance contraction on DMMA.8x8x4 ("dsweep").
//
// Why: on sm_100a every non-FP64 instruction of these sweeps costs the shared FP64/DMMA pipe about one issue
// cycle (measured: profiles/dsweep_ncu_r01.md, profiles/fp64_issue_model_r01.txt, tools/fp64_issue_model.cu, tools/fp64_mix_model.cu).  In the
// register-resident sweep (kmv_impl.cuh) a Matern32 pair at d = 11 is 27 FP64 instructions + ~15 integer / LDS /
// MUFU instructions.  DMMA.8x8x4 performs the 256 FMAs of 8 warp-wide DFMAs in ONE instruction (same pipe
// time), so moving the d + 1 distance FMAs onto it removes ~11 instructions per pair and most of the LDS
// traffic; the kernel map runs on the accumulator fragments as in widek.cu.  Measured on B200 (Gpairs/s, DMMA vs
// register kernel): d = 11 961 vs 914, d = 15 841 vs 697, d = 19 766 vs 644, d = 23 708 vs 580, d = 27 655 vs
// 520; no gain for d < 10 (the register kernel keeps those).
//
// Works directly on the packed layout of cglb_pack_inputs (row width DP = d + 1 rounded up to even).  The
// contraction length is DP rounded up to a multiple of 4; when DP is not one (d = 12, 13, 16, 17, ...) the
// last k-step reads two doubles of the NEXT packed row, which the A side multiplies by 0.  DP = 12, 20, 28
// (d = 10, 11, 18, 19, 26, 27) are = 4 or 12 mod 16 and bank-conflict free; the other widths take 2- to 8-way
// conflicts on the 3-9 fragment loads per 8 columns, which the measurements below include.  The last slot of a packed row holds |b|^2: the A-side fragment carries 1.0
// there and -2 a_k elsewhere, the accumulator of the first DMMA starts at |a|^2, and the contraction yields
// q = |a|^2 + |b|^2 - 2 a.b with no extra FP64 slot.
//
//   CTA = 8 warps, warp = 32 rows (4 m-tiles, A fragments in registers for the whole item);
//   work item = (row block of 256 rows, chunk of 1024 columns), columns at/after the row block only; tiles
//   overlapping the row block are evaluated as ordered pairs (row sums only), tiles beyond it feed y_i and y_j;
//   64-column tiles stream through an 8-stage cp.async.bulk/mbarrier ring;
//   software pipeline over n-tiles (8 columns), across tile boundaries: the DMMAs of n-tile t+1 are issued
//   before the kernel map of n-tile t;
//   the warps never synchronise with each other: column sums leave each warp as one coalesced 256-byte RED
//   per 32 columns (transposing butterfly over the 8 lanes that share a column pair), row sums once per item.
#pragma once
#include "kmv_impl.cuh"

namespace cglb {

constexpr int DS_STAGES = 8;                  // 64-column tiles in the ring (4 in flight ahead of warp 0)
constexpr int DS_RPC = 4;                     // row blocks per column chunk (chunk = 4 x rows per item)
// forward sweep: packed widths whose 8-stage ring leaves room for TWO slabs of per-warp column sums (split-phase hand-over)
__host__ __device__ constexpr bool DS_SPLIT(int dp) { return dp <= 16; }

__device__ __forceinline__ void dmma884c(double (&d)[2], double a, double b, double c0, double c1) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d[0]), "=d"(d[1])
        : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// Work items and their order.  An item pairs the row block I (ROWS rows) with the column chunk c (DS_RPC x ROWS columns) for
// every I < DS_RPC (c + 1): the columns at or after the row block.  Round 1 enumerated them chunk-major over ALL row blocks
// (t = DS_RPC c (c + 1) / 2 + I); at n = 2M the rows a chunk pairs with (up to 192 MB of packed inputs) no longer fit the L2
// and were re-read from DRAM for every chunk: 123 GB per launch (profiles/ncu_r02/dsweep_d11_n2M_raw.csv).  Now the row blocks
// are grouped into SUPER-ROWS of R = DS_RPC Q blocks (Q = SweepArgs::sr_chunks chunks, ~48 MB of packed rows): within super-row s
// the order is chunk-major over its own row blocks only, so its rows stay L2-resident while the column chunks stream by once.
//   super-row s, local chunk k = c - s Q >= 0: row blocks s R .. min((s + 1) R, DS_RPC (c + 1)) - 1
//   items per local chunk: DS_RPC (k + 1) for k < Q (triangular part), R for k >= Q (rectangular part)
// cglb_b200/distributed.py: decode_strip_item is the host mirror (tests/test_host_logic.py: every pair exactly once).
template <int ROWS>
struct DCursor {
    static constexpr int DS_ROWS = ROWS, DS_CHUNK = DS_RPC * ROWS;
    long tau;
    long r0;        // first row of the row block
    long c0;        // first column of the first tile
    int tile, ntiles;
    bool valid;
    __device__ __forceinline__ static void decode(long t, long n_chunks, long Q, long& I, long& c) {
        const long R = DS_RPC * Q;
        long s = 0;
        for (;; ++s) {                                   // a handful of super-rows (4 at n = 2M)
            const long nc = n_chunks - s * Q;            // chunks this super-row pairs with
            const long kk = nc < Q ? nc : Q;
            const long cnt = DS_RPC * kk * (kk + 1) / 2 + (nc > Q ? R * (nc - Q) : 0);
            if (t < cnt || nc <= Q) break;               // (the last super-row takes whatever is left)
            t -= cnt;
        }
        const long tri = DS_RPC * Q * (Q + 1) / 2;
        long k, iloc;
        if (t < tri) {
            // chunk-major enumeration of {(i, k) : i < DS_RPC (k + 1)}: prefix(k) = DS_RPC k (k + 1) / 2
            k = (long)((sqrt(1.0 + 8.0 * (double)t / DS_RPC) - 1.0) * 0.5);
            while (DS_RPC * k * (k + 1) / 2 > t) --k;
            while (DS_RPC * (k + 1) * (k + 2) / 2 <= t) ++k;
            iloc = t - DS_RPC * k * (k + 1) / 2;
        } else {
            k = Q + (t - tri) / R;
            iloc = (t - tri) % R;
        }
        c = s * Q + k;
        I = s * R + iloc;
    }
    __device__ __forceinline__ void load_item(const SweepArgs& a, long n_chunks) {
        for (;; tau += gridDim.x) {
            const long t = tau * a.nparts + a.part;
            valid = t < a.nitems;
            if (!valid) return;
            long I, c;
            decode(t, n_chunks, a.sr_chunks, I, c);
            if (I >= a.nb_rows || c >= n_chunks) continue;
            r0 = I * DS_ROWS;
            long cbeg = c * DS_CHUNK;
            if (cbeg < r0) cbeg = r0;
            long cend = (c + 1) * DS_CHUNK;
            if (cend > a.ncols) cend = a.ncols;
            if (cbeg >= cend) continue;
            c0 = cbeg;
            ntiles = (int)((cend - cbeg + kBJ - 1) / kBJ);
            tile = 0;
            return;
        }
    }
    __device__ __forceinline__ void start(const SweepArgs& a, long n_chunks) { tau = blockIdx.x; load_item(a, n_chunks); }
    __device__ __forceinline__ void next_tile(const SweepArgs& a, long n_chunks) {
        if (++tile == ntiles) { tau += gridDim.x; load_item(a, n_chunks); }
    }
};

// chunks per super-row: ~48 MB of packed rows (a chunk is DS_RPC x ROWS rows of dp doubles), overridden by the "superrow"
// option (tests exercise several super-rows on small problems)
static inline long dsweep_superrow_chunks(const Context* ctx, int dp, int rows) {
    if (ctx->opt_superrow > 0) return ctx->opt_superrow;
    const long q = (48L << 20) / ((long)DS_RPC * rows * dp * (long)sizeof(double));
    return q < 1 ? 1 : q;
}

// WARPS warps x MT m-tiles (8 rows each) per warp; lane 0 of warp 0 also drives the TMA ring, as in the
// register-resident sweep (a dedicated producer warp and 12 / 16-warp shapes were measured and are slower,
// profiles/dsweep_ncu_r01.md, profiles/fp64_issue_model_r01.txt)
template <int KIND, int DP, int WARPS, int MT>
__global__ void __launch_bounds__(WARPS * 32, 1) dmma_sweep_kernel(const SweepArgs args, const long n_chunks) {
    constexpr int KS = (DP + 3) / 4;    // k-steps; slots DP .. 4 KS - 1 belong to the next packed row (A carries 0 there)
    constexpr int DS_WARPS = WARPS, DS_THREADS = WARPS * 32, DS_ROWS = WARPS * 8 * MT;
    constexpr int WROWS = 8 * MT;       // rows per warp
    using Cur = DCursor<DS_ROWS>;
    static_assert(DP % 2 == 0 && DP >= 4, "packed row width");
    static_assert(DS_THREADS == DS_ROWS, "thread t adds row r0 + t and the columns c0 + t + 256 k of its CTA's copy");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_tab[kExpTabBig];                                  // static: LDS with an immediate base
    double* s_x = reinterpret_cast<double*>(smem_raw);                    // [DS_STAGES][kBJ*DP]
    double* s_v = s_x + DS_STAGES * kBJ * DP;                               // [DS_STAGES][kBJ]
    // Fixed summation order (no run-to-run differences): inside an item every warp parks its column sums in its own row
    // of a slab and its row sums in s_rsum; once all warps have done so the sums are added -- columns over the warps in
    // warp order -- to THIS CTA's copy of y (args.y + blockIdx.x * ystride), the copies are summed in CTA order by a second
    // kernel.  Address X of the copy is only ever touched by thread X % 256 (row blocks and chunk starts are multiples of
    // 256), in the static order of the CTA's items.
    // SPLIT (the widths whose ring leaves room for two slabs): the hand-over is split-phase.  A warp ARRIVES on s_item when
    // its sums of item k are parked and goes on to item k + 1 (other slab); it waits for that phase only after it has issued
    // the global loads of the next item's row fragments, and then adds its share of item k.  No warp waits at a barrier with
    // nothing to do (a CTA barrier at every item end cost 2.2 % at d = 11).
    constexpr bool SPLIT = DS_SPLIT(DP);
    constexpr int NSLAB = SPLIT ? 2 : 1;
    double* s_slab = s_v + DS_STAGES * kBJ;                                 // [NSLAB][WARPS][chunk] per-warp column sums
    double* s_rsum = s_slab + NSLAB * WARPS * Cur::DS_CHUNK;                // [NSLAB][DS_ROWS] row sums
    long* s_desc = reinterpret_cast<long*>(s_rsum + NSLAB * DS_ROWS);       // [NSLAB][4]: r0, c0, ntiles of the parked item
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_desc + NSLAB * 4);
    uint64_t* s_empty = s_full + DS_STAGES;
    uint64_t* s_item = s_empty + DS_STAGES;                                 // [1] arrivals: one per warp and item
    double* const yb = args.y + (long)blockIdx.x * args.ystride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < kExpTabBig; i += DS_THREADS) s_tab[i] = args.exp_tab[kExpTabSmall + i];
    if (DP % 4 != 0) {
        // the last k-step of a column reads 2 doubles of the next packed row (times 0 from the A side): every
        // byte a fragment load can touch must hold a finite number before the first tile lands
        for (int i = tid; i < DS_STAGES * kBJ * DP + DS_STAGES * kBJ; i += DS_THREADS) s_x[i] = 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], DS_WARPS); }
        mbar_init(s_item, DS_WARPS);
        mbar_fence_init();
    }
    __syncthreads();
    if (DP % 4 != 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero fill before the TMA writes

    int pstage = 0, stage = 0;
    uint32_t pphase = 0, phase = 0;
    auto produce = [&](Cur& pc) {
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DP + kBJ) * sizeof(double)));
        tma_load_1d(s_x + pstage * kBJ * DP, args.xp_cols + j0 * DP, kBJ * DP * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_v + pstage * kBJ, args.vcol + j0, kBJ * sizeof(double), &s_full[pstage]);
        if (++pstage == DS_STAGES) { pstage = 0; pphase ^= 1; }
        pc.next_tile(args, n_chunks);
    };

    Cur cc;
    cc.start(args, n_chunks);
    Cur pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < DS_STAGES / 2; ++i) produce(pc);
    }
    const double var = args.variance;
    int item = 0;                 // items this CTA has parked so far (slab / phase = item & 1)
    // adds the parked sums of item `it` (slab it & 1): every thread its own columns and its own row
    auto add_parked = [&](int it) {
        const int b = SPLIT ? (it & 1) : 0;
        const long* dsc = s_desc + b * 4;
        const long p_r0 = dsc[0], p_c0 = dsc[1];
        const int p_ncols = (int)dsc[2] * kBJ;
        const double* slab = s_slab + b * WARPS * Cur::DS_CHUNK;
        for (int col = tid; col < p_ncols; col += DS_THREADS) {
            const long jc = p_c0 + col;
            if (p_c0 + (col & ~(kBJ - 1)) >= p_r0 + DS_ROWS && jc < args.ncols) {      // tiles beyond the row block only
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) s += slab[w * Cur::DS_CHUNK + col];
                atomicAdd(yb + jc, var * s);
            }
        }
        if (p_r0 + tid < args.nrows) atomicAdd(yb + p_r0 + tid, var * s_rsum[b * DS_ROWS + tid]);
    };

    while (cc.valid) {
        const long r0 = cc.r0;
        const long c0 = cc.c0;
        const int ntiles = cc.ntiles;
        // A fragments: lane (g, t4) holds A[row g + 8 i][4 ks + t4], A = (-2 a_0 .. -2 a_{DP-2}, 1)
        double af[MT][KS], na[MT], vrow[MT], racc[MT];
        bool live[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const long row = r0 + warp * WROWS + i * 8 + g;
            live[i] = row < args.nrows;
            const double* src = args.xp_rows + (live[i] ? row : 0) * DP;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int k = 4 * ks + t4;
                const double val = (live[i] && k < DP) ? __ldg(src + k) : 0.0;
                af[i][ks] = (k == DP - 1) ? 1.0 : -2.0 * val;
            }
            na[i] = live[i] ? __ldg(src + DP - 1) : 0.0;
            vrow[i] = live[i] ? __ldg(args.vcol + row) : 0.0;
            racc[i] = 0.0;
        }
        // One n-tile = 8 columns x the warp's 8 MT rows: KS chained DMMAs per m-tile, started from |a|^2.
        auto dmma_ntile = [&](const double* sx, int col0, double (&acc)[MT][2]) {
            double bf[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) bf[ks] = sx[(col0 + g) * DP + 4 * ks + t4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                dmma884c(acc[i], af[i][0], bf[0], na[i], na[i]);
#pragma unroll
                for (int ks = 1; ks < KS; ++ks) dmma884c(acc[i], af[i][ks], bf[ks], acc[i][0], acc[i][1]);
            }
        };
        if (SPLIT && item > 0) {
            // the row fragments above are in flight: now wait for every warp to have parked the previous item, add it
            mbar_wait(s_item, (uint32_t)((item - 1) & 1));
            add_parked(item - 1);
        }
        // Software pipeline over n-tiles, across tile boundaries: the DMMAs of n-tile t+1 are issued before the
        // kernel map of n-tile t, so the map always has all 2 MT chains of a full n-tile to interleave.
        if (tid == 0) produce(pc);
        __syncwarp();
        mbar_wait(&s_full[stage], phase);
        double accn[MT][2];
        dmma_ntile(s_x + stage * kBJ * DP, 0, accn);
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            const double* sx = s_x + stage * kBJ * DP;
            const double* sv = s_v + stage * kBJ;
            const long j0 = c0 + (long)tile * kBJ;
            const bool offdiag = j0 >= r0 + DS_ROWS;
            const int nstage = (stage + 1 == DS_STAGES) ? 0 : stage + 1;
            const uint32_t nphase = (nstage == 0) ? (phase ^ 1) : phase;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                double c8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col0 = half * 32 + j * 8;
                    double acc[MT][2];
#pragma unroll
                    for (int i = 0; i < MT; ++i) { acc[i][0] = accn[i][0]; acc[i][1] = accn[i][1]; }
                    const double2 vv = *reinterpret_cast<const double2*>(sv + col0 + 2 * t4);
                    if (j < 3 || half == 0) {
                        dmma_ntile(sx, col0 + 8, accn);
                    } else if (tile + 1 < ntiles) {
                        if (tid == 0) produce(pc);
                        __syncwarp();
                        mbar_wait(&s_full[nstage], nphase);
                        dmma_ntile(s_x + nstage * kBJ * DP, 0, accn);
                    }
                    double cs0 = 0.0, cs1 = 0.0;
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const double k0 = kappa<KIND, 10>(acc[i][0], s_tab);
                        const double k1 = kappa<KIND, 10>(acc[i][1], s_tab);
                        racc[i] = fma(k0, vv.x, racc[i]);
                        racc[i] = fma(k1, vv.y, racc[i]);
                        cs0 = fma(k0, vrow[i], cs0);
                        cs1 = fma(k1, vrow[i], cs1);
                    }
                    c8[2 * j] = cs0;
                    c8[2 * j + 1] = cs1;
                }
                if (offdiag) {
                    // sum the 8 column partials over the 8 lanes sharing t4 (lane bits 2..4): transposing butterfly
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
                    const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // = 2 j + e
                    // every lane now holds one of the 32 column sums of this half tile (over the warp's rows): one
                    // conflict-free 256-byte store into the warp's slab row, no CTA barrier inside the item
                    s_slab[((SPLIT ? (item & 1) : 0) * WARPS + warp) * Cur::DS_CHUNK + tile * kBJ + half * 32 + (idx >> 1) * 8 + 2 * t4 +
                           (idx & 1)] = c8[0];
                }
            }
            // this warp is done with the stage (the first n-tile of the next stage has been read already)
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            stage = nstage;
            phase = nphase;
        }
        // item end: park the row sums (over the 4 lanes sharing g; every warp owns its 32 rows) and the item's coordinates
        {
            const int b = SPLIT ? (item & 1) : 0;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                double s = racc[i];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (t4 == 0) s_rsum[b * DS_ROWS + warp * WROWS + i * 8 + g] = s;
            }
            if (tid == 0) { s_desc[b * 4] = r0; s_desc[b * 4 + 1] = c0; s_desc[b * 4 + 2] = ntiles; }
        }
        if (SPLIT) {
            __syncwarp();
            if (lane == 0) mbar_arrive(s_item);       // release: this warp's sums of the item are visible to whoever waits
        } else {
            __syncthreads();
            add_parked(item);
            __syncthreads();                          // the slab is free again
        }
        ++item;
        cc.tau += gridDim.x;
        cc.load_item(args, n_chunks);
    }
    if (SPLIT && item > 0) {
        mbar_wait(s_item, (uint32_t)((item - 1) & 1));
        add_parked(item - 1);
    }
}

static inline size_t dsweep_smem_bytes(int dp, int warps, int chunk, int rows) {
    const int nslab = DS_SPLIT(dp) ? 2 : 1;
    return (size_t)(DS_STAGES * kBJ * dp + DS_STAGES * kBJ + nslab * (warps * chunk + rows + 4)) * sizeof(double) +
           (2 * DS_STAGES + 1) * sizeof(uint64_t);
}

template <int KIND, int DP, int WARPS = 8, int MT = 4>
static int run_dsweep(Context* ctx, SweepArgs a, cudaStream_t st) {
    constexpr int ROWS = WARPS * 8 * MT, CHUNK = DS_RPC * ROWS, THREADS = WARPS * 32;
    a.nb_rows = (a.nrows + ROWS - 1) / ROWS;
    const long n_chunks = (a.ncols + CHUNK - 1) / CHUNK;
    a.nb_cols = n_chunks;
    a.nitems = DS_RPC * n_chunks * (n_chunks + 1) / 2;
    a.sr_chunks = dsweep_superrow_chunks(ctx, DP, ROWS);
    auto kern = dmma_sweep_kernel<KIND, DP, WARPS, MT>;
    const size_t smem = dsweep_smem_bytes(DP, WARPS, CHUNK, ROWS);
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    const int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, THREADS, smem, st>>>(a, n_chunks);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}


// ---------------------------------------------------------------------------------------------
// K2 (fused backward sweep) on DMMA for 10 <= d <= 32.  Same tiling as dmma_sweep_kernel; per
// n-tile (8 columns x the warp's 32 rows):
//   S = A_I A_J^T on DMMA -> (kappa, e') on the fragments; omega = u_i w_j + w_i u_j; c = e' omega
//   gvar += kappa omega;  row sums of c in registers, column sums through the transposing butterfly + one RED
//   cross term: Y_I += C A_J as a SECOND DMMA product.  The contraction index (the 8 columns) can be permuted
//   freely, so the thread's own two c values (columns 2 t4, 2 t4 + 1 of the C fragment) serve directly as the
//   A fragment of two k-steps whose B fragments are rows 2 t4 / 2 t4 + 1 of the column tile: no shuffles, no
//   shared-memory staging of C.  X_q = sum_i a_iq Y_iq is formed once per item.
// Tiles overlapping the row block visit ordered pairs with halved weights and doubled row sums, as in the
// register-resident kernel.
// ---------------------------------------------------------------------------------------------
template <int KIND, int D, int WARPS, int MT>
__global__ void __launch_bounds__(WARPS * 32, 1) dmma_bwd_kernel(const SweepArgs args, const long n_chunks) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int KS = (DP + 3) / 4;
    constexpr int NQ = (D + 7) / 8;     // 8-slot coordinate tiles of Y (slots >= D are ignored)
    constexpr int DS_THREADS = WARPS * 32, DS_ROWS = WARPS * 8 * MT;
    constexpr int WROWS = 8 * MT;
    using Cur = DCursor<DS_ROWS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_tab[kExpTabBig];
    __shared__ double s_red[WARPS][4 * NQ + 1];
    double* s_x = reinterpret_cast<double*>(smem_raw);                    // [DS_STAGES][kBJ*DP]
    double* s_wu = s_x + DS_STAGES * kBJ * DP;                            // [DS_STAGES][2][kBJ]  (w, u)
    double* s_slab = s_wu + DS_STAGES * 2 * kBJ;                          // [WARPS][chunk] per-warp column sums of one item
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_slab + WARPS * Cur::DS_CHUNK);
    uint64_t* s_empty = s_full + DS_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < kExpTabBig; i += DS_THREADS) s_tab[i] = args.exp_tab[kExpTabSmall + i];
    // fragment loads may run past a packed row (into the next row, the next stage or the w/u slices): all finite
    for (int i = tid; i < DS_STAGES * kBJ * DP + DS_STAGES * 2 * kBJ; i += DS_THREADS) s_x[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    int pstage = 0, stage = 0;
    uint32_t pphase = 0, phase = 0;
    auto produce = [&](Cur& pc) {
        if (!pc.valid) return;
        mbar_wait(&s_empty[pstage], pphase ^ 1);
        const long j0 = pc.c0 + (long)pc.tile * kBJ;
        mbar_expect_tx(&s_full[pstage], (uint32_t)((kBJ * DP + 2 * kBJ) * sizeof(double)));
        tma_load_1d(s_x + pstage * kBJ * DP, args.xp_cols + j0 * DP, kBJ * DP * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_wu + pstage * 2 * kBJ, args.vcol + j0, kBJ * sizeof(double), &s_full[pstage]);
        tma_load_1d(s_wu + pstage * 2 * kBJ + kBJ, args.ucol + j0, kBJ * sizeof(double), &s_full[pstage]);
        if (++pstage == DS_STAGES) { pstage = 0; pphase ^= 1; }
        pc.next_tile(args, n_chunks);
    };

    Cur cc;
    cc.start(args, n_chunks);
    Cur pc = cc;
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < DS_STAGES / 2; ++i) produce(pc);
    }
    double gql[NQ][2], gvar = 0.0;       // thread-private: -2 X_q for slots 8 nq + 2 t4 + e, variance sum
#pragma unroll
    for (int nq = 0; nq < NQ; ++nq) gql[nq][0] = gql[nq][1] = 0.0;

    while (cc.valid) {
        const long r0 = cc.r0;
        const long c0 = cc.c0;
        const int ntiles = cc.ntiles;
        double af[MT][KS], na[MT], ui[MT], wi[MT], racc[MT], yq[MT][NQ][2];
        bool live[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const long row = r0 + warp * WROWS + i * 8 + g;
            live[i] = row < args.nrows;
            const double* src = args.xp_rows + (live[i] ? row : 0) * DP;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int k = 4 * ks + t4;
                const double val = (live[i] && k < DP) ? __ldg(src + k) : 0.0;
                af[i][ks] = (k == DP - 1) ? 1.0 : -2.0 * val;
            }
            na[i] = live[i] ? __ldg(src + DP - 1) : 0.0;
            ui[i] = live[i] ? __ldg(args.ucol + row) : 0.0;
            wi[i] = live[i] ? __ldg(args.vcol + row) : 0.0;
            racc[i] = 0.0;
#pragma unroll
            for (int nq = 0; nq < NQ; ++nq) yq[i][nq][0] = yq[i][nq][1] = 0.0;
        }
#pragma unroll 1
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tid == 0) produce(pc);
            __syncwarp();
            mbar_wait(&s_full[stage], phase);
            const double* sx = s_x + stage * kBJ * DP;
            const double* sw = s_wu + stage * 2 * kBJ;
            const double* su = sw + kBJ;
            const long j0 = c0 + (long)tile * kBJ;
            const bool offdiag = j0 >= r0 + DS_ROWS;
            const double half = offdiag ? 1.0 : 0.5, rs = offdiag ? 1.0 : 2.0;
            double hu[MT], hw[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) { hu[i] = half * ui[i]; hw[i] = half * wi[i]; }
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                double c8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col0 = hh * 32 + j * 8;
                    double bf[KS];
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) bf[ks] = sx[(col0 + g) * DP + 4 * ks + t4];
                    double acc[MT][2];
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        dmma884c(acc[i], af[i][0], bf[0], na[i], na[i]);
#pragma unroll
                        for (int ks = 1; ks < KS; ++ks) dmma884c(acc[i], af[i][ks], bf[ks], acc[i][0], acc[i][1]);
                    }
                    const double2 w2 = *reinterpret_cast<const double2*>(sw + col0 + 2 * t4);
                    const double2 u2 = *reinterpret_cast<const double2*>(su + col0 + 2 * t4);
                    double b2[2][NQ];      // B fragments of the cross-term product: rows 2 t4 + e of the column tile
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int nq = 0; nq < NQ; ++nq) b2[e][nq] = sx[(col0 + 2 * t4 + e) * DP + 8 * nq + g];
                    double cs0 = 0.0, cs1 = 0.0;
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        double kap0, ew0, kap1, ew1;
                        kappa_and_dweight<KIND, 10>(acc[i][0], s_tab, kap0, ew0);
                        kappa_and_dweight<KIND, 10>(acc[i][1], s_tab, kap1, ew1);
                        const double om0 = fma(hw[i], u2.x, hu[i] * w2.x);
                        const double om1 = fma(hw[i], u2.y, hu[i] * w2.y);
                        const double cw0 = ew0 * om0, cw1 = ew1 * om1;
                        gvar = fma(kap0, om0, gvar);
                        gvar = fma(kap1, om1, gvar);
                        racc[i] = fma(rs, cw0 + cw1, racc[i]);
                        cs0 += cw0;
                        cs1 += cw1;
#pragma unroll
                        for (int nq = 0; nq < NQ; ++nq) {
                            dmma884c(yq[i][nq], cw0, b2[0][nq], yq[i][nq][0], yq[i][nq][1]);
                            dmma884c(yq[i][nq], cw1, b2[1][nq], yq[i][nq][0], yq[i][nq][1]);
                        }
                    }
                    c8[2 * j] = cs0;
                    c8[2 * j + 1] = cs1;
                }
                if (offdiag) {
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
                    const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                    s_slab[warp * Cur::DS_CHUNK + tile * kBJ + hh * 32 + (idx >> 1) * 8 + 2 * t4 + (idx & 1)] = c8[0];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            if (++stage == DS_STAGES) { stage = 0; phase ^= 1; }
        }
        // item end: column sums of the tiles beyond the row block, summed over the warps in warp order and added to this
        // CTA's copy of R (fixed order, see dmma_sweep_kernel; the pointer is formed here: no register is left in the loop)
        __syncthreads();
        double* const yb = args.y + (long)blockIdx.x * args.ystride;
        for (int col = tid; col < ntiles * kBJ; col += DS_THREADS) {
            const long jc = c0 + col;
            if (c0 + (col & ~(kBJ - 1)) >= r0 + DS_ROWS && jc < args.ncols) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) s += s_slab[w * Cur::DS_CHUNK + col];
                atomicAdd(yb + jc, s);
            }
        }
        // row sums (over the 4 lanes sharing g) and the cross term X_q = sum_i a_iq Y_iq
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            double s = racc[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const long row = r0 + warp * WROWS + i * 8 + g;
            if (t4 == 0 && live[i]) atomicAdd(yb + row, s);
            if (live[i]) {
                const double* src = args.xp_rows + row * DP;
#pragma unroll
                for (int nq = 0; nq < NQ; ++nq)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int q = 8 * nq + 2 * t4 + e;
                        if (q < D) gql[nq][e] = fma(-2.0 * __ldg(src + q), yq[i][nq][e], gql[nq][e]);
                    }
            }
        }
        __syncthreads();
        cc.tau += gridDim.x;
        cc.load_item(args, n_chunks);
    }

    // block reduction: slot q = 8 nq + 2 t4 + e lives in the lanes with this t4 -> sum over g (lane bits 2..4), then warps
#pragma unroll
    for (int nq = 0; nq < NQ; ++nq)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double s = gql[nq][e];
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            gql[nq][e] = s;
        }
    {
        const double s = warp_sum(gvar);
        if (lane == 0) s_red[warp][4 * NQ] = s;
    }
    // lanes 0..3 (g = 0) hold the warp sums of slots 8 nq + 2 t4 + e; slots >= D are skipped below
    __shared__ double s_q[WARPS][8 * NQ];
    if (g == 0) {
#pragma unroll
        for (int nq = 0; nq < NQ; ++nq) {
            s_q[warp][8 * nq + 2 * t4] = gql[nq][0];
            s_q[warp][8 * nq + 2 * t4 + 1] = gql[nq][1];
        }
    }
    __syncthreads();
    double* const gb = args.gout + (long)blockIdx.x * args.gstride;       // this CTA's slot (summed in CTA order afterwards)
    if (tid < D) {
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) s += s_q[w][tid];
        gb[tid] = s;
    } else if (tid == D) {
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) s += s_red[w][4 * NQ];
        gb[D] = s;
    }
}

template <int KIND, int D, int WARPS = 8, int MT = 4>
static int run_dbwd(Context* ctx, SweepArgs a, cudaStream_t st) {
    constexpr int DP = SmemLayout<D>::DP;
    constexpr int ROWS = WARPS * 8 * MT, CHUNK = DS_RPC * ROWS, THREADS = WARPS * 32;
    a.nb_rows = (a.nrows + ROWS - 1) / ROWS;
    const long n_chunks = (a.ncols + CHUNK - 1) / CHUNK;
    a.nb_cols = n_chunks;
    a.nitems = DS_RPC * n_chunks * (n_chunks + 1) / 2;
    a.sr_chunks = dsweep_superrow_chunks(ctx, DP, ROWS);
    auto kern = dmma_bwd_kernel<KIND, D, WARPS, MT>;
    const size_t smem = (size_t)(DS_STAGES * kBJ * DP + DS_STAGES * 2 * kBJ + WARPS * CHUNK) * sizeof(double) + 2 * DS_STAGES * sizeof(uint64_t);
    CGLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long my_items = (a.nitems - a.part + a.nparts - 1) / a.nparts;
    if (my_items <= 0) return CGLB_OK;
    const int grid = (int)(my_items < ctx->num_sms ? my_items : ctx->num_sms);
    kern<<<grid, THREADS, smem, st>>>(a, n_chunks);
    ctx->launches++;
    CGLB_LAUNCH_OK();
    return CGLB_OK;
}

}  // namespace cglb
