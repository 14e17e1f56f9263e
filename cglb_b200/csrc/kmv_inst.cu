// One translation unit per input dimension D (compiled with -DCGLB_KMV_D=<d>), so the 16 x 2 x 8
// template instantiations of the sweeps build in parallel.
#include "kmv_impl.cuh"
#include "dsweep_impl.cuh"
#include "f32sweep_impl.cuh"
#include "kmv_multi_impl.cuh"

#ifndef CGLB_KMV_D
#error "compile with -DCGLB_KMV_D=<d>"
#endif

namespace cglb {

#define CGLB_CAT2(a, b) a##b
#define CGLB_CAT(a, b) CGLB_CAT2(a, b)

int CGLB_CAT(sweep_d, CGLB_KMV_D)(Context* ctx, int kind, int mode, const SweepArgs& a, cudaStream_t st) {
    constexpr int D = CGLB_KMV_D;
    if (mode == 3) {      // symmetric forward sweep with the distance contraction on DMMA (dsweep_impl.cuh)
        if constexpr (D >= 2) {
            constexpr int DP = SmemLayout<D>::DP;
            return kind == CGLB_MATERN32 ? run_dsweep<CGLB_MATERN32, DP>(ctx, a, st) : run_dsweep<CGLB_RBF, DP>(ctx, a, st);
        } else {
            set_error("dsweep: d=%d is not instantiated", D);
            return CGLB_ERR_UNSUPPORTED;
        }
    }
    if (mode == 4) {      // symmetric backward sweep on DMMA (dsweep_impl.cuh)
        if constexpr (D >= 2) {
            return kind == CGLB_MATERN32 ? run_dbwd<CGLB_MATERN32, D>(ctx, a, st) : run_dbwd<CGLB_RBF, D>(ctx, a, st);
        } else {
            set_error("dbwd: d=%d is not instantiated", D);
            return CGLB_ERR_UNSUPPORTED;
        }
    }
    if (mode == 5) return kind == CGLB_MATERN32 ? run_multi<CGLB_MATERN32, D, 2>(ctx, a, st) : run_multi<CGLB_RBF, D, 2>(ctx, a, st);
    if (mode == 6) return kind == CGLB_MATERN32 ? run_multi<CGLB_MATERN32, D, 4>(ctx, a, st) : run_multi<CGLB_RBF, D, 4>(ctx, a, st);
    if (kind == CGLB_MATERN32) {
        if (mode == 0) return run_fwd<CGLB_MATERN32, D, true>(ctx, a, st);
        if (mode == 1) return run_fwd<CGLB_MATERN32, D, false>(ctx, a, st);
        return run_bwd<CGLB_MATERN32, D>(ctx, a, st);
    } else {
        if (mode == 0) return run_fwd<CGLB_RBF, D, true>(ctx, a, st);
        if (mode == 1) return run_fwd<CGLB_RBF, D, false>(ctx, a, st);
        return run_bwd<CGLB_RBF, D>(ctx, a, st);
    }
}


int CGLB_CAT(f32_d, CGLB_KMV_D)(Context* ctx, int kind, const SweepArgsF32& a, cudaStream_t st) {
    constexpr int D = CGLB_KMV_D;
    return kind == CGLB_MATERN32 ? run_f32<CGLB_MATERN32, D>(ctx, a, st) : run_f32<CGLB_RBF, D>(ctx, a, st);
}

int CGLB_CAT(f32_bwd_d, CGLB_KMV_D)(Context* ctx, int kind, const BwdArgsF32& a, cudaStream_t st) {
    constexpr int D = CGLB_KMV_D;
    return kind == CGLB_MATERN32 ? run_f32_bwd<CGLB_MATERN32, D>(ctx, a, st) : run_f32_bwd<CGLB_RBF, D>(ctx, a, st);
}

int CGLB_CAT(knm_d, CGLB_KMV_D)(Context* ctx, int kind, int bwd, const KnmArgs& a, cudaStream_t st) {
    constexpr int D = CGLB_KMV_D;
    if (kind == CGLB_MATERN32) return run_knm<CGLB_MATERN32, D>(ctx, bwd, a, st);
    return run_knm<CGLB_RBF, D>(ctx, bwd, a, st);
}

}  // namespace cglb
