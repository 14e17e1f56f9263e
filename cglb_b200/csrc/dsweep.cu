// Policy of the DMMA-distance sweep (kernel: dsweep_impl.cuh, instantiated per input dimension in kmv_inst.cu).
#include "common.cuh"

namespace cglb {

// true if the DMMA sweep should handle this symmetric K*v: d = 10, 11, 13..32 and enough work items to fill the GPU
// (nparts = 0: skip the size test -- CGLB_DSWEEP=2, tests of small shapes)
bool dsweep_supported(const Context* ctx, int d, long n, int nparts) {
    if (d > CGLB_MAX_REGISTER_D) return false;
    if (nparts <= 0) return d >= 2;       // forced: every instantiated width
    // measured on B200 (tools/dev_time_sweeps.py, n = 100k, Gpairs/s DMMA vs register): d = 8 1114 vs 1227, 9 941 vs 967,
    // 10 950 vs 926, 11 961 vs 914, 12 842 vs 865 (13 useful slots in 4 k-steps), 13 841 vs 758, 16 916 vs 785,
    // 17 762 vs 666, 24 653 vs 562, 29 606 vs 509, 32 550 vs 477
    if (d < 10 || d == 12) return false;
    const long n_chunks = (n + 1023) / 1024;
    const long items = 4 * n_chunks * (n_chunks + 1) / 2 / nparts;
    return items >= 8L * ctx->num_sms;
}

// same question for the backward sweep (dmma_bwd_kernel)
bool dbwd_supported(const Context* ctx, int d, long n, int nparts) {
    if (d > CGLB_MAX_REGISTER_D) return false;
    if (nparts <= 0) return d >= 2;
    // measured on B200 (tools/dev_time_sweeps.py, Gpairs/s DMMA vs register): d = 3 833 vs 854, 4 855 vs 931, 5 741 vs 730,
    // 6 863 vs 779, 7 742 vs 645, 8 772 vs 697, 9 564 vs 573 (9 coordinates in two 8-slot tiles), 10 562 vs 557,
    // 11 568 vs 515, 13 525 vs 388, 16 558 vs 359, 19 407 vs 294, 24 365 vs 233, 32 318 vs 170
    if (d < 6 || d == 9) return false;
    const long n_chunks = (n + 1023) / 1024;
    const long items = 4 * n_chunks * (n_chunks + 1) / 2 / nparts;
    return items >= 8L * ctx->num_sms;
}

}  // namespace cglb
