// repose_lm.cu — the LM kernels' translation unit.  Built WITHOUT -fmad=false so nvcc contracts the
// Jacobian arithmetic into FMAs (half the FP64 instructions); everything that must match the
// reference bit for bit lives in repose_b200.cu, which is built with -fmad=false.
#include "rp_lm_kernel.cuh"

#ifndef RP_LM_THREADS
#define RP_LM_THREADS 128   // block-per-problem variant; registers per thread = 65536 / (threads x blocks per SM)
#endif

namespace rp {

template <class K>
static int occupancy_grid(int sms, K kernel, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb * sms;
}

template <int VARIANT, int NP>
static void launch_variant(int sms, const LMArgs &a, cudaStream_t st) {
    // the LO refinement (use_final = 0) always uses the TRUNCATED loss: compile-time specialisation
    // (the one-problem-per-pair launches have too few problems for warp granularity: 2-3 per warp leave a long tail)
    if (a.warp_kernel && (a.prob_list || a.warp_single) && !a.mask && !a.use_final)
        lm_warp_kernel<VARIANT, NP, RP_LOSS_TRUNCATED>
            <<<occupancy_grid(sms, lm_warp_kernel<VARIANT, NP, RP_LOSS_TRUNCATED>, LMW_WPB[VARIANT] * 32), LMW_WPB[VARIANT] * 32, 0, st>>>(a);
    else if (!a.use_final)
        lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_TRUNCATED>
            <<<occupancy_grid(sms, lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_TRUNCATED>, RP_LM_THREADS), RP_LM_THREADS, 0, st>>>(a);
    // the final refinement with the two losses the reference's drivers and defaults use, fixed at compile time (no loss
    // switch and no unused loss code in the loop: the launch is sensitive to instruction fetch), any other loss at run time
    else if (a.loss_type == RP_LOSS_TRUNCATED_CAUCHY)
        lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_TRUNCATED_CAUCHY>
            <<<occupancy_grid(sms, lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_TRUNCATED_CAUCHY>, RP_LM_THREADS), RP_LM_THREADS, 0, st>>>(a);
    else if (a.loss_type == RP_LOSS_CAUCHY)
        lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_CAUCHY>
            <<<occupancy_grid(sms, lm_kernel<VARIANT, NP, RP_LM_THREADS, RP_LOSS_CAUCHY>, RP_LM_THREADS), RP_LM_THREADS, 0, st>>>(a);
    else
        lm_kernel<VARIANT, NP, RP_LM_THREADS, -1>
            <<<occupancy_grid(sms, lm_kernel<VARIANT, NP, RP_LM_THREADS, -1>, RP_LM_THREADS), RP_LM_THREADS, 0, st>>>(a);
}

int launch_lm_kernel(int sms, int variant, const LMArgs &a, cudaStream_t st) {
    switch (variant) {
    case RP_CALIB: launch_variant<RP_CALIB, 7>(sms, a, st); break;
    case RP_CALIB_SHIFT: launch_variant<RP_CALIB_SHIFT, 9>(sms, a, st); break;
    case RP_SHARED: launch_variant<RP_SHARED, 8>(sms, a, st); break;
    default: launch_variant<RP_VARYING, 9>(sms, a, st); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace rp
