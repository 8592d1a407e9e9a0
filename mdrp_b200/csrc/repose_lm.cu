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
static void launch_variant(int sms, bool warp_per_problem, const LMArgs &a, cudaStream_t st) {
    if (warp_per_problem)
        lm_kernel<VARIANT, NP, 32><<<occupancy_grid(sms, lm_kernel<VARIANT, NP, 32>, 32), 32, 0, st>>>(a);
    else
        lm_kernel<VARIANT, NP, RP_LM_THREADS><<<occupancy_grid(sms, lm_kernel<VARIANT, NP, RP_LM_THREADS>, RP_LM_THREADS),
                                                  RP_LM_THREADS, 0, st>>>(a);
}

int launch_lm_kernel(int sms, int variant, bool warp_per_problem, const LMArgs &a, cudaStream_t st) {
    switch (variant) {
    case RP_CALIB: launch_variant<RP_CALIB, 7>(sms, warp_per_problem, a, st); break;
    case RP_CALIB_SHIFT: launch_variant<RP_CALIB_SHIFT, 9>(sms, warp_per_problem, a, st); break;
    case RP_SHARED: launch_variant<RP_SHARED, 8>(sms, warp_per_problem, a, st); break;
    default: launch_variant<RP_VARYING, 9>(sms, warp_per_problem, a, st); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace rp
