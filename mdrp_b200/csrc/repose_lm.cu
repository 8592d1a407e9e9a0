// repose_lm.cu — the LM kernels' translation unit.  Built WITHOUT -fmad=false so nvcc contracts the
// Jacobian arithmetic into FMAs (half the FP64 instructions); everything that must match the
// reference bit for bit lives in repose_b200.cu, which is built with -fmad=false.
#include "rp_lm_kernel.cuh"

namespace rp {

template <class K>
static int occupancy_grid(int sms, K kernel, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb * sms;
}

int launch_lm_kernel(int sms, int variant, const LMArgs &a, cudaStream_t st) {
    switch (variant) {
    case RP_CALIB:
        lm_kernel<RP_CALIB, 7><<<occupancy_grid(sms, lm_kernel<RP_CALIB, 7>, LM_THREADS), LM_THREADS, 0, st>>>(a);
        break;
    case RP_CALIB_SHIFT:
        lm_kernel<RP_CALIB_SHIFT, 9><<<occupancy_grid(sms, lm_kernel<RP_CALIB_SHIFT, 9>, LM_THREADS), LM_THREADS, 0, st>>>(a);
        break;
    case RP_SHARED:
        lm_kernel<RP_SHARED, 8><<<occupancy_grid(sms, lm_kernel<RP_SHARED, 8>, LM_THREADS), LM_THREADS, 0, st>>>(a);
        break;
    default:
        lm_kernel<RP_VARYING, 9><<<occupancy_grid(sms, lm_kernel<RP_VARYING, 9>, LM_THREADS), LM_THREADS, 0, st>>>(a);
        break;
    }
    return (int)cudaGetLastError();
}

}  // namespace rp
