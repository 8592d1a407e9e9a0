// rp_types.cuh — constants, per-pair parameters and packed point layouts shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include "rp_common.cuh"

namespace rp {

constexpr int SEG = 1024;          // RANSAC iterations per solve block (4 rounds of 256 threads)
constexpr int SOLVE_THREADS = 256;
constexpr int HB = 128;            // hypotheses per scoring work item
constexpr int SCORE_THREADS = 256;
constexpr int SCORE_WARPS = SCORE_THREADS / 32;
constexpr int PT = 4;              // correspondences per thread per slice
constexpr int EV = 256;            // trigger events kept per pair (expected ~2 ln H ~ 20)

struct PairParams {
    long long off;        // first correspondence of this pair in the packed arrays
    int n;                // number of correspondences
    int valid;            // n >= 3
    double thr;           // max_epipolar_error in normalised units
    double sq_thr;        // thr^2
    double scale_reproj;  // (thr_epi/thr_reproj)^2 or 0
    double lo_loss_scale; // loss_scale of the LO refinement (see varying-focal note in DESIGN.md)
    double final_loss_scale;
    double nscale;        // focal variants: normalize_points scale (1 for calibrated)
    double Mmax, mmax;    // bounds used by the FP32 filter
    long long pbase;      // first entry of this pair in the pair-interleaved FP32 layout (units of 2 points)
};

struct alignas(16) Pt64 {
    double x1_0, x1_1, x2_0, x2_1;
};
struct Bear {
    double b1x, b1y, b1z, b2x, b2y, b2z;
};

// ---------------------------------------------------------------------------------------------
// warp helpers
RP_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
RP_D double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace rp
