// rp_score.cuh — per-(hypothesis, correspondence) scoring primitives.
//
//   SC  compute_sampson_msac_score(CameraPose,…) so@0x4f61d0  Sampson + cheirality
//   SF  compute_sampson_msac_score(Matrix3d F,…)  so@0x4f65d0  Sampson only
//   I1  get_inliers so@0x4f7a10 / so@0x4f77f0
//
// The inlier decision of the reference is `fl(fl(C*C)/fl(Cx+Cy)) < thr^2 [&& cheirality]` in
// FP64 without FMA.  The kernels reproduce it bit for bit in two tiers:
//   tier 0  an FP32 FMA filter that only ever answers "certainly an outlier" (rigorous error
//           bound below) — the common case, ~19 FFMA per point-score;
//   tier 1  the exact FP64 evaluation in the reference's operation order for everything else.
// SURVEY.md §8a rows SC/SF/I1.
#pragma once
#include "rp_common.cuh"

namespace rp {

// exact Sampson r^2 in the reference's operation order (no FMA)
RP_HD double sampson_r2_exact(const M3 &E, double x1_0, double x1_1, double x2_0, double x2_1) {
    const double Ex1_0 = E.r0.x * x1_0 + E.r0.y * x1_1 + E.r0.z;
    const double Ex1_1 = E.r1.x * x1_0 + E.r1.y * x1_1 + E.r1.z;
    const double Ex1_2 = E.r2.x * x1_0 + E.r2.y * x1_1 + E.r2.z;
    const double Ex2_0 = E.r0.x * x2_0 + E.r1.x * x2_1 + E.r2.x;
    const double Ex2_1 = E.r0.y * x2_0 + E.r1.y * x2_1 + E.r2.y;
    const double C = x2_0 * Ex1_0 + x2_1 * Ex1_1 + Ex1_2;
    const double Cx = Ex1_0 * Ex1_0 + Ex1_1 * Ex1_1;
    const double Cy = Ex2_0 * Ex2_0 + Ex2_1 * Ex2_1;
    return C * C / (Cx + Cy);
}

// x.homogeneous().normalized() of the reference: (x, y, 1)/sqrt(x^2+y^2+1), per point (not per
// hypothesis), so the prepare kernel computes it once
RP_HD V3 bearing(double x, double y) {
    const double n = sqrt(x * x + y * y + 1.0);
    return v3(x / n, y / n, 1.0 / n);
}

// check_cheirality so@0x1dce00 with min_depth = 0.01 on unit bearings
RP_HD bool cheirality_exact(Quat q, V3 t, V3 b1, V3 b2) {
    const V3 Rb1 = quat_rotate(q, b1);
    const double a = -(Rb1.x * b2.x + Rb1.y * b2.y + Rb1.z * b2.z);
    const double be1 = -(Rb1.x * t.x + Rb1.y * t.y + Rb1.z * t.z);
    const double be2 = b2.x * t.x + b2.y * t.y + b2.z * t.z;
    const double lambda1 = be1 - a * be2;
    const double lambda2 = -a * be1 + be2;
    const double min_depth = 0.01 * (1.0 - a * a);
    return lambda1 > min_depth && lambda2 > min_depth;
}

// ---- tier 0: FP32 "certainly an outlier" filter -------------------------------------------
// With u = 2^-24, E~ = fl32(E), x~ = fl32(x), M = max_k (|x1|_1+1)(|x2|_1+1), m = max_k max(|x1|_1+1,
// |x2|_1+1) over the pair's points and Emax = max|E_ij|:
//   |C~ - C|            <= 7.2 u Emax M          (two input roundings + two FMA roundings per term)
//   sqrt(den) <= sqrt(den~)(1+2.1u) + 8.2 u Emax m
// so   (|C~| - eps)_+^2 > g * den~   with  eps = 16 u Emax (M + thr m),  g = thr^2 (1+1e-5)
// implies C^2/den > thr^2 (1+1e-6): the FP64 reference would also say "outlier".  The constants
// carry >2x slack over the derived bounds; hypotheses whose Emax / thr leave the range where FP32
// products stay normal get eps = +inf (filter disabled, everything goes to tier 1).
struct alignas(16) Filter32 {
    float e00, e01, e02, e10, e11, e12, e20, e21, e22;
    float eps, g, pad;
};

RP_HD Filter32 make_filter32(const M3 &E, double thr, double Mmax, double mmax) {
    Filter32 f;
    f.e00 = (float)E.r0.x; f.e01 = (float)E.r0.y; f.e02 = (float)E.r0.z;
    f.e10 = (float)E.r1.x; f.e11 = (float)E.r1.y; f.e12 = (float)E.r1.z;
    f.e20 = (float)E.r2.x; f.e21 = (float)E.r2.y; f.e22 = (float)E.r2.z;
    double emax = fmax(fmax(fmax(fabs(E.r0.x), fabs(E.r0.y)), fmax(fabs(E.r0.z), fabs(E.r1.x))),
                       fmax(fmax(fabs(E.r1.y), fabs(E.r1.z)), fmax(fmax(fabs(E.r2.x), fabs(E.r2.y)), fabs(E.r2.z))));
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double eps = 16.0 * u * emax * (Mmax + thr * mmax);
    const bool sane = emax > 1e-10 && emax < 1e10 && thr > 1e-10 && thr < 1e3 && Mmax < 1e6 && eps == eps;
    f.eps = sane ? (float)(eps * 1.0001) : INFINITY;
    f.g = (float)(thr * thr * (1.0 + 1e-5));
    f.pad = 0.f;
    return f;
}

// true  => the FP64 reference test r2 < thr^2 is certainly false for this point
RP_HD bool certain_outlier32(const Filter32 &f, float x1_0, float x1_1, float x2_0, float x2_1) {
    const float a0 = fmaf_(f.e00, x1_0, fmaf_(f.e01, x1_1, f.e02));
    const float a1 = fmaf_(f.e10, x1_0, fmaf_(f.e11, x1_1, f.e12));
    const float a2 = fmaf_(f.e20, x1_0, fmaf_(f.e21, x1_1, f.e22));
    const float b0 = fmaf_(f.e00, x2_0, fmaf_(f.e10, x2_1, f.e20));
    const float b1 = fmaf_(f.e01, x2_0, fmaf_(f.e11, x2_1, f.e21));
    const float C = fmaf_(x2_0, a0, fmaf_(x2_1, a1, a2));
    const float den = fmaf_(a0, a0, fmaf_(a1, a1, fmaf_(b0, b0, b1 * b1)));
    const float tt = fmaxf(fabsf(C) - f.eps, 0.0f);
    return tt * tt > f.g * den;
}


// ---- FP32 cheirality screen ------------------------------------------------------------------
// check_cheirality so@0x1dce00 evaluated in FP32 with R(q) as a matrix and bearings recomputed as
// (x,y,1)*rsqrt(x^2+y^2+1).  With u = 2^-24, |b| <= 1, |R_ij| <= 1, T = |t|_1:
//   bearings carry <= 8u relative error (input rounding, sum, rsqrt <= 2 ulp, product),
//   |d(R b1)_i| <= 12u,  |da| <= 38u,  |d beta1| <= 16uT,  |d beta2| <= 12uT,
//   |d lambda_i| <= 70uT,  |d(0.01(1-a^2))| <= 0.8u.
// The screen answers +1 / -1 only when min(lambda1,lambda2) - 0.01(1-a^2) clears
// mu = 128uT + 2u (again ~2x slack); 0 sends the point to the exact FP64 test.
struct alignas(16) Cheir32 {
    float r00, r01, r02, r10, r11, r12, r20, r21, r22;
    float tx, ty, tz, mu, pad0, pad1, pad2;
};

RP_HD Cheir32 make_cheir32(Quat q, V3 t) {
    const M3 R = quat_to_rotmat(q);
    Cheir32 c;
    c.r00 = (float)R.r0.x; c.r01 = (float)R.r0.y; c.r02 = (float)R.r0.z;
    c.r10 = (float)R.r1.x; c.r11 = (float)R.r1.y; c.r12 = (float)R.r1.z;
    c.r20 = (float)R.r2.x; c.r21 = (float)R.r2.y; c.r22 = (float)R.r2.z;
    c.tx = (float)t.x; c.ty = (float)t.y; c.tz = (float)t.z;
    const double T = fabs(t.x) + fabs(t.y) + fabs(t.z);
    const double u = 5.9604644775390625e-08;
    const double qn = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    const bool sane = fabs(qn - 1.0) < 1e-6 && T > 1e-12 && T < 1e12;
    c.mu = sane ? (float)((128.0 * u * T + 2.0 * u) * 1.0001) : INFINITY;
    c.pad0 = c.pad1 = c.pad2 = 0.f;
    return c;
}

RP_HD float rsqrt32(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// +1: the FP64 reference test certainly passes, -1: certainly fails, 0: undecided
RP_HD int cheirality32(const Cheir32 &c, float x1_0, float x1_1, float x2_0, float x2_1) {
    const float n1 = rsqrt32(fmaf_(x1_0, x1_0, fmaf_(x1_1, x1_1, 1.0f)));
    const float n2 = rsqrt32(fmaf_(x2_0, x2_0, fmaf_(x2_1, x2_1, 1.0f)));
    const float b1x = x1_0 * n1, b1y = x1_1 * n1, b1z = n1;
    const float b2x = x2_0 * n2, b2y = x2_1 * n2, b2z = n2;
    const float rx = fmaf_(c.r00, b1x, fmaf_(c.r01, b1y, c.r02 * b1z));
    const float ry = fmaf_(c.r10, b1x, fmaf_(c.r11, b1y, c.r12 * b1z));
    const float rz = fmaf_(c.r20, b1x, fmaf_(c.r21, b1y, c.r22 * b1z));
    const float a = -fmaf_(rx, b2x, fmaf_(ry, b2y, rz * b2z));
    const float be1 = -fmaf_(rx, c.tx, fmaf_(ry, c.ty, rz * c.tz));
    const float be2 = fmaf_(b2x, c.tx, fmaf_(b2y, c.ty, b2z * c.tz));
    const float l1 = fmaf_(-a, be2, be1);
    const float l2 = fmaf_(-a, be1, be2);
    const float md = 0.01f * fmaf_(-a, a, 1.0f);
    const float m = fminf(l1, l2) - md;
    return m > c.mu ? 1 : (m < -c.mu ? -1 : 0);
}

}  // namespace rp
