// rp_common.cuh — shared types and small FP64 algebra for the RePoseD kernels.
//
// Everything here is `RP_HD` (host+device) so tests/hostcheck can compile the very same
// math with g++ and compare it with the oracle on the CPU-only build box.  The product
// (librepose_b200.so) never executes it on the host.
//
// Numerics contract: the translation unit is compiled with -fmad=false (host: -ffp-contract=off),
// so `a*b+c` is two IEEE roundings exactly like the reference's SSE2 build; fused multiply-adds
// appear only where written explicitly (rp::fma_) on paths that need speed, not bit parity.
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>
#include "../../include/repose_b200.h"

#if defined(__CUDACC__)
#define RP_HD __host__ __device__ __forceinline__
#define RP_D __device__ __forceinline__
#else
#define RP_HD inline
#define RP_D inline
#endif

namespace rp {

RP_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
RP_HD float fmaf_(float a, float b, float c) { return ::fmaf(a, b, c); }

struct V3 {
    double x, y, z;
};
RP_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RP_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RP_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RP_HD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RP_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
RP_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RP_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// 3x3, row-major, kept as named rows so everything stays in registers
struct M3 {
    V3 r0, r1, r2;
};
RP_HD V3 mul(const M3 &M, V3 v) { return v3(dot(M.r0, v), dot(M.r1, v), dot(M.r2, v)); }
RP_HD V3 mulT(const M3 &M, V3 v) {
    return v3(M.r0.x * v.x + M.r1.x * v.y + M.r2.x * v.z, M.r0.y * v.x + M.r1.y * v.y + M.r2.y * v.z,
              M.r0.z * v.x + M.r1.z * v.y + M.r2.z * v.z);
}
RP_HD V3 col0(const M3 &M) { return v3(M.r0.x, M.r1.x, M.r2.x); }
RP_HD V3 col1(const M3 &M) { return v3(M.r0.y, M.r1.y, M.r2.y); }
RP_HD V3 col2(const M3 &M) { return v3(M.r0.z, M.r1.z, M.r2.z); }
RP_HD M3 from_cols(V3 a, V3 b, V3 c) {
    M3 M;
    M.r0 = v3(a.x, b.x, c.x);
    M.r1 = v3(a.y, b.y, c.y);
    M.r2 = v3(a.z, b.z, c.z);
    return M;
}
RP_HD M3 matmul(const M3 &A, const M3 &B) {
    const V3 c0 = col0(B), c1 = col1(B), c2 = col2(B);
    M3 O;
    O.r0 = v3(dot(A.r0, c0), dot(A.r0, c1), dot(A.r0, c2));
    O.r1 = v3(dot(A.r1, c0), dot(A.r1, c1), dot(A.r1, c2));
    O.r2 = v3(dot(A.r2, c0), dot(A.r2, c1), dot(A.r2, c2));
    return O;
}
// cofactor inverse; same evaluation order as the oracle's inv3 so the solvers agree to the last bit
RP_HD M3 inverse(const M3 &M) {
    const double c00 = M.r1.y * M.r2.z - M.r1.z * M.r2.y;
    const double c01 = M.r1.z * M.r2.x - M.r1.x * M.r2.z;
    const double c02 = M.r1.x * M.r2.y - M.r1.y * M.r2.x;
    const double det = M.r0.x * c00 + M.r0.y * c01 + M.r0.z * c02;
    const double id = 1.0 / det;
    M3 O;
    O.r0 = v3(c00 * id, (M.r0.z * M.r2.y - M.r0.y * M.r2.z) * id, (M.r0.y * M.r1.z - M.r0.z * M.r1.y) * id);
    O.r1 = v3(c01 * id, (M.r0.x * M.r2.z - M.r0.z * M.r2.x) * id, (M.r0.z * M.r1.x - M.r0.x * M.r1.z) * id);
    O.r2 = v3(c02 * id, (M.r0.y * M.r2.x - M.r0.x * M.r2.y) * id, (M.r0.x * M.r1.y - M.r0.y * M.r1.x) * id);
    return O;
}

struct Quat {
    double w, x, y, z;
};

// Eigen::Quaterniond(w,x,y,z).toRotationMatrix() — the reference's CameraPose::R()
RP_HD M3 quat_to_rotmat(Quat q) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 R;
    R.r0 = v3(1.0 - (tyy + tzz), txy - twz, txz + twy);
    R.r1 = v3(txy + twz, 1.0 - (txx + tzz), tyz - twx);
    R.r2 = v3(txz - twy, tyz + twx, 1.0 - (txx + tyy));
    return R;
}

// Eigen::Quaterniond(R) followed by normalisation (PoseLib rotmat_to_quat)
RP_HD Quat rotmat_to_quat(const M3 &R) {
    const double m00 = R.r0.x, m11 = R.r1.y, m22 = R.r2.z;
    const double tr = m00 + m11 + m22;
    double w, x, y, z;
    if (tr > 0.0) {
        double t = sqrt(tr + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (R.r2.y - R.r1.z) * t;
        y = (R.r0.z - R.r2.x) * t;
        z = (R.r1.x - R.r0.y) * t;
    } else {
        int i = 0;
        if (m11 > m00) i = 1;
        if (m22 > (i == 0 ? m00 : m11)) i = 2;
        if (i == 0) {
            double t = sqrt(m00 - m11 - m22 + 1.0);
            x = 0.5 * t;
            t = 0.5 / t;
            w = (R.r2.y - R.r1.z) * t;
            y = (R.r1.x + R.r0.y) * t;
            z = (R.r2.x + R.r0.z) * t;
        } else if (i == 1) {
            double t = sqrt(m11 - m22 - m00 + 1.0);
            y = 0.5 * t;
            t = 0.5 / t;
            w = (R.r0.z - R.r2.x) * t;
            z = (R.r2.y + R.r1.z) * t;
            x = (R.r0.y + R.r1.x) * t;
        } else {
            double t = sqrt(m22 - m00 - m11 + 1.0);
            z = 0.5 * t;
            t = 0.5 / t;
            w = (R.r1.x - R.r0.y) * t;
            x = (R.r0.z + R.r2.x) * t;
            y = (R.r1.z + R.r2.y) * t;
        }
    }
    const double n = sqrt(w * w + x * x + y * y + z * z);
    Quat q;
    q.w = w / n; q.x = x / n; q.y = y / n; q.z = z / n;
    return q;
}

// PoseLib quat_rotate (misc/quaternion.h): R(q) p through the quaternion sandwich, in the
// reference's operation order (check_cheirality depends on it bit for bit)
RP_HD V3 quat_rotate(Quat q, V3 p) {
    const double q1 = q.w, q2 = q.x, q3 = q.y, q4 = q.z;
    const double p1 = p.x, p2 = p.y, p3 = p.z;
    const double px1 = -p1 * q2 - p2 * q3 - p3 * q4;
    const double px2 = p1 * q1 - p2 * q4 + p3 * q3;
    const double px3 = p2 * q1 + p1 * q4 - p3 * q2;
    const double px4 = p2 * q2 - p1 * q3 + p3 * q1;
    return v3(px2 * q1 - px1 * q2 - px3 * q4 + px4 * q3, px3 * q1 - px1 * q3 + px2 * q4 - px4 * q2,
              px3 * q2 - px2 * q3 - px1 * q4 + px4 * q1);
}

RP_HD Quat quat_mul(Quat a, Quat b) {
    Quat o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}

// quat_exp so@0x262fd0: rotation vector -> unit quaternion
RP_HD Quat quat_exp(V3 w) {
    const double th2 = w.x * w.x + w.y * w.y + w.z * w.z;
    const double th = sqrt(th2);
    const double a = 0.5 * th;
    double re, im;
    if (th > 1e-6) {
        re = cos(a);
        im = sin(a) / th;
    } else {
        const double a2 = th2 * 0.25;
        re = 1.0 - a2 / 2.0 + a2 * a2 / 24.0;
        im = 0.5 - a2 / 12.0 + a2 * a2 / 240.0;
    }
    Quat q;
    q.w = re; q.x = im * w.x; q.y = im * w.y; q.z = im * w.z;
    return q;
}

// device-side model, identical layout to rp_model (12 doubles)
struct Model {
    Quat q;
    V3 t;
    double scale, shift1, shift2, f1, f2;
};
static_assert(sizeof(Model) == sizeof(rp_model), "Model must match the C ABI layout");

RP_HD Model identity_model() {
    Model m;
    m.q.w = 1.0; m.q.x = m.q.y = m.q.z = 0.0;
    m.t = v3(0.0, 0.0, 0.0);
    m.scale = 1.0; m.shift1 = m.shift2 = 0.0; m.f1 = m.f2 = 1.0;
    return m;
}

// essential_from_motion so@0x1dcb60: E = [t]x R(q) evaluated as Eigen's 3x3 product
// (sum over k = 0,1,2 including the structural zeros of [t]x)
RP_HD M3 essential_from_motion(Quat q, V3 t) {
    const M3 R = quat_to_rotmat(q);
    // the structural zeros of [t]x only add +-0 terms, which never change a finite sum
    M3 E;
    E.r0 = v3((-t.z) * R.r1.x + t.y * R.r2.x, (-t.z) * R.r1.y + t.y * R.r2.y, (-t.z) * R.r1.z + t.y * R.r2.z);
    E.r1 = v3(t.z * R.r0.x + (-t.x) * R.r2.x, t.z * R.r0.y + (-t.x) * R.r2.y, t.z * R.r0.z + (-t.x) * R.r2.z);
    E.r2 = v3((-t.y) * R.r0.x + t.x * R.r1.x, (-t.y) * R.r0.y + t.x * R.r1.y, (-t.y) * R.r0.z + t.x * R.r1.z);
    return E;
}

// F = diag(1,1,f2) E diag(1,1,f1): focal estimators' score_model so@0x4fac60 / so@0x4faf90
RP_HD M3 fundamental_from_model(const Model &m) {
    M3 F = essential_from_motion(m.q, m.t);
    F.r0.z = F.r0.z * m.f1;
    F.r1.z = F.r1.z * m.f1;
    F.r2.x = m.f2 * F.r2.x;
    F.r2.y = m.f2 * F.r2.y;
    F.r2.z = (m.f2 * F.r2.z) * m.f1;
    return F;
}

RP_HD int n_params(int variant) { return variant == RP_CALIB ? 7 : (variant == RP_SHARED ? 8 : 9); }

}  // namespace rp
