// rp_solvers.cuh — depth-aware minimal solvers, one thread per RANSAC hypothesis.
//
//   S1  calibrated, scale only     P3P (Ding et al. CVPR'23, PoseLib p3p so@0xecd50) + scale,
//                                  RelativePoseMonoDepthEstimator::generate_models so@0x4fe090
//   S2  calibrated, scale+shifts   relpose_monodepth_3pt so@0x155ca0
//   S3  shared unknown focal       relpose_monodepth_3pt_shared_focal so@0x18fdf0
//   S4  two unknown focals         relpose_monodepth_3pt_varying_focal so@0x19bcd0
//
// FP64 closed forms only (cubic/quartic roots, 3x3 inverses, triangle alignment); no dynamic
// indexing, so each hypothesis lives in registers.  SURVEY.md §8a rows S1-S4.
#pragma once
#include "rp_common.cuh"

namespace rp {

// ---- univariate.h -----------------------------------------------------------------------
// solve_cubic_single_real so@0x1dabf0: one real root of x^3 + c2 x^2 + c1 x + c0.
// Returns true when that root is the only real one.
RP_HD bool solve_cubic_single_real(double c2, double c1, double c0, double &root) {
    const double a = c1 - c2 * c2 / 3.0;
    double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
    double c = b * b / 4.0 + a * a * a / 27.0;
    if (c != 0) {
        if (c > 0) {
            c = sqrt(c);
            b *= -0.5;
            root = cbrt(b + c) + cbrt(b - c) - c2 / 3.0;
            return true;
        }
        c = 3.0 * b / (2.0 * a) * sqrt(-3.0 / a);
        root = 2.0 * sqrt(-a / 3.0) * cos(acos(c) / 3.0) - c2 / 3.0;
    } else {
        root = -c2 / 3.0 + (a != 0 ? (3.0 * b / a) : 0);
    }
    return false;
}

RP_HD double sign1(double x) { return x < 0 ? -1.0 : 1.0; }

// solve_quartic_real so@0x1dc570: real roots of x^4 + b x^3 + c x^2 + d x + e via the resolvent
// cubic and two quadratics, one Newton step each.  roots r0..r3, returns how many.
RP_HD int solve_quartic_real(double b, double c, double d, double e, double &r0, double &r1, double &r2,
                             double &r3) {
    const double p = c - 3.0 * b * b / 8.0;
    const double q = b * b * b / 8.0 - 0.5 * b * c + d;
    const double r = (-3.0 * b * b * b * b + 256.0 * e - 64.0 * b * d + 16.0 * b * b * c) / 256.0;
    double u2;
    solve_cubic_single_real(2.0 * p, p * p - 4.0 * r, -q * q, u2);
    if (u2 < 0) return 0;
    const double u = sqrt(u2);
    const double s = -u;
    const double t = (p + u * u + q / u) / 2.0;
    const double v = (p + u * u - q / u) / 2.0;
    int sols = 0;
    double disc = u * u - 4.0 * v;
    if (disc > 0) {
        r0 = (-u - sign1(u) * sqrt(disc)) / 2.0;
        r1 = v / r0;
        sols = 2;
    }
    disc = s * s - 4.0 * t;
    if (disc > 0) {
        const double ra = (-s - sign1(s) * sqrt(disc)) / 2.0;
        const double rb = t / ra;
        if (sols == 0) { r0 = ra; r1 = rb; } else { r2 = ra; r3 = rb; }
        sols += 2;
    }
#define RP_NEWTON(x)                                                                                  \
    do {                                                                                              \
        x = x - b / 4.0;                                                                              \
        const double x2_ = x * x, x3_ = x * x2_;                                                      \
        x = x + (-(x2_ * x2_ + b * x3_ + c * x2_ + d * x + e) / (4.0 * x3_ + 3.0 * b * x2_ + 2.0 * c * x + d)); \
    } while (0)
    if (sols >= 2) { RP_NEWTON(r0); RP_NEWTON(r1); }
    if (sols == 4) { RP_NEWTON(r2); RP_NEWTON(r3); }
#undef RP_NEWTON
    return sols;
}

// root2real of p3p.cc: real roots of x^2 + b x + c
RP_HD bool root2real(double b, double c, double &r1, double &r2) {
    const double THRESHOLD = -1.0e-12;
    const double v = b * b - 4.0 * c;
    if (v < THRESHOLD) {
        r1 = r2 = -0.5 * b;
        return v >= 0;
    }
    if (v > THRESHOLD && v < 0.0) {
        r1 = -0.5 * b;
        r2 = -2;
        return true;
    }
    const double y = sqrt(v);
    if (b < 0) {
        r1 = 0.5 * (-b + y);
        r2 = 0.5 * (-b - y);
    } else {
        r1 = 2.0 * c / (-b + y);
        r2 = 2.0 * c / (-b - y);
    }
    return true;
}

// ---- rigid alignment of two congruent triangles: Y_i = R X_i + t --------------------------
RP_HD void align_triangles(V3 X0, V3 X1, V3 X2, V3 Y0, V3 Y1, V3 Y2, Quat &q, V3 &t) {
    const V3 a1 = X1 - X0, b1 = X2 - X0, a2 = Y1 - Y0, b2 = Y2 - Y0;
    const M3 M1 = from_cols(a1, b1, cross(a1, b1));
    const M3 M2 = from_cols(a2, b2, cross(a2, b2));
    const M3 R = matmul(M2, inverse(M1));
    q = rotmat_to_quat(R);
    t = Y0 - mul(quat_to_rotmat(q), X0);
}

// ---- S1: P3P -------------------------------------------------------------------------------
struct P3PSolutions {
    Quat q[4];
    V3 t[4];
    int n;
};

RP_HD void p3p_refine_lambda(double &l1, double &l2, double &l3, double a12, double a13, double a23, double b12,
                             double b13, double b23) {
    for (int iter = 0; iter < 5; ++iter) {
        const double r1 = (l1 * l1 - 2.0 * l1 * l2 * b12 + l2 * l2 - a12);
        const double r2 = (l1 * l1 - 2.0 * l1 * l3 * b13 + l3 * l3 - a13);
        const double r3 = (l2 * l2 - 2.0 * l2 * l3 * b23 + l3 * l3 - a23);
        if (fabs(r1) + fabs(r2) + fabs(r3) < 1e-10) return;
        const double x11 = l1 - l2 * b12, x12 = l2 - l1 * b12;
        const double x21 = l1 - l3 * b13, x23 = l3 - l1 * b13;
        const double x32 = l2 - l3 * b23, x33 = l3 - l2 * b23;
        const double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
        l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
        l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
        l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
    }
}

// The solver is split like the scale+shift one (see ShiftCand below): p3p_setup + p3p_candidates are the
// uniform per-sample part, p3p_finish the per-solution part (Gauss-Newton on the three depths + pose assembly).
struct P3PSetup {
    V3 x0, x1, x2;      // bearings after relabelling (side (1,2) is the longest)
    V3 Xa;              // first 3-D point after relabelling
    V3 X01, X02;
    double a01, a02, a12, m01, m02, m12;
};
struct P3PCand {
    double d0, d1, d2;  // depths along the bearings before the refinement
};

RP_HD P3PSetup p3p_setup(V3 x0, V3 x1, V3 x2, V3 Xa, V3 Xb, V3 Xc) {
    V3 X01 = Xa - Xb, X02 = Xa - Xc, X12 = Xb - Xc;
    double a01 = dot(X01, X01), a02 = dot(X02, X02), a12 = dot(X12, X12);
    // relabel so that side (1,2) is the longest
    if (a01 > a02) {
        if (a01 > a12) {
            V3 tv = x0; x0 = x2; x2 = tv;
            tv = Xa; Xa = Xc; Xc = tv;
            double td = a01; a01 = a12; a12 = td;
            X01 = neg(X12);
            X02 = neg(X02);
        }
    } else if (a02 > a12) {
        V3 tv = x0; x0 = x1; x1 = tv;
        tv = Xa; Xa = Xb; Xb = tv;
        double td = a02; a02 = a12; a12 = td;
        X01 = neg(X01);
        X02 = X12;
    }
    P3PSetup S;
    S.x0 = x0; S.x1 = x1; S.x2 = x2; S.Xa = Xa; S.X01 = X01; S.X02 = X02;
    S.a01 = a01; S.a02 = a02; S.a12 = a12;
    S.m01 = dot(x0, x1); S.m02 = dot(x0, x2); S.m12 = dot(x1, x2);
    return S;
}

RP_HD void p3p_push(int &n, P3PCand &c0, P3PCand &c1, P3PCand &c2, P3PCand &c3, double d0, double d1, double d2) {
    P3PCand c;
    c.d0 = d0; c.d1 = d1; c.d2 = d2;
    // static indexing keeps the candidates in registers
    if (n == 0) c0 = c;
    else if (n == 1) c1 = c;
    else if (n == 2) c2 = c;
    else c3 = c;
    ++n;
}

// depth triples of the (at most 4) solutions, in the order p3p.cc emits them
RP_HD int p3p_candidates(const P3PSetup &S, P3PCand &c0, P3PCand &c1, P3PCand &c2, P3PCand &c3) {
    const double a01 = S.a01, a02 = S.a02, a12 = S.a12, m01 = S.m01, m02 = S.m02, m12 = S.m12;
    const double a12d = 1.0 / a12;
    const double a = a01 * a12d, b = a02 * a12d;
    const double m12sq = -m12 * m12 + 1.0, m02sq = -1.0 + m02 * m02, m01sq = -1.0 + m01 * m01;
    const double ab = a * b, bsq = b * b, asq = a * a;
    const double m013 = -2.0 + 2.0 * m01 * m02 * m12;
    const double bsqm12sq = bsq * m12sq, asqm12sq = asq * m12sq, abm12sq = 2.0 * ab * m12sq;
    const double k3_inv = 1.0 / (bsqm12sq + b * m02sq);
    const double k2 = k3_inv * ((-1.0 + a) * m02sq + abm12sq + bsqm12sq + b * m013);
    const double k1 = k3_inv * (asqm12sq + abm12sq + a * m013 + (-1.0 + b) * m01sq);
    const double k0 = k3_inv * (asqm12sq + a * m01sq);
    double s;
    const bool G = solve_cubic_single_real(k2, k1, k0, s);

    // degenerate conic C = D1 + s D2 and its two lines p, q (compute_pq of p3p.cc)
    double C00 = -a + s * (1 - b), C01 = -m02 * s, C02 = a * m12 + b * m12 * s;
    double C11 = s + 1, C12 = -m01, C22 = -a - b * s + 1;
    double C10 = C01, C20 = C02, C21 = C12;
    const double A00 = C12 * C21 - C11 * C22, A11 = C02 * C20 - C00 * C22, A22 = C01 * C10 - C00 * C11;
    const double A01 = C01 * C22 - C02 * C21, A02 = C02 * C11 - C01 * C12, A12 = C00 * C12 - C02 * C10;
    double v0, v1, v2;
    if (A00 > A11 ? (A00 > A22) : false) {
        const double sq = sqrt(A00);
        v0 = A00 / sq; v1 = A01 / sq; v2 = A02 / sq;
    } else if (A00 > A11 ? false : (A11 > A22)) {
        const double sq = sqrt(A11);
        v0 = A01 / sq; v1 = A11 / sq; v2 = A12 / sq;
    } else {
        const double sq = sqrt(A22);
        v0 = A02 / sq; v1 = A12 / sq; v2 = A22 / sq;
    }
    C01 -= v2; C02 += v1; C12 -= v0;
    C10 += v2; C20 -= v1; C21 += v0;

    int n = 0;
    for (int i = 0; i < 2; ++i) {
        const double p0 = C00, p1 = (i == 0) ? C10 : C01, p2 = (i == 0) ? C20 : C02;
        const bool switch_12 = fabs(p0) <= fabs(p1);
        double tau0, tau1;
        if (switch_12) {
            const double w0 = -p0 / p1, w1 = -p2 / p1;
            const double ca = 1.0 / (w1 * w1 - b);
            const double cb = 2.0 * (b * m12 - m02 * w1 + w0 * w1) * ca;
            const double cc = (w0 * w0 - 2 * m02 * w0 - b + 1.0) * ca;
            if (!root2real(cb, cc, tau0, tau1)) continue;
            for (int k = 0; k < 2; ++k) {
                const double tau = k == 0 ? tau0 : tau1;
                if (tau <= 0) continue;
                double d2 = sqrt(a12 / (tau * (tau - 2.0 * m12) + 1.0));
                double d1 = tau * d2;
                double d0 = (w0 * d2 + w1 * d1);
                if (d0 < 0) continue;
                p3p_push(n, c0, c1, c2, c3, d0, d1, d2);
                if (n == 4) return n;
            }
        } else {
            const double w0 = -p1 / p0, w1 = -p2 / p0;
            const double ca = 1.0 / (-a * w1 * w1 + 2 * a * m12 * w1 - a + 1);
            const double cb = 2 * (a * m12 * w0 - m01 - a * w0 * w1) * ca;
            const double cc = (1 - a * w0 * w0) * ca;
            if (!root2real(cb, cc, tau0, tau1)) continue;
            for (int k = 0; k < 2; ++k) {
                const double tau = k == 0 ? tau0 : tau1;
                if (tau <= 0) continue;
                double d0 = sqrt(a01 / (tau * (tau - 2.0 * m01) + 1.0));
                double d1 = tau * d0;
                double d2 = w0 * d0 + w1 * d1;
                if (d2 < 0) continue;
                p3p_push(n, c0, c1, c2, c3, d0, d1, d2);
                if (n == 4) return n;
            }
        }
        if (n > 0 && G) break;
    }
    return n;
}

// refinement of the depths and pose assembly: d_i x_i = R X_i + t
RP_HD void p3p_finish(const P3PSetup &S, const P3PCand &c, Quat &q, V3 &t) {
    double d0 = c.d0, d1 = c.d1, d2 = c.d2;
    p3p_refine_lambda(d0, d1, d2, S.a01, S.a02, S.a12, S.m01, S.m02, S.m12);
    const M3 XX = inverse(from_cols(S.X01, S.X02, cross(S.X01, S.X02)));
    const V3 v1 = d0 * S.x0 - d1 * S.x1;
    const V3 v2 = d0 * S.x0 - d2 * S.x2;
    const M3 YY = from_cols(v1, v2, cross(v1, v2));
    const M3 R = matmul(YY, XX);
    q = rotmat_to_quat(R);
    t = d0 * S.x0 - mul(R, S.Xa);
}

// x: unit bearings in camera 2, X: 3-D points in camera 1.  Finds R,t with d_i x_i = R X_i + t.
RP_HD void p3p(V3 x0, V3 x1, V3 x2, V3 Xa, V3 Xb, V3 Xc, P3PSolutions &out) {
    out.n = 0;
    const P3PSetup S = p3p_setup(x0, x1, x2, Xa, Xb, Xc);
    P3PCand c0, c1, c2, c3;
    const int n = p3p_candidates(S, c0, c1, c2, c3);
    if (n > 0) { p3p_finish(S, c0, out.q[0], out.t[0]); }
    if (n > 1) { p3p_finish(S, c1, out.q[1], out.t[1]); }
    if (n > 2) { p3p_finish(S, c2, out.q[2], out.t[2]); }
    if (n > 3) { p3p_finish(S, c3, out.q[3], out.t[3]); }
    out.n = n;
}

// one minimal sample: three correspondences as homogeneous (x, y, 1) points plus depths
struct Triplet {
    V3 p1[3], p2[3];
    double d1[3], d2[3];
};

struct ModelSet {
    Model m[4];
    int n;
};

RP_HD void set_model(ModelSet &out, const Model &m) {
    const int k = out.n;
    if (k == 0) out.m[0] = m;
    else if (k == 1) out.m[1] = m;
    else if (k == 2) out.m[2] = m;
    else out.m[3] = m;
    out.n = k + 1;
}

// S1: X_i = d1_i (x1_i, 1), bearings b_i = (x2_i, 1)/|.|, P3P, then
// scale = (R X_0 + t).x / (d2_0 x2_0.x) from the first sampled point; shifts stay 0.
RP_HD void calib_scale_inputs(const Triplet &s, V3 (&X)[3], V3 (&b)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V3 h = s.p2[i];
        const double n = sqrt(h.x * h.x + h.y * h.y + h.z * h.z);
        X[i] = v3(s.d1[i] * s.p1[i].x, s.d1[i] * s.p1[i].y, s.d1[i] * s.p1[i].z);
        b[i] = v3(h.x / n, h.y / n, h.z / n);
    }
}
RP_HD int solve_roots(const Triplet &s, P3PCand &c0, P3PCand &c1, P3PCand &c2, P3PCand &c3) {
    V3 X[3], b[3];
    calib_scale_inputs(s, X, b);
    return p3p_candidates(p3p_setup(b[0], b[1], b[2], X[0], X[1], X[2]), c0, c1, c2, c3);
}
RP_HD Model solve_finish(const Triplet &s, const P3PCand &c) {
    V3 X[3], b[3];
    calib_scale_inputs(s, X, b);
    Model m = identity_model();
    p3p_finish(p3p_setup(b[0], b[1], b[2], X[0], X[1], X[2]), c, m.q, m.t);
    const M3 R = quat_to_rotmat(m.q);
    m.scale = (dot(R.r0, X[0]) + m.t.x) / (s.d2[0] * s.p2[0].x);
    return m;
}
RP_HD void solve_calib_scale(const Triplet &s, ModelSet &out) {
    out.n = 0;
    P3PCand c0, c1, c2, c3;
    const int n = solve_roots(s, c0, c1, c2, c3);
    if (n > 0) set_model(out, solve_finish(s, c0));
    if (n > 1) set_model(out, solve_finish(s, c1));
    if (n > 2) set_model(out, solve_finish(s, c2));
    if (n > 3) set_model(out, solve_finish(s, c3));
}

// S2: unknowns s = scale^2, u = shift1, v = shift2.  For the pairs (0,1),(0,2),(1,2):
//   s |(d2_i+v) p2_i - (d2_j+v) p2_j|^2 = |(d1_i+u) p1_i - (d1_j+u) p1_j|^2
//   <=> c0 s v^2 + c1 u^2 + c2 s v + c3 s + c4 u + c5 = 0.
// Eliminating (s v^2, s v, s) linearly leaves quadratics in u; (s v)^2 = (s v^2) s is a quartic
// in u.  Raw roots are filtered (s>0, all six shifted depths >0) BEFORE the five Gauss-Newton
// polish steps (refine_suv so@0x15de40, tolerance 1e-10 on sum |r|), as in the binary.
struct ShiftEq {
    double c0, c1, c2, c3, c4, c5;
};
RP_HD ShiftEq shift_equation(V3 p1i, V3 p1j, V3 p2i, V3 p2j, double ai, double aj, double bi, double bj) {
    const double n1i = dot(p1i, p1i), n1j = dot(p1j, p1j), c1 = dot(p1i, p1j);
    const double n2i = dot(p2i, p2i), n2j = dot(p2j, p2j), c2 = dot(p2i, p2j);
    ShiftEq e;
    e.c0 = n2i + n2j - 2.0 * c2;
    e.c1 = -(n1i + n1j - 2.0 * c1);
    e.c2 = 2.0 * (bi * n2i + bj * n2j - c2 * (bi + bj));
    e.c3 = bi * bi * n2i + bj * bj * n2j - 2.0 * bi * bj * c2;
    e.c4 = -2.0 * (ai * n1i + aj * n1j - c1 * (ai + aj));
    e.c5 = -(ai * ai * n1i + aj * aj * n1j - 2.0 * ai * aj * c1);
    return e;
}

// The solver is split in two so that the RANSAC kernel can run the cheap, uniform part (equations, quartic,
// root filter) with one thread per sample and the expensive per-root part (Newton polish + triangle
// alignment) with one thread per surviving root (solve2_kernel): a sample has 0..4 roots, and a
// thread-per-sample loop over them leaves most lanes idle.
struct ShiftSystem {
    ShiftEq e0, e1, e2;
};
struct ShiftCand {
    double u, s, v;  // shift1, scale^2, shift2 before the polish
};

RP_HD ShiftSystem shift_system(const Triplet &t) {
    ShiftSystem S;
    S.e0 = shift_equation(t.p1[0], t.p1[1], t.p2[0], t.p2[1], t.d1[0], t.d1[1], t.d2[0], t.d2[1]);
    S.e1 = shift_equation(t.p1[0], t.p1[2], t.p2[0], t.p2[2], t.d1[0], t.d1[2], t.d2[0], t.d2[2]);
    S.e2 = shift_equation(t.p1[1], t.p1[2], t.p2[1], t.p2[2], t.d1[1], t.d1[2], t.d2[1], t.d2[2]);
    return S;
}

// roots of the quartic in u = shift1 that pass the sign filters, in root order; returns their number
RP_HD int solve_calib_shift_roots(const Triplet &t, const ShiftSystem &S, ShiftCand &c0, ShiftCand &c1, ShiftCand &c2,
                                  ShiftCand &c3) {
    const ShiftEq &e0 = S.e0, &e1 = S.e1, &e2 = S.e2;
    M3 A;
    A.r0 = v3(e0.c0, e0.c2, e0.c3);
    A.r1 = v3(e1.c0, e1.c2, e1.c3);
    A.r2 = v3(e2.c0, e2.c2, e2.c3);
    const double c00 = A.r1.y * A.r2.z - A.r1.z * A.r2.y;
    const double c01 = A.r1.z * A.r2.x - A.r1.x * A.r2.z;
    const double c02 = A.r1.x * A.r2.y - A.r1.y * A.r2.x;
    if (A.r0.x * c00 + A.r0.y * c01 + A.r0.z * c02 == 0.0) return 0;
    const M3 Ai = inverse(A);
    // rows of P: s v^2, s v, s as quadratics in u (coefficients of u^2, u, 1)
    const V3 q2 = v3(e0.c1, e1.c1, e2.c1), q1 = v3(e0.c4, e1.c4, e2.c4), q0 = v3(e0.c5, e1.c5, e2.c5);
    const double b0 = -dot(Ai.r0, q2), b1 = -dot(Ai.r0, q1), b2 = -dot(Ai.r0, q0);  // s v^2
    const double a0 = -dot(Ai.r1, q2), a1 = -dot(Ai.r1, q1), a2 = -dot(Ai.r1, q0);  // s v
    const double g0 = -dot(Ai.r2, q2), g1 = -dot(Ai.r2, q1), g2 = -dot(Ai.r2, q0);  // s
    const double k4 = a0 * a0 - b0 * g0;
    const double k3 = 2.0 * a0 * a1 - (b0 * g1 + b1 * g0);
    const double k2 = 2.0 * a0 * a2 + a1 * a1 - (b0 * g2 + b1 * g1 + b2 * g0);
    const double k1 = 2.0 * a1 * a2 - (b1 * g2 + b2 * g1);
    const double k0 = a2 * a2 - b2 * g2;
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    const int nr = solve_quartic_real(k3 / k4, k2 / k4, k1 / k4, k0 / k4, r0, r1, r2, r3);
    int n = 0;
#pragma unroll
    for (int ir = 0; ir < 4; ++ir) {
        if (ir >= nr) break;
        const double u = ir == 0 ? r0 : (ir == 1 ? r1 : (ir == 2 ? r2 : r3));
        const double s = (g0 * u + g1) * u + g2;
        const double sv = (a0 * u + a1) * u + a2;
        const double v = sv / s;
        if (!(s > 0)) continue;
        if (!(t.d1[0] + u > 0 && t.d1[1] + u > 0 && t.d1[2] + u > 0)) continue;
        if (!(t.d2[0] + v > 0 && t.d2[1] + v > 0 && t.d2[2] + v > 0)) continue;
        ShiftCand c;
        c.u = u; c.s = s; c.v = v;
        // static indexing keeps the candidates in registers
        if (n == 0) c0 = c;
        else if (n == 1) c1 = c;
        else if (n == 2) c2 = c;
        else c3 = c;
        ++n;
    }
    return n;
}

// Newton polish of (s, u, v) on the three distance equations, then the exact 3-point alignment
RP_HD Model solve_calib_shift_finish(const Triplet &t, const ShiftSystem &S, const ShiftCand &c) {
    const ShiftEq &e0 = S.e0, &e1 = S.e1, &e2 = S.e2;
    double u = c.u, s = c.s, v = c.v;
    for (int it = 0; it < 5; ++it) {
        const double ra = e0.c0 * s * v * v + e0.c1 * u * u + e0.c2 * s * v + e0.c3 * s + e0.c4 * u + e0.c5;
        const double rb = e1.c0 * s * v * v + e1.c1 * u * u + e1.c2 * s * v + e1.c3 * s + e1.c4 * u + e1.c5;
        const double rc = e2.c0 * s * v * v + e2.c1 * u * u + e2.c2 * s * v + e2.c3 * s + e2.c4 * u + e2.c5;
        if (fabs(ra) + fabs(rb) + fabs(rc) < 1e-10) break;
        M3 J;
        J.r0 = v3(e0.c0 * v * v + e0.c2 * v + e0.c3, 2.0 * e0.c1 * u + e0.c4, 2.0 * e0.c0 * s * v + e0.c2 * s);
        J.r1 = v3(e1.c0 * v * v + e1.c2 * v + e1.c3, 2.0 * e1.c1 * u + e1.c4, 2.0 * e1.c0 * s * v + e1.c2 * s);
        J.r2 = v3(e2.c0 * v * v + e2.c2 * v + e2.c3, 2.0 * e2.c1 * u + e2.c4, 2.0 * e2.c0 * s * v + e2.c2 * s);
        const double j00 = J.r1.y * J.r2.z - J.r1.z * J.r2.y;
        const double j01 = J.r1.z * J.r2.x - J.r1.x * J.r2.z;
        const double j02 = J.r1.x * J.r2.y - J.r1.y * J.r2.x;
        if (J.r0.x * j00 + J.r0.y * j01 + J.r0.z * j02 == 0.0) break;
        const V3 dx = mul(inverse(J), v3(ra, rb, rc));
        s -= dx.x; u -= dx.y; v -= dx.z;
    }
    Model m = identity_model();
    m.scale = sqrt(s);
    m.shift1 = u;
    m.shift2 = v;
    V3 X[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a = t.d1[i] + u, bb = t.d2[i] + v;
        X[i] = v3(a * t.p1[i].x, a * t.p1[i].y, a * t.p1[i].z);
        Y[i] = v3(m.scale * bb * t.p2[i].x, m.scale * bb * t.p2[i].y, m.scale * bb * t.p2[i].z);
    }
    align_triangles(X[0], X[1], X[2], Y[0], Y[1], Y[2], m.q, m.t);
    return m;
}

RP_HD int solve_roots(const Triplet &t, ShiftCand &c0, ShiftCand &c1, ShiftCand &c2, ShiftCand &c3) {
    return solve_calib_shift_roots(t, shift_system(t), c0, c1, c2, c3);
}
RP_HD Model solve_finish(const Triplet &t, const ShiftCand &c) { return solve_calib_shift_finish(t, shift_system(t), c); }

RP_HD void solve_calib_shift(const Triplet &t, ModelSet &out) {
    out.n = 0;
    const ShiftSystem S = shift_system(t);
    ShiftCand c0, c1, c2, c3;
    const int n = solve_calib_shift_roots(t, S, c0, c1, c2, c3);
    if (n > 0) set_model(out, solve_calib_shift_finish(t, S, c0));
    if (n > 1) set_model(out, solve_calib_shift_finish(t, S, c1));
    if (n > 2) set_model(out, solve_calib_shift_finish(t, S, c2));
    if (n > 3) set_model(out, solve_calib_shift_finish(t, S, c3));
}

// S4: linear 3x3 in (a = 1/f1^2, b = scale^2, c = scale^2/f2^2); valid iff a,b,c > 0.
RP_HD void solve_varying_focal(const Triplet &t, ModelSet &out) {
    out.n = 0;
    M3 A;
    double rhs[3];
    V3 rows[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = r == 2 ? 1 : 0, j = r == 0 ? 1 : 2;
        const double ax = t.d1[i] * t.p1[i].x - t.d1[j] * t.p1[j].x, ay = t.d1[i] * t.p1[i].y - t.d1[j] * t.p1[j].y;
        const double bx = t.d2[i] * t.p2[i].x - t.d2[j] * t.p2[j].x, by = t.d2[i] * t.p2[i].y - t.d2[j] * t.p2[j].y;
        const double dz1 = t.d1[i] - t.d1[j], dz2 = t.d2[i] - t.d2[j];
        rows[r] = v3(ax * ax + ay * ay, -dz2 * dz2, -(bx * bx + by * by));
        rhs[r] = -dz1 * dz1;
    }
    A.r0 = rows[0]; A.r1 = rows[1]; A.r2 = rows[2];
    const double c00 = A.r1.y * A.r2.z - A.r1.z * A.r2.y;
    const double c01 = A.r1.z * A.r2.x - A.r1.x * A.r2.z;
    const double c02 = A.r1.x * A.r2.y - A.r1.y * A.r2.x;
    if (A.r0.x * c00 + A.r0.y * c01 + A.r0.z * c02 == 0.0) return;
    const V3 sol = mul(inverse(A), v3(rhs[0], rhs[1], rhs[2]));
    if (!(sol.x > 0 && sol.y > 0 && sol.z > 0)) return;
    Model m = identity_model();
    m.f1 = 1.0 / sqrt(sol.x);
    m.scale = sqrt(sol.y);
    m.f2 = sqrt(sol.y / sol.z);
    V3 X[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        X[i] = v3(t.d1[i] * t.p1[i].x / m.f1, t.d1[i] * t.p1[i].y / m.f1, t.d1[i]);
        Y[i] = v3(m.scale * t.d2[i] * t.p2[i].x / m.f2, m.scale * t.d2[i] * t.p2[i].y / m.f2, m.scale * t.d2[i]);
    }
    align_triangles(X[0], X[1], X[2], Y[0], Y[1], Y[2], m.q, m.t);
    set_model(out, m);
}

// S3: unknowns f, scale and nu = (depth of point 2 in camera 2)/scale; points 0,1 align in 3-D,
// point 2 only reprojects.  With g = 1/f^2: S1 = g A1_01 + B1_01, S2 = g A2_01 + B2_01,
// scale^2 = S1/S2; the (0,2)-(1,2) difference is linear in nu: nu = N/(2 S1 L); substituting
// into the (0,2) equation leaves a quintic in g with zero constant term (g = 0 <=> f = inf),
// i.e. a quartic.  Valid iff g > 0 and nu > 0.  Polynomials in g are fixed-degree structs so
// all products unroll into registers.
template <int D>
struct Poly {
    double c[D + 1];
};
template <int DA, int DB>
RP_HD Poly<DA + DB> pmul(const Poly<DA> &a, const Poly<DB> &b) {
    Poly<DA + DB> r;
#pragma unroll
    for (int i = 0; i <= DA + DB; ++i) r.c[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= DA; ++i)
#pragma unroll
        for (int j = 0; j <= DB; ++j) r.c[i + j] += a.c[i] * b.c[j];
    return r;
}
RP_HD Poly<1> plin(double c1, double c0) {
    Poly<1> p;
    p.c[0] = c0;
    p.c[1] = c1;
    return p;
}
template <int D>
RP_HD double peval(const Poly<D> &p, double x) {
    double v = 0;
#pragma unroll
    for (int i = D; i >= 0; --i) v = v * x + p.c[i];
    return v;
}

struct FocalCand {
    double g, sc2, nu;  // 1/f^2, scale^2, depth of point 2 in camera 2 over scale
};

RP_HD int solve_roots(const Triplet &t, FocalCand &c0, FocalCand &c1, FocalCand &c2, FocalCand &c3) {
    double A1[3], B1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = r == 2 ? 1 : 0, j = r == 0 ? 1 : 2;
        const double vx = t.d1[i] * t.p1[i].x - t.d1[j] * t.p1[j].x, vy = t.d1[i] * t.p1[i].y - t.d1[j] * t.p1[j].y;
        A1[r] = vx * vx + vy * vy;
        B1[r] = (t.d1[i] - t.d1[j]) * (t.d1[i] - t.d1[j]);
    }
    const double wx = t.d2[0] * t.p2[0].x - t.d2[1] * t.p2[1].x, wy = t.d2[0] * t.p2[0].y - t.d2[1] * t.p2[1].y;
    const double A2 = wx * wx + wy * wy, B2 = (t.d2[0] - t.d2[1]) * (t.d2[0] - t.d2[1]);
    const double n2 = t.p2[2].x * t.p2[2].x + t.p2[2].y * t.p2[2].y;
    const double c02 = t.d2[0] * (t.p2[0].x * t.p2[2].x + t.p2[0].y * t.p2[2].y);
    const double c12 = t.d2[1] * (t.p2[1].x * t.p2[2].x + t.p2[1].y * t.p2[2].y);
    const double m0 = t.d2[0] * t.d2[0] * (t.p2[0].x * t.p2[0].x + t.p2[0].y * t.p2[0].y);
    const double m1 = t.d2[1] * t.d2[1] * (t.p2[1].x * t.p2[1].x + t.p2[1].y * t.p2[1].y);
    const Poly<1> S1 = plin(A1[0], B1[0]), S2 = plin(A2, B2), T02 = plin(A1[1], B1[1]);
    const Poly<1> Td = plin(1.0 * A1[1] + -1.0 * A1[2], 1.0 * B1[1] + -1.0 * B1[2]);
    const Poly<1> L = plin(c02 - c12, t.d2[0] - t.d2[1]);
    const Poly<2> Na = pmul(S1, plin(m0 - m1, t.d2[0] * t.d2[0] - t.d2[1] * t.d2[1]));
    const Poly<2> Nb = pmul(S2, Td);
    Poly<2> N;
#pragma unroll
    for (int i = 0; i <= 2; ++i) N.c[i] = 1.0 * Na.c[i] + -1.0 * Nb.c[i];
    const Poly<2> LL = pmul(L, L);
    const Poly<5> lhs = pmul(pmul(pmul(S2, T02), S1), LL);
    const Poly<5> r1 = pmul(pmul(N, N), plin(n2, 1.0));
    const Poly<5> r2 = pmul(pmul(pmul(S1, N), L), plin(c02, t.d2[0]));
    const Poly<5> r3 = pmul(pmul(pmul(S1, S1), LL), plin(m0, t.d2[0] * t.d2[0]));
    double P[6];
#pragma unroll
    for (int i = 0; i <= 5; ++i)
        P[i] = 1.0 * (4.0 * lhs.c[i] + -1.0 * r1.c[i]) + 1.0 * (4.0 * r2.c[i] + -4.0 * r3.c[i]);
    double r0_ = 0, r1_ = 0, r2_ = 0, r3_ = 0;
    const int nr = solve_quartic_real(P[4] / P[5], P[3] / P[5], P[2] / P[5], P[1] / P[5], r0_, r1_, r2_, r3_);
    int n = 0;
#pragma unroll
    for (int ir = 0; ir < 4; ++ir) {
        if (ir >= nr) break;
        double g = ir == 0 ? r0_ : (ir == 1 ? r1_ : (ir == 2 ? r2_ : r3_));
        for (int it = 0; it < 3; ++it) {
            const double v = (((P[5] * g + P[4]) * g + P[3]) * g + P[2]) * g + P[1];
            const double dv = ((4.0 * P[5] * g + 3.0 * P[4]) * g + 2.0 * P[3]) * g + P[2];
            if (dv == 0.0) break;
            g -= v / dv;
        }
        if (!(g > 0.0)) continue;
        const double s1 = peval(S1, g), s2 = peval(S2, g);
        const double sc2 = s1 / s2;
        if (!(sc2 > 0.0)) continue;
        const double nu = peval(N, g) / (2.0 * s1 * peval(L, g));
        if (!(nu > 0.0)) continue;
        FocalCand c;
        c.g = g; c.sc2 = sc2; c.nu = nu;
        // static indexing keeps the candidates in registers
        if (n == 0) c0 = c;
        else if (n == 1) c1 = c;
        else if (n == 2) c2 = c;
        else c3 = c;
        ++n;
    }
    return n;
}

RP_HD Model solve_finish(const Triplet &t, const FocalCand &c) {
    const double w = sqrt(c.g);
    Model m = identity_model();
    m.scale = sqrt(c.sc2);
    m.f1 = 1.0 / w;
    m.f2 = m.f1;
    V3 X[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        X[i] = v3(t.d1[i] * t.p1[i].x * w, t.d1[i] * t.p1[i].y * w, t.d1[i]);
        const double dep = m.scale * (i < 2 ? t.d2[i] : c.nu);
        Y[i] = v3(dep * t.p2[i].x * w, dep * t.p2[i].y * w, dep);
    }
    align_triangles(X[0], X[1], X[2], Y[0], Y[1], Y[2], m.q, m.t);
    return m;
}

RP_HD void solve_shared_focal(const Triplet &t, ModelSet &out) {
    out.n = 0;
    FocalCand c0, c1, c2, c3;
    const int n = solve_roots(t, c0, c1, c2, c3);
    if (n > 0) set_model(out, solve_finish(t, c0));
    if (n > 1) set_model(out, solve_finish(t, c1));
    if (n > 2) set_model(out, solve_finish(t, c2));
    if (n > 3) set_model(out, solve_finish(t, c3));
}

RP_HD void solve_minimal(int variant, const Triplet &t, ModelSet &out) {
    switch (variant) {
    case RP_CALIB: solve_calib_scale(t, out); break;
    case RP_CALIB_SHIFT: solve_calib_shift(t, out); break;
    case RP_SHARED: solve_shared_focal(t, out); break;
    default: solve_varying_focal(t, out); break;
    }
}

}  // namespace rp
