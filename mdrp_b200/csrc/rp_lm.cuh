// rp_lm.cuh — residuals and Jacobians of the hybrid monodepth refinement.
//
//   L2  refine_monodepth_relpose so@0x261030 (7 params, 9 with shifts),
//       refine_monodepth_shared_focal_relpose so@0x2592e0 (8), …_varying_focal_… so@0x260fa0 (9)
//       = lm_impl<MonoDepth*JacobianAccumulator<Loss,Weights>> of PoseLib bundle.cc/jacobian_impl.h
//
// cost = sum_k  w_s rho(r_s^2) + rho(s_r |pi(R (d1+u) p1 + t) - x2|^2)[z>0]
//                              + rho(s_r |pi(R^T (sigma (d2+v) p2 - t)) - x1|^2)[z>0]
// with p_i = K_i^-1 (x_i, 1) and r_s the Sampson residual of F = K2^-T [t]x R K1^-1.
// Parameter order: rotation tangent w (R <- R exp([w]x)), t, scale, then (shift1, shift2) | f | (f1, f2).
// SURVEY.md §8a rows L1/L2.  One residual row at a time is pushed into a packed lower-triangular
// J^T J held in registers (explicit FMAs: this path is not bit-pinned, the oracle comparison is
// cost <= oracle*(1+1e-9) and pose within 1e-6).
#pragma once
#include "rp_common.cuh"

namespace rp {

RP_HD double loss_eval(int type, double thr, double r2) {
    const double t2 = thr * thr;
    switch (type) {
    case RP_LOSS_TRIVIAL: return r2;
    case RP_LOSS_TRUNCATED: return r2 < t2 ? r2 : t2;
    case RP_LOSS_HUBER: { const double r = sqrt(r2); return r <= thr ? r2 : thr * (2.0 * r - thr); }
    case RP_LOSS_CAUCHY: return t2 * log1p(r2 / t2);
    case RP_LOSS_TRUNCATED_CAUCHY: return r2 > t2 ? t2 * log1p(1.0) : t2 * log1p(r2 / t2);
    default: return r2 < t2 ? r2 : t2;
    }
}
RP_HD double loss_weight(int type, double thr, double r2) {
    const double t2 = thr * thr;
    switch (type) {
    case RP_LOSS_TRIVIAL: return 1.0;
    case RP_LOSS_TRUNCATED: return r2 < t2 ? 1.0 : 0.0;
    case RP_LOSS_HUBER: { const double r = sqrt(r2); return r <= thr ? 1.0 : thr / r; }
    case RP_LOSS_CAUCHY: return 1.0 / (1.0 + r2 / t2);
    case RP_LOSS_TRUNCATED_CAUCHY: return r2 > t2 ? 0.0 : 1.0 / (1.0 + r2 / t2);
    default: return r2 < t2 ? 1.0 : 0.0;
    }
}

struct LMParams {
    double scale_reproj, weight_sampson, loss_scale;
    int loss_type;
};

// everything about the current model that is constant across points
struct LMFrame {
    M3 R, E;
    V3 t;
    double scale, shift1, shift2, f1, f2, if1sq, if2sq;
};

RP_HD LMFrame make_frame(const Model &m) {
    LMFrame F;
    F.R = quat_to_rotmat(m.q);
    F.E = essential_from_motion(m.q, m.t);
    F.t = m.t;
    F.scale = m.scale; F.shift1 = m.shift1; F.shift2 = m.shift2;
    F.f1 = m.f1; F.f2 = m.f2;
    F.if1sq = 1.0 / (m.f1 * m.f1);
    F.if2sq = 1.0 / (m.f2 * m.f2);
    return F;
}

template <int NP>
struct NormalEq {
    double A[NP * (NP + 1) / 2];  // packed lower triangle, row-major: (i,j<=i) at i(i+1)/2+j
    double g[NP];
    RP_HD void clear() {
#pragma unroll
        for (int i = 0; i < NP * (NP + 1) / 2; ++i) A[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) g[i] = 0.0;
    }
    RP_HD void add_row(double w, const double (&J)[NP], double r) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const double wi = w * J[i];
            g[i] = fma_(wi, r, g[i]);
#pragma unroll
            for (int j = 0; j <= i; ++j) A[i * (i + 1) / 2 + j] = fma_(wi, J[j], A[i * (i + 1) / 2 + j]);
        }
    }
};

RP_HD V3 unit(int i) { return v3(i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0); }
RP_HD double comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
RP_HD V3 row(const M3 &M, int i) { return i == 0 ? M.r0 : (i == 1 ? M.r1 : M.r2); }

// cost contribution of one correspondence
template <int VARIANT>
RP_HD double point_cost(const LMFrame &F, const LMParams &P, double x1_0, double x1_1, double x2_0, double x2_1,
                        double d1, double d2) {
    const V3 p1 = v3(x1_0 / F.f1, x1_1 / F.f1, 1.0);
    const V3 p2 = v3(x2_0 / F.f2, x2_1 / F.f2, 1.0);
    double cost = 0.0;
    if (P.weight_sampson > 0.0) {
        const V3 Ep1 = mul(F.E, p1), Etp2 = mulT(F.E, p2);
        const double C = dot(p2, Ep1);
        const double A = Ep1.x * Ep1.x + Ep1.y * Ep1.y, B = Etp2.x * Etp2.x + Etp2.y * Etp2.y;
        const double inv = 1.0 / sqrt(A * F.if2sq + B * F.if1sq);
        const double rs = C * inv;
        cost += P.weight_sampson * loss_eval(P.loss_type, P.loss_scale, rs * rs);
    }
    if (P.scale_reproj > 0.0) {
        const double a = d1 + F.shift1;
        V3 Z = mul(F.R, v3(a * p1.x, a * p1.y, a * p1.z));
        Z = Z + F.t;
        if (Z.z > 0.0) {
            const double iz = 1.0 / Z.z;
            const double r0 = F.f2 * (Z.x * iz) - x2_0, r1 = F.f2 * (Z.y * iz) - x2_1;
            cost += loss_eval(P.loss_type, P.loss_scale, P.scale_reproj * (r0 * r0 + r1 * r1));
        }
        const double b = F.scale * (d2 + F.shift2);
        const V3 Y = mulT(F.R, v3(b * p2.x - F.t.x, b * p2.y - F.t.y, b * p2.z - F.t.z));
        if (Y.z > 0.0) {
            const double iz = 1.0 / Y.z;
            const double r0 = F.f1 * (Y.x * iz) - x1_0, r1 = F.f1 * (Y.y * iz) - x1_1;
            cost += loss_eval(P.loss_type, P.loss_scale, P.scale_reproj * (r0 * r0 + r1 * r1));
        }
    }
    return cost;
}

// J^T J / J^T r contribution of one correspondence
template <int VARIANT, int NP>
RP_HD void point_accumulate(const LMFrame &F, const LMParams &P, double x1_0, double x1_1, double x2_0,
                            double x2_1, double d1, double d2, NormalEq<NP> &N) {
    constexpr bool FOCAL = (VARIANT == RP_SHARED || VARIANT == RP_VARYING);
    constexpr int CF2 = (VARIANT == RP_SHARED) ? 7 : 8;  // column of f2 (== f column when shared)
    const V3 p1 = v3(x1_0 / F.f1, x1_1 / F.f1, 1.0);
    const V3 p2 = v3(x2_0 / F.f2, x2_1 / F.f2, 1.0);
    double J[NP];

    // ---- Sampson row ----
    if (P.weight_sampson > 0.0) {
        const V3 Ep1 = mul(F.E, p1), Etp2 = mulT(F.E, p2);
        const double C = dot(p2, Ep1);
        const double A = Ep1.x * Ep1.x + Ep1.y * Ep1.y, B = Etp2.x * Etp2.x + Etp2.y * Etp2.y;
        const double inv = 1.0 / sqrt(A * F.if2sq + B * F.if1sq);
        const double rs = C * inv;
        const double w = P.weight_sampson * loss_weight(P.loss_type, P.loss_scale, rs * rs);
        if (w != 0.0) {
#pragma unroll
            for (int i = 0; i < NP; ++i) J[i] = 0.0;
            const double k = 0.5 * C * inv * inv * inv;
            const V3 Rp1 = mul(F.R, p1);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const V3 e = unit(i);
                // d/dw_i : dE = E [e_i]x
                const V3 exp1 = cross(e, p1);
                V3 dEp1 = mul(F.E, exp1);
                V3 tmp = cross(e, Etp2);
                double dC = dot(Etp2, exp1);
                double dden = 2.0 * (Ep1.x * dEp1.x + Ep1.y * dEp1.y) * F.if2sq +
                              2.0 * (Etp2.x * (-tmp.x) + Etp2.y * (-tmp.y)) * F.if1sq;
                J[i] = dC * inv - k * dden;
                // d/dt_i : dE = [e_i]x R
                dEp1 = cross(e, Rp1);
                tmp = mulT(F.R, cross(e, p2));
                dC = dot(p2, dEp1);
                dden = 2.0 * (Ep1.x * dEp1.x + Ep1.y * dEp1.y) * F.if2sq +
                       2.0 * (Etp2.x * (-tmp.x) + Etp2.y * (-tmp.y)) * F.if1sq;
                J[3 + i] = dC * inv - k * dden;
            }
            if (FOCAL) {
                const V3 dp1 = v3(-p1.x / F.f1, -p1.y / F.f1, 0.0);
                const V3 dp2 = v3(-p2.x / F.f2, -p2.y / F.f2, 0.0);
                const V3 dEp1 = mul(F.E, dp1), dEtp2 = mulT(F.E, dp2);
                const double dden1 = 2.0 * (Ep1.x * dEp1.x + Ep1.y * dEp1.y) * F.if2sq - 2.0 * B * F.if1sq / F.f1;
                const double dden2 = 2.0 * (Etp2.x * dEtp2.x + Etp2.y * dEtp2.y) * F.if1sq - 2.0 * A * F.if2sq / F.f2;
                const double j1 = dot(Etp2, dp1) * inv - k * dden1;
                const double j2 = dot(Ep1, dp2) * inv - k * dden2;
                if (VARIANT == RP_SHARED) J[7] = j1 + j2;
                else { J[7] = j1; J[CF2] = j2; }
            }
            N.add_row(w, J, rs);
        }
    }
    if (!(P.scale_reproj > 0.0)) return;

    // ---- reprojection 1 -> 2 ----
    {
        const double a = d1 + F.shift1;
        const V3 Pt = v3(a * p1.x, a * p1.y, a * p1.z);
        V3 Z = mul(F.R, Pt);
        Z = Z + F.t;
        if (Z.z > 0.0) {
            const double iz = 1.0 / Z.z;
            const double u0 = Z.x * iz, u1 = Z.y * iz;
            const double r0 = F.f2 * u0 - x2_0, r1 = F.f2 * u1 - x2_1;
            const double w = P.scale_reproj *
                             loss_weight(P.loss_type, P.loss_scale, P.scale_reproj * (r0 * r0 + r1 * r1));
            if (w != 0.0) {
                const double g = F.f2 * iz;
                double J0[NP], J1[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) { J0[i] = 0.0; J1[i] = 0.0; }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const V3 dZ = mul(F.R, cross(unit(i), Pt));
                    J0[i] = g * (dZ.x - u0 * dZ.z);
                    J1[i] = g * (dZ.y - u1 * dZ.z);
                    const V3 e = unit(i);
                    J0[3 + i] = g * (e.x - u0 * e.z);
                    J1[3 + i] = g * (e.y - u1 * e.z);
                }
                if (VARIANT == RP_CALIB_SHIFT) {
                    const V3 dZ = mul(F.R, p1);
                    J0[7] = g * (dZ.x - u0 * dZ.z);
                    J1[7] = g * (dZ.y - u1 * dZ.z);
                }
                if (FOCAL) {
                    const V3 dZ = mul(F.R, v3(-a * p1.x / F.f1, -a * p1.y / F.f1, 0.0));
                    J0[7] = g * (dZ.x - u0 * dZ.z);
                    J1[7] = g * (dZ.y - u1 * dZ.z);
                    J0[CF2] += u0;
                    J1[CF2] += u1;
                }
                N.add_row(w, J0, r0);
                N.add_row(w, J1, r1);
            }
        }
    }
    // ---- reprojection 2 -> 1 ----
    {
        const double bb = d2 + F.shift2;
        const double b = F.scale * bb;
        const V3 Y = mulT(F.R, v3(b * p2.x - F.t.x, b * p2.y - F.t.y, b * p2.z - F.t.z));
        if (Y.z > 0.0) {
            const double iz = 1.0 / Y.z;
            const double u0 = Y.x * iz, u1 = Y.y * iz;
            const double r0 = F.f1 * u0 - x1_0, r1 = F.f1 * u1 - x1_1;
            const double w = P.scale_reproj *
                             loss_weight(P.loss_type, P.loss_scale, P.scale_reproj * (r0 * r0 + r1 * r1));
            if (w != 0.0) {
                const double g = F.f1 * iz;
                double J0[NP], J1[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) { J0[i] = 0.0; J1[i] = 0.0; }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const V3 dY = cross(Y, unit(i));
                    J0[i] = g * (dY.x - u0 * dY.z);
                    J1[i] = g * (dY.y - u1 * dY.z);
                    const V3 dT = neg(row(F.R, i));
                    J0[3 + i] = g * (dT.x - u0 * dT.z);
                    J1[3 + i] = g * (dT.y - u1 * dT.z);
                }
                {
                    const V3 dY = mulT(F.R, v3(bb * p2.x, bb * p2.y, bb * p2.z));
                    J0[6] = g * (dY.x - u0 * dY.z);
                    J1[6] = g * (dY.y - u1 * dY.z);
                }
                if (VARIANT == RP_CALIB_SHIFT) {
                    const V3 dY = mulT(F.R, v3(F.scale * p2.x, F.scale * p2.y, F.scale * p2.z));
                    J0[8] = g * (dY.x - u0 * dY.z);
                    J1[8] = g * (dY.y - u1 * dY.z);
                }
                if (FOCAL) {
                    const V3 dY = mulT(F.R, v3(-b * p2.x / F.f2, -b * p2.y / F.f2, 0.0));
                    J0[CF2] += g * (dY.x - u0 * dY.z);
                    J1[CF2] += g * (dY.y - u1 * dY.z);
                    J0[7] += u0;
                    J1[7] += u1;
                }
                N.add_row(w, J0, r0);
                N.add_row(w, J1, r1);
            }
        }
    }
}

// parameter update of lm_impl's problem.step(): R <- R exp([dw]x), everything else additive
template <int VARIANT>
RP_HD Model model_step(const Model &m, const double *dp) {
    Model o = m;
    o.q = quat_mul(m.q, quat_exp(v3(dp[0], dp[1], dp[2])));
    o.t = v3(m.t.x + dp[3], m.t.y + dp[4], m.t.z + dp[5]);
    o.scale = m.scale + dp[6];
    if (VARIANT == RP_CALIB_SHIFT) { o.shift1 = m.shift1 + dp[7]; o.shift2 = m.shift2 + dp[8]; }
    if (VARIANT == RP_SHARED) { o.f1 = m.f1 + dp[7]; o.f2 = o.f1; }
    if (VARIANT == RP_VARYING) { o.f1 = m.f1 + dp[7]; o.f2 = m.f2 + dp[8]; }
    return o;
}

// (J^T J + lambda I) x = b by Cholesky on the packed lower triangle (Eigen's
// selfadjointView<Lower>().llt().solve()); a non-PD matrix propagates NaN like Eigen does.
template <int NP>
RP_HD void llt_solve(const double *A, double lambda, const double *b, double *x) {
    double L[NP * (NP + 1) / 2];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        double s = A[j * (j + 1) / 2 + j] + lambda;
#pragma unroll
        for (int k = 0; k < j; ++k) s -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
        const double d = sqrt(s);
        L[j * (j + 1) / 2 + j] = d;
#pragma unroll
        for (int i = j + 1; i < NP; ++i) {
            double v = A[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            L[i * (i + 1) / 2 + j] = v / d;
        }
    }
    double y[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * y[k];
        y[i] = v / L[i * (i + 1) / 2 + i];
    }
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < NP; ++k) v -= L[k * (k + 1) / 2 + i] * x[k];
        x[i] = v / L[i * (i + 1) / 2 + i];
    }
}

}  // namespace rp
