// hostcheck.cpp — compiles the product's RP_HD device math (mdrp_b200/csrc/*.cuh) for the HOST so
// the CPU-only test tier can compare it with the oracle.  Test scaffolding only: this file is not
// part of librepose_b200.so and the product has no host execution path.
#include "../../mdrp_b200/csrc/rp_solvers.cuh"
#include "../../mdrp_b200/csrc/rp_score.cuh"
#include "../../mdrp_b200/csrc/rp_lm.cuh"
#include <cstring>

using namespace rp;
#define HC extern "C" __attribute__((visibility("default")))

HC int hc_solve(int variant, const double *x1h, const double *x2h, const double *d1, const double *d2,
                rp_model *out) {
    Triplet t;
    for (int i = 0; i < 3; ++i) {
        t.p1[i] = v3(x1h[3 * i], x1h[3 * i + 1], x1h[3 * i + 2]);
        t.p2[i] = v3(x2h[3 * i], x2h[3 * i + 1], x2h[3 * i + 2]);
        t.d1[i] = d1[i];
        t.d2[i] = d2[i];
    }
    ModelSet ms;
    solve_minimal(variant, t, ms);
    for (int k = 0; k < ms.n; ++k) std::memcpy(&out[k], &ms.m[k], sizeof(rp_model));
    return ms.n;
}

static long g_screen_contradictions = 0, g_screen_decided = 0, g_screen_undecided = 0;
HC void hc_screen_stats(long *out) { out[0] = g_screen_contradictions; out[1] = g_screen_decided; out[2] = g_screen_undecided; }

// full two-tier point decision as the scoring kernel makes it; also reports how many points
// the FP32 filter rejected (tier 0) so the test can check it never rejects a reference inlier
HC double hc_score(int variant, const rp_model *model, const double *x1, const double *x2, long n,
                   double sq_thr, long *count, long *tier0, unsigned char *mask) {
    Model m;
    std::memcpy(&m, model, sizeof(m));
    const bool pose = variant == RP_CALIB || variant == RP_CALIB_SHIFT;
    const M3 E = pose ? essential_from_motion(m.q, m.t) : fundamental_from_model(m);
    double Mmax = 0, mmax = 0;
    for (long k = 0; k < n; ++k) {
        double a = fabs(x1[2 * k]) + fabs(x1[2 * k + 1]) + 1.0, b = fabs(x2[2 * k]) + fabs(x2[2 * k + 1]) + 1.0;
        Mmax = fmax(Mmax, a * b);
        mmax = fmax(mmax, fmax(a, b));
    }
    const Filter32 f = make_filter32(E, sqrt(sq_thr), Mmax, mmax);
    const Cheir32 ch = make_cheir32(m.q, m.t);
    long c = 0, t0 = 0;
    double sum = 0.0;
    for (long k = 0; k < n; ++k) {
        bool inl = false;
        if (certain_outlier32(f, (float)x1[2 * k], (float)x1[2 * k + 1], (float)x2[2 * k], (float)x2[2 * k + 1])) {
            ++t0;
        } else {
            const double r2 = sampson_r2_exact(E, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]);
            if (r2 < sq_thr) {
                if (!pose) inl = true;
                else {
                    const bool exact = cheirality_exact(m.q, m.t, bearing(x1[2 * k], x1[2 * k + 1]), bearing(x2[2 * k], x2[2 * k + 1]));
                    const int scr = cheirality32(ch, (float)x1[2 * k], (float)x1[2 * k + 1], (float)x2[2 * k], (float)x2[2 * k + 1]);
                    if (scr != 0 && (scr > 0) != exact) g_screen_contradictions++;
                    if (scr != 0) g_screen_decided++; else g_screen_undecided++;
                    inl = scr != 0 ? scr > 0 : exact;
                }
                if (inl) { ++c; sum += r2; }
            }
        }
        if (mask) mask[k] = inl;
    }
    *count = c;
    *tier0 = t0;
    return sum + sq_thr * (double)(n - c);
}

template <int V, int NP>
static void accumulate_t(const Model &m, const LMParams &P, const double *x1, const double *x2, const double *d1,
                         const double *d2, long n, double *JtJ, double *Jtr, double *cost) {
    const LMFrame F = make_frame(m);
    NormalEq<NP> N;
    N.clear();
    double c = 0;
    LogProd lp;   // (only the device path parks Cauchy terms in it; on the host every term comes back through the return value)
    lp.init();
    for (long k = 0; k < n; ++k) {
        point_accumulate<V, NP>(F, P, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1], d1[k], d2[k], N);
        c += point_cost<V>(F, P, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1], d1[k], d2[k], lp);
    }
    c += P.loss_scale * P.loss_scale * lp.total(P.weight_sampson);
    std::memset(JtJ, 0, 81 * sizeof(double));
    std::memset(Jtr, 0, 9 * sizeof(double));
    for (int i = 0; i < NP; ++i) {
        Jtr[i] = N.g[i];
        for (int j = 0; j <= i; ++j) JtJ[9 * i + j] = JtJ[9 * j + i] = N.A[i * (i + 1) / 2 + j];
    }
    *cost = c;
}

HC void hc_accumulate(int variant, const rp_model *model, const double *x1, const double *x2, const double *d1,
                      const double *d2, long n, double scale_reproj, double weight_sampson, int loss_type,
                      double loss_scale, double *JtJ, double *Jtr, double *cost) {
    Model m;
    std::memcpy(&m, model, sizeof(m));
    LMParams P;
    P.scale_reproj = scale_reproj; P.weight_sampson = weight_sampson; P.loss_scale = loss_scale; P.loss_type = loss_type;
    P.inv_t2 = 1.0 / (loss_scale * loss_scale);
    switch (variant) {
    case RP_CALIB: accumulate_t<RP_CALIB, 7>(m, P, x1, x2, d1, d2, n, JtJ, Jtr, cost); break;
    case RP_CALIB_SHIFT: accumulate_t<RP_CALIB_SHIFT, 9>(m, P, x1, x2, d1, d2, n, JtJ, Jtr, cost); break;
    case RP_SHARED: accumulate_t<RP_SHARED, 8>(m, P, x1, x2, d1, d2, n, JtJ, Jtr, cost); break;
    default: accumulate_t<RP_VARYING, 9>(m, P, x1, x2, d1, d2, n, JtJ, Jtr, cost); break;
    }
}

HC void hc_llt_solve(int np, const double *JtJ, double lambda, const double *b, double *x) {
    double A[45];
    for (int i = 0; i < np; ++i)
        for (int j = 0; j <= i; ++j) A[i * (i + 1) / 2 + j] = JtJ[9 * i + j];
    if (np == 7) llt_solve<7>(A, lambda, b, x);
    else if (np == 8) llt_solve<8>(A, lambda, b, x);
    else llt_solve<9>(A, lambda, b, x);
}

HC void hc_step(int variant, const rp_model *model, const double *dp, rp_model *out) {
    Model m, o;
    std::memcpy(&m, model, sizeof(m));
    switch (variant) {
    case RP_CALIB: o = model_step<RP_CALIB>(m, dp); break;
    case RP_CALIB_SHIFT: o = model_step<RP_CALIB_SHIFT>(m, dp); break;
    case RP_SHARED: o = model_step<RP_SHARED>(m, dp); break;
    default: o = model_step<RP_VARYING>(m, dp); break;
    }
    std::memcpy(out, &o, sizeof(o));
}
