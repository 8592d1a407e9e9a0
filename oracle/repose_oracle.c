/*
 * repose_oracle.c — CPU restatement of the RePoseD LO-RANSAC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mdrp_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * The algorithm lives in a third-party dependency that is absent from
 * /root/reference as source: PoseLib 2.0.5 + PR #152 (kocurvik/PoseLib@pr-mdrp,
 * /root/reference/README.md:52), shipped as the binary wheel
 * /root/reference/demo/poselib-2.0.5-cp312-cp312-linux_x86_64.whl
 * (sha256 509a74fd…).  This file restates the published algorithm (PoseLib's
 * ransac_impl.h / sampling.cc / utils.cc / bundle.cc / p3p.cc / univariate.h
 * and the PR's monodepth solvers) and follows the call sites listed in
 * SURVEY.md §3.2/§8a; "so@0x…" are symbol addresses inside the wheel's _core.so.
 *
 * PARITY PINNING: the reference has no tests for this path, so
 * the restatement is pinned against outputs of the reference binary itself:
 * tests/test_oracle_vs_ref.py calls the wheel's exported C++ symbols through
 * ctypes (oracle/ref_wheel.py) on the same inputs, and tests/golden/ (npz files) holds
 * wheel-generated vectors (make_golden.py, make_golden_prosac.py, make_golden_extra.py)
 * for boxes without it.  Randomised end-to-end runs against the binary
 * (2 400 small scenes over all options) are what found the four quirks marked
 * "verified on the binary" below (weight_sampson^2 in the normal equations,
 * the focal accumulators' robust-weight argument, the final refinement not
 * writing model_score, std::min in the truncated loss).
 *
 * All arithmetic is FP64; build with -ffp-contract=off (no FMA), as the wheel is
 * plain SSE2 scalar code.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RO_API __attribute__((visibility("default")))

/* model: q (w first), t, scale, shift1, shift2, f1, f2 — 12 doubles
 * (MonoDepthTwoViewGeometry / MonoDepthImagePair, _core.pyi:171-204) */
typedef struct {
    double q[4];
    double t[3];
    double scale, shift1, shift2;
    double f1, f2;
} ro_model;

enum { RO_CALIB = 0, RO_CALIB_SHIFT = 1, RO_SHARED = 2, RO_VARYING = 3 };
enum { RO_LOSS_TRIVIAL = 0, RO_LOSS_TRUNCATED = 1, RO_LOSS_HUBER = 2, RO_LOSS_CAUCHY = 3,
       RO_LOSS_TRUNCATED_CAUCHY = 4, RO_LOSS_TRUNCATED_LE_ZACH = 5 };

typedef struct {
    int64_t max_iterations;
    int loss_type;
    double loss_scale, gradient_tol, step_tol, initial_lambda, min_lambda, max_lambda;
} ro_bundle_opt;

typedef struct {
    int64_t iterations;
    double initial_cost, cost, lambda;
    int64_t invalid_steps;
    double step_norm, grad_norm;
} ro_bundle_stats;

typedef struct {
    int64_t max_iterations, min_iterations;
    double dyn_num_trials_mult, success_prob, max_reproj_error, max_epipolar_error;
    uint64_t seed;
    int estimate_shift;
    double weight_sampson;
    int progressive_sampling;          /* RansacOptions::progressive_sampling (PROSAC) */
    int64_t max_prosac_iterations;     /* RansacOptions::max_prosac_iterations (default 100000) */
} ro_ransac_opt;

typedef struct {
    int64_t refinements, iterations, num_inliers;
    double inlier_ratio, model_score;
} ro_ransac_stats;

/* ------------------------------------------------------------------------- */
/* R2: sampler — RandomSampler::generate_sample so@0x4f8970, draw_sample
 * so@0x4f87f0, random_int so@0x4f87a0 (SplitMix64, low 32 bits as int).      */
RO_API int ro_random_int(uint64_t *state) {
    *state += 0x9e3779b97f4a7c15ULL;
    uint64_t z = *state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z = z ^ (z >> 31);
    return (int)z;
}

RO_API void ro_draw_sample(size_t sample_sz, size_t n, uint64_t *state, size_t *out) {
    for (size_t i = 0; i < sample_sz; ++i) {
        int done = 0;
        while (!done) {
            out[i] = (size_t)(int64_t)ro_random_int(state) % n;
            done = 1;
            for (size_t j = 0; j < i; ++j)
                if (out[i] == out[j]) { done = 0; break; }
        }
    }
}

/* RandomSampler (robust/sampling.cc): the PROSAC branch of generate_sample so@0x4f8970 and the growth
 * table of initialize_prosac so@0x4f8a20.  The data are assumed sorted by decreasing quality.  While
 * sample_k < max_prosac_iterations a sample is sample_sz-1 distinct indices out of the first subset_sz-1
 * points plus the point subset_sz-1 itself; the subset grows by one whenever sample_k passes
 * growth[subset_sz-1].  Afterwards (or with use_prosac = 0) it is the uniform draw_sample.          */
typedef struct {
    size_t num_data, sample_sz;
    uint64_t state;
    int use_prosac;
    size_t max_prosac_iterations, sample_k, subset_sz;
    size_t *growth;
} ro_sampler;

RO_API void ro_prosac_growth(size_t num_data, size_t sample_sz, size_t max_prosac_iterations, size_t *growth) {
    double T_n = (double)max_prosac_iterations;
    for (size_t i = 0; i < sample_sz; ++i) T_n *= (double)(sample_sz - i) / (double)(num_data - i);
    for (size_t i = 0; i < sample_sz; ++i) growth[i] = 1;
    size_t T_n_prime = 1;
    for (size_t n = sample_sz; n < num_data; ++n) {
        const double T_n_next = ((double)n + 1.0) * T_n / (((double)n + 1.0) - (double)sample_sz);
        T_n_prime = (size_t)((double)T_n_prime + ceil(T_n_next - T_n));
        growth[n] = T_n_prime;
        T_n = T_n_next;
    }
}

static void ro_sampler_init(ro_sampler *s, size_t num_data, size_t sample_sz, uint64_t seed, int use_prosac,
                            size_t max_prosac_iterations) {
    s->num_data = num_data; s->sample_sz = sample_sz; s->state = seed;
    s->use_prosac = use_prosac; s->max_prosac_iterations = max_prosac_iterations;
    s->sample_k = 0; s->subset_sz = 0; s->growth = NULL;
    if (use_prosac) {
        const size_t len = num_data > sample_sz ? num_data : sample_sz;
        s->growth = (size_t *)calloc(len, sizeof(size_t));
        ro_prosac_growth(num_data, sample_sz, max_prosac_iterations, s->growth);
        s->sample_k = 1;
        s->subset_sz = sample_sz;
    }
}

static void ro_sampler_generate(ro_sampler *s, size_t *sample) {
    if (s->use_prosac && s->sample_k < s->max_prosac_iterations) {
        ro_draw_sample(s->sample_sz - 1, s->subset_sz - 1, &s->state, sample);
        sample[s->sample_sz - 1] = s->subset_sz - 1;
        s->sample_k++;
        if (s->sample_k < s->max_prosac_iterations && s->sample_k > s->growth[s->subset_sz - 1]) {
            if (++s->subset_sz > s->num_data) s->subset_sz = s->num_data;
        }
    } else {
        ro_draw_sample(s->sample_sz, s->num_data, &s->state, sample);
    }
}

/* `iters` consecutive samples of a fresh sampler: out [iters, sample_sz] */
RO_API void ro_generate_samples(size_t num_data, size_t sample_sz, uint64_t seed, int use_prosac,
                                size_t max_prosac_iterations, size_t iters, size_t *out) {
    ro_sampler s;
    ro_sampler_init(&s, num_data, sample_sz, seed, use_prosac, max_prosac_iterations);
    for (size_t i = 0; i < iters; ++i) ro_sampler_generate(&s, out + i * sample_sz);
    free(s.growth);
}

/* ------------------------------------------------------------------------- */
/* small algebra                                                              */
static void quat_to_rotmat(const double q[4], double R[9]) {
    /* Eigen::Quaterniond(w,x,y,z).toRotationMatrix(), row-major out */
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

static void rotmat_to_quat(const double R[9], double q[4]) {
    /* Eigen::Quaterniond(R) then normalise (PoseLib rotmat_to_quat) */
    double tr = R[0] + R[4] + R[8];
    double w, x, y, z;
    if (tr > 0.0) {
        double t = sqrt(tr + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (R[7] - R[5]) * t;
        y = (R[2] - R[6]) * t;
        z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        double t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        w = (R[3 * k + j] - R[3 * j + k]) * t;
        v[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        v[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        x = v[0]; y = v[1]; z = v[2];
    }
    double n = sqrt(w * w + x * x + y * y + z * z);
    q[0] = w / n; q[1] = x / n; q[2] = y / n; q[3] = z / n;
}

static void quat_rotate(const double q[4], const double p[3], double out[3]) {
    /* PoseLib quat_rotate (misc/quaternion.h) */
    const double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
    const double p1 = p[0], p2 = p[1], p3 = p[2];
    const double px1 = -p1 * q2 - p2 * q3 - p3 * q4;
    const double px2 = p1 * q1 - p2 * q4 + p3 * q3;
    const double px3 = p2 * q1 + p1 * q4 - p3 * q2;
    const double px4 = p2 * q2 - p1 * q3 + p3 * q1;
    out[0] = px2 * q1 - px1 * q2 - px3 * q4 + px4 * q3;
    out[1] = px3 * q1 - px1 * q3 + px2 * q4 - px4 * q2;
    out[2] = px3 * q2 - px2 * q3 - px1 * q4 + px4 * q1;
}

static void quat_multiply(const double a[4], const double b[4], double o[4]) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

static void quat_exp(const double w[3], double q[4]) {
    /* so@0x262fd0 quat_exp: unit quaternion of the rotation vector w */
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
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
    q[0] = re; q[1] = im * w[0]; q[2] = im * w[1]; q[3] = im * w[2];
}

static void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(const double M[9], const double v[3], double o[3]) {
    o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}
static void matTvec3(const double M[9], const double v[3], double o[3]) {
    o[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    o[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    o[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
static int inv3(const double M[9], double o[9]) {
    double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (M[2] * M[7] - M[1] * M[8]) * id; o[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    o[3] = c01 * id; o[4] = (M[0] * M[8] - M[2] * M[6]) * id; o[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    o[6] = c02 * id; o[7] = (M[1] * M[6] - M[0] * M[7]) * id; o[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return det != 0.0;
}
static void matmul3(const double A[9], const double B[9], double O[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            O[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* essential_from_motion so@0x1dcb60: E = [t]x * R(q), row-major out */
RO_API void ro_essential_from_motion(const double q[4], const double t[3], double E[9]) {
    double R[9];
    quat_to_rotmat(q, R);
    const double T[9] = {0.0, -t[2], t[1], t[2], 0.0, -t[0], -t[1], t[0], 0.0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            E[3 * i + j] = (T[3 * i] * R[j] + T[3 * i + 1] * R[3 + j]) + T[3 * i + 2] * R[6 + j];
}

/* check_cheirality so@0x1dce00 on (x,y,1)/|.| bearings */
static int cheirality(const double q[4], const double t[3], double x1_0, double x1_1, double x2_0,
                      double x2_1, double min_depth) {
    double n1 = sqrt(x1_0 * x1_0 + x1_1 * x1_1 + 1.0);
    double n2 = sqrt(x2_0 * x2_0 + x2_1 * x2_1 + 1.0);
    double b1[3] = {x1_0 / n1, x1_1 / n1, 1.0 / n1};
    double b2[3] = {x2_0 / n2, x2_1 / n2, 1.0 / n2};
    double Rb1[3];
    quat_rotate(q, b1, Rb1);
    const double a = -(Rb1[0] * b2[0] + Rb1[1] * b2[1] + Rb1[2] * b2[2]);
    const double be1 = -(Rb1[0] * t[0] + Rb1[1] * t[1] + Rb1[2] * t[2]);
    const double be2 = b2[0] * t[0] + b2[1] * t[1] + b2[2] * t[2];
    const double lambda1 = be1 - a * be2;
    const double lambda2 = -a * be1 + be2;
    min_depth = min_depth * (1.0 - a * a);
    return lambda1 > min_depth && lambda2 > min_depth;
}

/* SC: compute_sampson_msac_score(CameraPose,…) so@0x4f61d0 (SURVEY.md §8a row SC).
 * x1,x2: [n,2] row-major normalised image points. */
static double sampson_r2(const double E[9], double x1_0, double x1_1, double x2_0, double x2_1) {
    const double Ex1_0 = E[0] * x1_0 + E[1] * x1_1 + E[2];
    const double Ex1_1 = E[3] * x1_0 + E[4] * x1_1 + E[5];
    const double Ex1_2 = E[6] * x1_0 + E[7] * x1_1 + E[8];
    const double Ex2_0 = E[0] * x2_0 + E[3] * x2_1 + E[6];
    const double Ex2_1 = E[1] * x2_0 + E[4] * x2_1 + E[7];
    const double C = x2_0 * Ex1_0 + x2_1 * Ex1_1 + Ex1_2;
    const double Cx = Ex1_0 * Ex1_0 + Ex1_1 * Ex1_1;
    const double Cy = Ex2_0 * Ex2_0 + Ex2_1 * Ex2_1;
    return C * C / (Cx + Cy);
}

RO_API double ro_msac_score_pose(const double q[4], const double t[3], const double *x1, const double *x2,
                                 size_t n, double sq_thr, size_t *count) {
    double E[9];
    ro_essential_from_motion(q, t, E);
    double score = 0.0;
    size_t c = 0;
    for (size_t k = 0; k < n; ++k) {
        const double r2 = sampson_r2(E, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]);
        if (r2 < sq_thr && cheirality(q, t, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1], 0.01)) {
            c++;
            score += r2;
        } else {
            score += sq_thr;
        }
    }
    *count = c;
    return score;
}

/* SF: compute_sampson_msac_score(Matrix3d F,…) so@0x4f65d0 — no cheirality. F row-major. */
RO_API double ro_msac_score_F(const double F[9], const double *x1, const double *x2, size_t n, double sq_thr,
                              size_t *count) {
    double score = 0.0;
    size_t c = 0;
    for (size_t k = 0; k < n; ++k) {
        const double r2 = sampson_r2(F, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]);
        if (r2 < sq_thr) {
            c++;
            score += r2;
        } else {
            score += sq_thr;
        }
    }
    *count = c;
    return score;
}

/* I1: get_inliers so@0x4f7a10 / so@0x4f77f0 */
RO_API size_t ro_get_inliers_pose(const double q[4], const double t[3], const double *x1, const double *x2,
                                  size_t n, double sq_thr, char *mask) {
    double E[9];
    ro_essential_from_motion(q, t, E);
    size_t c = 0;
    for (size_t k = 0; k < n; ++k) {
        const double r2 = sampson_r2(E, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]);
        int in = r2 < sq_thr && cheirality(q, t, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1], 0.01);
        mask[k] = (char)in;
        c += in;
    }
    return c;
}

RO_API size_t ro_get_inliers_F(const double F[9], const double *x1, const double *x2, size_t n, double sq_thr,
                               char *mask) {
    size_t c = 0;
    for (size_t k = 0; k < n; ++k) {
        int in = sampson_r2(F, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]) < sq_thr;
        mask[k] = (char)in;
        c += in;
    }
    return c;
}

/* F = diag(1,1,f2) * E * diag(1,1,f1) — focal estimators' score_model so@0x4fac60 / so@0x4faf90 */
RO_API void ro_fundamental_from_model(const ro_model *m, double F[9]) {
    double E[9];
    ro_essential_from_motion(m->q, m->t, E);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = E[3 * i + j];
            if (i == 2) v = m->f2 * v;
            if (j == 2) v = v * m->f1;
            F[3 * i + j] = v;
        }
}

/* ------------------------------------------------------------------------- */
/* univariate.h: solve_cubic_single_real so@0x1dabf0, solve_quartic_real so@0x1dc570 */
RO_API int ro_solve_cubic_single_real(double c2, double c1, double c0, double *root) {
    double a = c1 - c2 * c2 / 3.0;
    double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
    double c = b * b / 4.0 + a * a * a / 27.0;
    if (c != 0) {
        if (c > 0) {
            c = sqrt(c);
            b *= -0.5;
            *root = cbrt(b + c) + cbrt(b - c) - c2 / 3.0;
            return 1;
        } else {
            c = 3.0 * b / (2.0 * a) * sqrt(-3.0 / a);
            *root = 2.0 * sqrt(-a / 3.0) * cos(acos(c) / 3.0) - c2 / 3.0;
        }
    } else {
        *root = -c2 / 3.0 + (a != 0 ? (3.0 * b / a) : 0);
    }
    return 0;
}

static double sgn(double x) { return x < 0 ? -1.0 : 1.0; }

RO_API int ro_solve_quartic_real(double b, double c, double d, double e, double roots[4]) {
    /* x^4 + b x^3 + c x^2 + d x + e: depressed quartic, resolvent cubic, two quadratics,
     * one Newton step per root */
    double p = c - 3.0 * b * b / 8.0;
    double q = b * b * b / 8.0 - 0.5 * b * c + d;
    double r = (-3.0 * b * b * b * b + 256.0 * e - 64.0 * b * d + 16.0 * b * b * c) / 256.0;
    double bb = 2.0 * p;
    double cc = p * p - 4.0 * r;
    double dd = -q * q;
    double u2;
    ro_solve_cubic_single_real(bb, cc, dd, &u2);
    if (u2 < 0) return 0;
    double u = sqrt(u2);
    double s = -u;
    double t = (p + u * u + q / u) / 2.0;
    double v = (p + u * u - q / u) / 2.0;
    int sols = 0;
    double disc = u * u - 4.0 * v;
    if (disc > 0) {
        roots[0] = (-u - sgn(u) * sqrt(disc)) / 2.0;
        roots[1] = v / roots[0];
        sols += 2;
    }
    disc = s * s - 4.0 * t;
    if (disc > 0) {
        roots[sols] = (-s - sgn(s) * sqrt(disc)) / 2.0;
        roots[sols + 1] = t / roots[sols];
        sols += 2;
    }
    for (int i = 0; i < sols; i++) {
        roots[i] = roots[i] - b / 4.0;
        double x = roots[i];
        double x2 = x * x;
        double x3 = x * x2;
        double dx = -(x2 * x2 + b * x3 + c * x2 + d * x + e) / (4.0 * x3 + 3.0 * b * x2 + 2.0 * c * x + d);
        roots[i] = x + dx;
    }
    return sols;
}

static int root2real(double b, double c, double *r1, double *r2) {
    double THRESHOLD = -1.0e-12;
    double v = b * b - 4.0 * c;
    if (v < THRESHOLD) {
        *r1 = *r2 = -0.5 * b;
        return v >= 0;
    }
    if (v > THRESHOLD && v < 0.0) {
        *r1 = -0.5 * b;
        *r2 = -2;
        return 1;
    }
    double y = sqrt(v);
    if (b < 0) {
        *r1 = 0.5 * (-b + y);
        *r2 = 0.5 * (-b - y);
    } else {
        *r1 = 2.0 * c / (-b + y);
        *r2 = 2.0 * c / (-b - y);
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* S1: p3p so@0xecd50 — Ding et al., "Revisiting the P3P Problem" (CVPR'23) as in
 * PoseLib solvers/p3p.cc: cubic -> degenerate conic -> two lines -> quadratics,
 * Gauss-Newton polish of the three depths, pose from the aligned triangle.      */
static void refine_lambda(double *l1, double *l2, double *l3, double a12, double a13, double a23, double b12,
                          double b13, double b23) {
    for (int iter = 0; iter < 5; ++iter) {
        double r1 = (*l1 * *l1 - 2.0 * *l1 * *l2 * b12 + *l2 * *l2 - a12);
        double r2 = (*l1 * *l1 - 2.0 * *l1 * *l3 * b13 + *l3 * *l3 - a13);
        double r3 = (*l2 * *l2 - 2.0 * *l2 * *l3 * b23 + *l3 * *l3 - a23);
        if (fabs(r1) + fabs(r2) + fabs(r3) < 1e-10) return;
        double x11 = *l1 - *l2 * b12, x12 = *l2 - *l1 * b12;
        double x21 = *l1 - *l3 * b13, x23 = *l3 - *l1 * b13;
        double x32 = *l2 - *l3 * b23, x33 = *l3 - *l2 * b23;
        double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
        *l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
        *l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
        *l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
    }
}

static void compute_pq(double C[9], double p[3], double q[3]) {
    double A[9];
    A[0] = C[5] * C[7] - C[4] * C[8];
    A[4] = C[2] * C[6] - C[0] * C[8];
    A[8] = C[1] * C[3] - C[0] * C[4];
    A[1] = C[1] * C[8] - C[2] * C[7];
    A[2] = C[2] * C[4] - C[1] * C[5];
    A[3] = A[1];
    A[5] = C[0] * C[5] - C[2] * C[3];
    A[6] = A[2];
    A[7] = A[5];
    double v[3];
    int col;
    if (A[0] > A[4]) col = (A[0] > A[8]) ? 0 : 2;
    else col = (A[4] > A[8]) ? 1 : 2;
    double s = sqrt(A[4 * col]);
    v[0] = A[col] / s; v[1] = A[3 + col] / s; v[2] = A[6 + col] / s;
    C[1] -= v[2]; C[2] += v[1]; C[5] -= v[0];
    C[3] += v[2]; C[6] -= v[1]; C[7] += v[0];
    p[0] = C[0]; p[1] = C[3]; p[2] = C[6]; /* col 0 */
    q[0] = C[0]; q[1] = C[1]; q[2] = C[2]; /* row 0 */
}

/* x: 3 unit bearings (row-major 3x3), X: 3 points. out: up to 4 (q,t). */
RO_API int ro_p3p(const double *x_in, const double *X_in, double *q_out, double *t_out) {
    double X[3][3], x[3][3];
    memcpy(X, X_in, sizeof(X));
    memcpy(x, x_in, sizeof(x));
    double X01[3], X02[3], X12[3];
    for (int i = 0; i < 3; ++i) {
        X01[i] = X[0][i] - X[1][i];
        X02[i] = X[0][i] - X[2][i];
        X12[i] = X[1][i] - X[2][i];
    }
    double a01 = dot3(X01, X01), a02 = dot3(X02, X02), a12 = dot3(X12, X12);
    double tmp[3], tv;
#define SWAPV(u, v) do { memcpy(tmp, u, 24); memcpy(u, v, 24); memcpy(v, tmp, 24); } while (0)
#define SWAPD(u, v) do { tv = u; u = v; v = tv; } while (0)
    /* make BC (12) the largest side */
    if (a01 > a02) {
        if (a01 > a12) {
            SWAPV(x[0], x[2]); SWAPV(X[0], X[2]); SWAPD(a01, a12);
            for (int i = 0; i < 3; ++i) { X01[i] = -X12[i]; X02[i] = -X02[i]; }
        }
    } else if (a02 > a12) {
        SWAPV(x[0], x[1]); SWAPV(X[0], X[1]); SWAPD(a02, a12);
        for (int i = 0; i < 3; ++i) { X01[i] = -X01[i]; X02[i] = X12[i]; }
    }
    const double a12d = 1.0 / a12;
    const double a = a01 * a12d, b = a02 * a12d;
    const double m01 = dot3(x[0], x[1]), m02 = dot3(x[0], x[2]), m12 = dot3(x[1], x[2]);
    const double m12sq = -m12 * m12 + 1.0, m02sq = -1.0 + m02 * m02, m01sq = -1.0 + m01 * m01;
    const double ab = a * b, bsq = b * b, asq = a * a;
    const double m013 = -2.0 + 2.0 * m01 * m02 * m12;
    const double bsqm12sq = bsq * m12sq, asqm12sq = asq * m12sq, abm12sq = 2.0 * ab * m12sq;
    const double k3_inv = 1.0 / (bsqm12sq + b * m02sq);
    const double k2 = k3_inv * ((-1.0 + a) * m02sq + abm12sq + bsqm12sq + b * m013);
    const double k1 = k3_inv * (asqm12sq + abm12sq + a * m013 + (-1.0 + b) * m01sq);
    const double k0 = k3_inv * (asqm12sq + a * m01sq);
    double s;
    int G = ro_solve_cubic_single_real(k2, k1, k0, &s);
    double C[9];
    C[0] = -a + s * (1 - b); C[1] = -m02 * s; C[2] = a * m12 + b * m12 * s;
    C[3] = C[1]; C[4] = s + 1; C[5] = -m01;
    C[6] = C[2]; C[7] = C[5]; C[8] = -a - b * s + 1;
    double pq[2][3];
    compute_pq(C, pq[0], pq[1]);
    double XXm[9], XX[9], cx[3];
    cross3(X01, X02, cx);
    for (int i = 0; i < 3; ++i) { XXm[3 * i] = X01[i]; XXm[3 * i + 1] = X02[i]; XXm[3 * i + 2] = cx[i]; }
    inv3(XXm, XX);
    int n_sols = 0;
    for (int i = 0; i < 2; ++i) {
        double p0 = pq[i][0], p1 = pq[i][1], p2 = pq[i][2];
        int switch_12 = fabs(p0) <= fabs(p1);
        double taus[2];
        if (switch_12) {
            double w0 = -p0 / p1, w1 = -p2 / p1;
            double ca = 1.0 / (w1 * w1 - b);
            double cb = 2.0 * (b * m12 - m02 * w1 + w0 * w1) * ca;
            double cc = (w0 * w0 - 2 * m02 * w0 - b + 1.0) * ca;
            if (!root2real(cb, cc, &taus[0], &taus[1])) continue;
            for (int k = 0; k < 2; ++k) {
                double tau = taus[k];
                if (tau <= 0) continue;
                double d2 = sqrt(a12 / (tau * (tau - 2.0 * m12) + 1.0));
                double d1 = tau * d2;
                double d0 = (w0 * d2 + w1 * d1);
                if (d0 < 0) continue;
                refine_lambda(&d0, &d1, &d2, a01, a02, a12, m01, m02, m12);
                double v1[3], v2[3], v3[3], YY[9], R[9];
                for (int j = 0; j < 3; ++j) { v1[j] = d0 * x[0][j] - d1 * x[1][j]; v2[j] = d0 * x[0][j] - d2 * x[2][j]; }
                cross3(v1, v2, v3);
                for (int j = 0; j < 3; ++j) { YY[3 * j] = v1[j]; YY[3 * j + 1] = v2[j]; YY[3 * j + 2] = v3[j]; }
                matmul3(YY, XX, R);
                double RX[3];
                matvec3(R, X[0], RX);
                rotmat_to_quat(R, q_out + 4 * n_sols);
                for (int j = 0; j < 3; ++j) t_out[3 * n_sols + j] = d0 * x[0][j] - RX[j];
                ++n_sols;
                if (n_sols == 4) return n_sols;
            }
        } else {
            double w0 = -p1 / p0, w1 = -p2 / p0;
            double ca = 1.0 / (-a * w1 * w1 + 2 * a * m12 * w1 - a + 1);
            double cb = 2 * (a * m12 * w0 - m01 - a * w0 * w1) * ca;
            double cc = (1 - a * w0 * w0) * ca;
            if (!root2real(cb, cc, &taus[0], &taus[1])) continue;
            for (int k = 0; k < 2; ++k) {
                double tau = taus[k];
                if (tau <= 0) continue;
                double d0 = sqrt(a01 / (tau * (tau - 2.0 * m01) + 1.0));
                double d1 = tau * d0;
                double d2 = w0 * d0 + w1 * d1;
                if (d2 < 0) continue;
                refine_lambda(&d0, &d1, &d2, a01, a02, a12, m01, m02, m12);
                double v1[3], v2[3], v3[3], YY[9], R[9];
                for (int j = 0; j < 3; ++j) { v1[j] = d0 * x[0][j] - d1 * x[1][j]; v2[j] = d0 * x[0][j] - d2 * x[2][j]; }
                cross3(v1, v2, v3);
                for (int j = 0; j < 3; ++j) { YY[3 * j] = v1[j]; YY[3 * j + 1] = v2[j]; YY[3 * j + 2] = v3[j]; }
                matmul3(YY, XX, R);
                double RX[3];
                matvec3(R, X[0], RX);
                rotmat_to_quat(R, q_out + 4 * n_sols);
                for (int j = 0; j < 3; ++j) t_out[3 * n_sols + j] = d0 * x[0][j] - RX[j];
                ++n_sols;
                if (n_sols == 4) return n_sols;
            }
        }
        if (n_sols > 0 && G) break;
    }
    return n_sols;
}

/* exact rigid alignment of two congruent triangles: R,t with Y_i = R X_i + t */
static void align3(const double X[3][3], const double Y[3][3], double q[4], double t[3]) {
    double a1[3], b1[3], c1[3], a2[3], b2[3], c2[3];
    for (int i = 0; i < 3; ++i) {
        a1[i] = X[1][i] - X[0][i]; b1[i] = X[2][i] - X[0][i];
        a2[i] = Y[1][i] - Y[0][i]; b2[i] = Y[2][i] - Y[0][i];
    }
    cross3(a1, b1, c1);
    cross3(a2, b2, c2);
    double M1[9], M2[9], M1i[9], R[9];
    for (int i = 0; i < 3; ++i) {
        M1[3 * i] = a1[i]; M1[3 * i + 1] = b1[i]; M1[3 * i + 2] = c1[i];
        M2[3 * i] = a2[i]; M2[3 * i + 1] = b2[i]; M2[3 * i + 2] = c2[i];
    }
    inv3(M1, M1i);
    matmul3(M2, M1i, R);
    rotmat_to_quat(R, q);
    double Rq[9], RX[3];
    quat_to_rotmat(q, Rq);
    matvec3(Rq, X[0], RX);
    for (int i = 0; i < 3; ++i) t[i] = Y[0][i] - RX[i];
}

static void model_init(ro_model *m) {
    m->q[0] = 1; m->q[1] = m->q[2] = m->q[3] = 0;
    m->t[0] = m->t[1] = m->t[2] = 0;
    m->scale = 1; m->shift1 = m->shift2 = 0; m->f1 = m->f2 = 1;
}

/* S1: RelativePoseMonoDepthEstimator::generate_models so@0x4fe090, scale-only branch.
 * x1h,x2h: 3 homogeneous (x,y,1) normalised points. */
RO_API int ro_solve_calib_scale(const double *x1h, const double *x2h, const double *d1, const double *d2,
                                ro_model *out) {
    double X[9], b[9];
    for (int i = 0; i < 3; ++i) {
        double n = sqrt(x2h[3 * i] * x2h[3 * i] + x2h[3 * i + 1] * x2h[3 * i + 1] + x2h[3 * i + 2] * x2h[3 * i + 2]);
        for (int j = 0; j < 3; ++j) {
            X[3 * i + j] = d1[i] * x1h[3 * i + j];
            b[3 * i + j] = x2h[3 * i + j] / n;
        }
    }
    double qs[16], ts[12];
    int n = ro_p3p(b, X, qs, ts);
    for (int k = 0; k < n; ++k) {
        model_init(&out[k]);
        memcpy(out[k].q, qs + 4 * k, 32);
        memcpy(out[k].t, ts + 3 * k, 24);
        double R[9], RX[3];
        quat_to_rotmat(out[k].q, R);
        matvec3(R, X, RX);
        out[k].scale = (RX[0] + out[k].t[0]) / (d2[0] * x2h[0]);
    }
    return n;
}

/* S2: relpose_monodepth_3pt so@0x155ca0 (solver_p3p_mono_3d so@0x154760 + refine_suv so@0x15de40).
 * Unknowns s = scale^2, u = shift1, v = shift2.  For the three point pairs (i,j):
 *   s*|(d2_i+v) x2_i - (d2_j+v) x2_j|^2 = |(d1_i+u) x1_i - (d1_j+u) x1_j|^2 ,
 * i.e. c0*s*v^2 + c1*u^2 + c2*s*v + c3*s + c4*u + c5 = 0.  Linear elimination of
 * (s v^2, s v, s) gives quadratics in u; (s v)^2 = (s v^2)(s) is a quartic in u.
 * Raw roots are filtered (s>0, all six shifted depths >0), then polished by five
 * Gauss-Newton steps (tolerance 1e-10 on the sum of |residuals|), as the binary does. */
static int solve3(const double A[9], const double b[3], double x[3]) {
    double Ai[9];
    if (!inv3(A, Ai)) return 0;
    matvec3(Ai, b, x);
    return 1;
}

RO_API int ro_solve_calib_shift(const double *x1h, const double *x2h, const double *d1, const double *d2,
                                ro_model *out) {
    static const int PI[3] = {0, 0, 1}, PJ[3] = {1, 2, 2};
    double cf[3][6];
    for (int r = 0; r < 3; ++r) {
        const double *p1i = x1h + 3 * PI[r], *p1j = x1h + 3 * PJ[r];
        const double *p2i = x2h + 3 * PI[r], *p2j = x2h + 3 * PJ[r];
        double n1i = dot3(p1i, p1i), n1j = dot3(p1j, p1j), c1 = dot3(p1i, p1j);
        double n2i = dot3(p2i, p2i), n2j = dot3(p2j, p2j), c2 = dot3(p2i, p2j);
        double ai = d1[PI[r]], aj = d1[PJ[r]], bi = d2[PI[r]], bj = d2[PJ[r]];
        cf[r][0] = n2i + n2j - 2.0 * c2;
        cf[r][1] = -(n1i + n1j - 2.0 * c1);
        cf[r][2] = 2.0 * (bi * n2i + bj * n2j - c2 * (bi + bj));
        cf[r][3] = bi * bi * n2i + bj * bj * n2j - 2.0 * bi * bj * c2;
        cf[r][4] = -2.0 * (ai * n1i + aj * n1j - c1 * (ai + aj));
        cf[r][5] = -(ai * ai * n1i + aj * aj * n1j - 2.0 * ai * aj * c1);
    }
    /* [s v^2, s v, s]^T = -A^-1 * (c1 u^2 + c4 u + c5) */
    double A[9], Ai[9];
    for (int r = 0; r < 3; ++r) { A[3 * r] = cf[r][0]; A[3 * r + 1] = cf[r][2]; A[3 * r + 2] = cf[r][3]; }
    if (!inv3(A, Ai)) return 0;
    double P[3][3]; /* rows: s v^2, s v, s ; cols: u^2, u, 1 */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            int idx = (c == 0) ? 1 : (c == 1 ? 4 : 5);
            P[r][c] = -(Ai[3 * r] * cf[0][idx] + Ai[3 * r + 1] * cf[1][idx] + Ai[3 * r + 2] * cf[2][idx]);
        }
    /* quartic (s v)^2 - (s v^2)(s) */
    const double *a = P[1], *b = P[0], *c = P[2];
    double k4 = a[0] * a[0] - b[0] * c[0];
    double k3 = 2.0 * a[0] * a[1] - (b[0] * c[1] + b[1] * c[0]);
    double k2 = 2.0 * a[0] * a[2] + a[1] * a[1] - (b[0] * c[2] + b[1] * c[1] + b[2] * c[0]);
    double k1 = 2.0 * a[1] * a[2] - (b[1] * c[2] + b[2] * c[1]);
    double k0 = a[2] * a[2] - b[2] * c[2];
    double roots[4];
    int nr = ro_solve_quartic_real(k3 / k4, k2 / k4, k1 / k4, k0 / k4, roots);
    int n = 0;
    for (int ir = 0; ir < nr; ++ir) {
        double u = roots[ir];
        double s = (c[0] * u + c[1]) * u + c[2];
        double sv = (a[0] * u + a[1]) * u + a[2];
        double v = sv / s;
        if (!(s > 0)) continue;
        if (!(d1[0] + u > 0 && d1[1] + u > 0 && d1[2] + u > 0)) continue;
        if (!(d2[0] + v > 0 && d2[1] + v > 0 && d2[2] + v > 0)) continue;
        /* refine_suv: 5 Gauss-Newton steps on the three equations */
        for (int it = 0; it < 5; ++it) {
            double r[3], J[9];
            for (int e = 0; e < 3; ++e) {
                const double *k = cf[e];
                r[e] = k[0] * s * v * v + k[1] * u * u + k[2] * s * v + k[3] * s + k[4] * u + k[5];
                J[3 * e] = k[0] * v * v + k[2] * v + k[3];
                J[3 * e + 1] = 2.0 * k[1] * u + k[4];
                J[3 * e + 2] = 2.0 * k[0] * s * v + k[2] * s;
            }
            if (fabs(r[0]) + fabs(r[1]) + fabs(r[2]) < 1e-10) break;
            double dx[3];
            if (!solve3(J, r, dx)) break;
            s -= dx[0]; u -= dx[1]; v -= dx[2];
        }
        ro_model *m = &out[n];
        model_init(m);
        m->scale = sqrt(s);
        m->shift1 = u;
        m->shift2 = v;
        double X[3][3], Y[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                X[i][j] = (d1[i] + u) * x1h[3 * i + j];
                Y[i][j] = m->scale * (d2[i] + v) * x2h[3 * i + j];
            }
        align3(X, Y, m->q, m->t);
        ++n;
    }
    return n;
}

/* S4: relpose_monodepth_3pt_varying_focal so@0x19bcd0 — linear 3x3 in
 * (a=1/f1^2, b=scale^2, c=scale^2/f2^2), SURVEY.md §8a row S4. x: (x,y,1) pixels (pp-centred). */
RO_API int ro_solve_varying_focal(const double *x1h, const double *x2h, const double *d1, const double *d2,
                                  ro_model *out) {
    static const int PI[3] = {0, 0, 1}, PJ[3] = {1, 2, 2};
    double A[9], rhs[3], sol[3];
    for (int r = 0; r < 3; ++r) {
        int i = PI[r], j = PJ[r];
        double ax = d1[i] * x1h[3 * i] - d1[j] * x1h[3 * j], ay = d1[i] * x1h[3 * i + 1] - d1[j] * x1h[3 * j + 1];
        double bx = d2[i] * x2h[3 * i] - d2[j] * x2h[3 * j], by = d2[i] * x2h[3 * i + 1] - d2[j] * x2h[3 * j + 1];
        double dz1 = d1[i] - d1[j], dz2 = d2[i] - d2[j];
        A[3 * r] = ax * ax + ay * ay;
        A[3 * r + 1] = -dz2 * dz2;
        A[3 * r + 2] = -(bx * bx + by * by);
        rhs[r] = -dz1 * dz1;
    }
    if (!solve3(A, rhs, sol)) return 0;
    if (!(sol[0] > 0 && sol[1] > 0 && sol[2] > 0)) return 0;
    double f1 = 1.0 / sqrt(sol[0]);
    double scale = sqrt(sol[1]);
    double f2 = sqrt(sol[1] / sol[2]);
    ro_model *m = &out[0];
    model_init(m);
    m->scale = scale; m->f1 = f1; m->f2 = f2;
    double X[3][3], Y[3][3];
    for (int i = 0; i < 3; ++i) {
        X[i][0] = d1[i] * x1h[3 * i] / f1; X[i][1] = d1[i] * x1h[3 * i + 1] / f1; X[i][2] = d1[i];
        Y[i][0] = scale * d2[i] * x2h[3 * i] / f2; Y[i][1] = scale * d2[i] * x2h[3 * i + 1] / f2; Y[i][2] = scale * d2[i];
    }
    align3(X, Y, m->q, m->t);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* L2: refine_monodepth_relpose so@0x261030 (+ shared so@0x2592e0, varying so@0x260fa0):
 * lm_impl<MonoDepth*JacobianAccumulator> (bundle.cc / jacobian_impl.h).
 * cost = sum_k [ w_s*rho(r_sampson^2) + rho(s_r*|pi(R (d1+u) p1 + t) - x2|^2) [z>0]
 *                                      + rho(s_r*|pi(R^T (sigma (d2+v) p2 - t)) - x1|^2) [z>0] ]
 * (SURVEY.md §8a row L2).  Parameters: rotation tangent (3, R <- R*exp([w]x)), t (3),
 * scale, then (shift1, shift2) | f | (f1, f2).                                   */
static double loss_eval(int type, double thr, double r2) {
    const double t2 = thr * thr;
    switch (type) {
    case RO_LOSS_TRIVIAL: return r2;
    case RO_LOSS_TRUNCATED: return t2 < r2 ? t2 : r2; /* std::min(r2, t2): a NaN residual stays NaN */
    case RO_LOSS_HUBER: { double r = sqrt(r2); return r <= thr ? r2 : thr * (2.0 * r - thr); }
    case RO_LOSS_CAUCHY: return t2 * log1p(r2 / t2);
    case RO_LOSS_TRUNCATED_CAUCHY: return r2 > t2 ? t2 * log1p(1.0) : t2 * log1p(r2 / t2);
    default: return r2 < t2 ? r2 : t2;
    }
}
static double loss_weight(int type, double thr, double r2) {
    const double t2 = thr * thr;
    switch (type) {
    case RO_LOSS_TRIVIAL: return 1.0;
    case RO_LOSS_TRUNCATED: return r2 < t2 ? 1.0 : 0.0;
    case RO_LOSS_HUBER: { double r = sqrt(r2); return r <= thr ? 1.0 : thr / r; }
    case RO_LOSS_CAUCHY: return 1.0 / (1.0 + r2 / t2);
    case RO_LOSS_TRUNCATED_CAUCHY: return r2 > t2 ? 0.0 : 1.0 / (1.0 + r2 / t2);
    default: return r2 < t2 ? 1.0 : 0.0;
    }
}

static int n_params(int variant) {
    return variant == RO_CALIB ? 7 : (variant == RO_SHARED ? 8 : 9);
}

typedef struct {
    double rs;          /* sampson residual */
    double r12[2];      /* reprojection 1->2 */
    double r21[2];      /* reprojection 2->1 */
    int v12, v21;       /* positive-depth flags */
    double Js[9], J12[2][9], J21[2][9];
} pt_terms;

typedef struct {
    int variant;
    double R[9], E[9];
    ro_model m;
} lm_ctx;

static void lm_ctx_init(lm_ctx *c, int variant, const ro_model *m) {
    c->variant = variant;
    c->m = *m;
    quat_to_rotmat(m->q, c->R);
    ro_essential_from_motion(m->q, m->t, c->E);
}

static void point_terms(const lm_ctx *c, const double x1[2], const double x2[2], double d1, double d2,
                        int want_jac, pt_terms *o) {
    const int variant = c->variant;
    const double *R = c->R, *E = c->E;
    const ro_model *m = &c->m;
    const double f1 = m->f1, f2 = m->f2;
    const double p1[3] = {x1[0] / f1, x1[1] / f1, 1.0};
    const double p2[3] = {x2[0] / f2, x2[1] / f2, 1.0};
    const int np = n_params(variant);
    if (want_jac) memset(o->Js, 0, sizeof(o->Js) + sizeof(o->J12) + sizeof(o->J21));
    /* --- Sampson --- */
    double Ep1[3], Etp2[3];
    matvec3(E, p1, Ep1);
    matTvec3(E, p2, Etp2);
    const double C = dot3(p2, Ep1);
    const double A = Ep1[0] * Ep1[0] + Ep1[1] * Ep1[1];
    const double B = Etp2[0] * Etp2[0] + Etp2[1] * Etp2[1];
    const double if1sq = 1.0 / (f1 * f1), if2sq = 1.0 / (f2 * f2);
    const double den = A * if2sq + B * if1sq;
    const double inv = 1.0 / sqrt(den);
    o->rs = C * inv;
    if (want_jac) {
        const double k = 0.5 * C * inv * inv * inv;
        double Rp1[3];
        matvec3(R, p1, Rp1);
        for (int i = 0; i < 3; ++i) {
            double e[3] = {0, 0, 0}; e[i] = 1.0;
            double exp1[3], dEp1[3], dEtp2[3], tmp[3];
            /* rotation */
            cross3(e, p1, exp1);
            matvec3(E, exp1, dEp1);
            cross3(e, Etp2, tmp);
            dEtp2[0] = -tmp[0]; dEtp2[1] = -tmp[1];
            double dC = dot3(Etp2, exp1);
            double dden = 2.0 * (Ep1[0] * dEp1[0] + Ep1[1] * dEp1[1]) * if2sq +
                          2.0 * (Etp2[0] * dEtp2[0] + Etp2[1] * dEtp2[1]) * if1sq;
            o->Js[i] = dC * inv - k * dden;
            /* translation */
            cross3(e, Rp1, dEp1);
            cross3(e, p2, tmp);
            double rt[3];
            matTvec3(R, tmp, rt);
            dEtp2[0] = -rt[0]; dEtp2[1] = -rt[1];
            dC = dot3(p2, dEp1);
            dden = 2.0 * (Ep1[0] * dEp1[0] + Ep1[1] * dEp1[1]) * if2sq +
                   2.0 * (Etp2[0] * dEtp2[0] + Etp2[1] * dEtp2[1]) * if1sq;
            o->Js[3 + i] = dC * inv - k * dden;
        }
        if (variant == RO_SHARED || variant == RO_VARYING) {
            const double dp1[3] = {-p1[0] / f1, -p1[1] / f1, 0.0};
            const double dp2[3] = {-p2[0] / f2, -p2[1] / f2, 0.0};
            double dEp1[3], dEtp2[3];
            matvec3(E, dp1, dEp1);
            matTvec3(E, dp2, dEtp2);
            double dC1 = dot3(Etp2, dp1);
            double dden1 = 2.0 * (Ep1[0] * dEp1[0] + Ep1[1] * dEp1[1]) * if2sq - 2.0 * B * if1sq / f1;
            double dC2 = dot3(Ep1, dp2);
            double dden2 = 2.0 * (Etp2[0] * dEtp2[0] + Etp2[1] * dEtp2[1]) * if1sq - 2.0 * A * if2sq / f2;
            double j1 = dC1 * inv - k * dden1, j2 = dC2 * inv - k * dden2;
            if (variant == RO_SHARED) o->Js[7] = j1 + j2;
            else { o->Js[7] = j1; o->Js[8] = j2; }
        }
    }
    /* --- reprojection 1 -> 2 --- */
    {
        const double a = d1 + m->shift1;
        const double P[3] = {a * p1[0], a * p1[1], a * p1[2]};
        double Z[3];
        matvec3(R, P, Z);
        Z[0] += m->t[0]; Z[1] += m->t[1]; Z[2] += m->t[2];
        o->v12 = Z[2] > 0.0;
        const double iz = 1.0 / Z[2];
        const double u0 = Z[0] * iz, u1 = Z[1] * iz;
        o->r12[0] = f2 * u0 - x2[0];
        o->r12[1] = f2 * u1 - x2[1];
        if (want_jac && o->v12) {
            const double g = f2 * iz;
#define PROJ_JAC(dZ, col)                                   \
    do {                                                    \
        o->J12[0][col] = g * ((dZ)[0] - u0 * (dZ)[2]);      \
        o->J12[1][col] = g * ((dZ)[1] - u1 * (dZ)[2]);      \
    } while (0)
            for (int i = 0; i < 3; ++i) {
                double e[3] = {0, 0, 0}; e[i] = 1.0;
                double exP[3], dZ[3];
                cross3(e, P, exP);
                matvec3(R, exP, dZ);
                PROJ_JAC(dZ, i);
                PROJ_JAC(e, 3 + i);
            }
            if (variant == RO_CALIB_SHIFT) {
                double dZ[3];
                matvec3(R, p1, dZ);
                PROJ_JAC(dZ, 7);
            }
            if (variant == RO_SHARED || variant == RO_VARYING) {
                const double dP[3] = {-a * p1[0] / f1, -a * p1[1] / f1, 0.0};
                double dZ[3];
                matvec3(R, dP, dZ);
                PROJ_JAC(dZ, 7);
                int c2 = variant == RO_SHARED ? 7 : 8;
                o->J12[0][c2] += u0;
                o->J12[1][c2] += u1;
            }
#undef PROJ_JAC
        }
    }
    /* --- reprojection 2 -> 1 --- */
    {
        const double bb = d2 + m->shift2;
        const double b = m->scale * bb;
        const double Q[3] = {b * p2[0] - m->t[0], b * p2[1] - m->t[1], b * p2[2] - m->t[2]};
        double Y[3];
        matTvec3(R, Q, Y);
        o->v21 = Y[2] > 0.0;
        const double iz = 1.0 / Y[2];
        const double u0 = Y[0] * iz, u1 = Y[1] * iz;
        o->r21[0] = f1 * u0 - x1[0];
        o->r21[1] = f1 * u1 - x1[1];
        if (want_jac && o->v21) {
            const double g = f1 * iz;
#define PROJ_JAC(dY, col)                                   \
    do {                                                    \
        o->J21[0][col] = g * ((dY)[0] - u0 * (dY)[2]);      \
        o->J21[1][col] = g * ((dY)[1] - u1 * (dY)[2]);      \
    } while (0)
            for (int i = 0; i < 3; ++i) {
                double e[3] = {0, 0, 0}; e[i] = 1.0;
                double dY[3];
                cross3(Y, e, dY);
                PROJ_JAC(dY, i);
                double dT[3] = {-R[3 * i], -R[3 * i + 1], -R[3 * i + 2]};
                PROJ_JAC(dT, 3 + i);
            }
            {
                const double dQ[3] = {bb * p2[0], bb * p2[1], bb * p2[2]};
                double dY[3];
                matTvec3(R, dQ, dY);
                PROJ_JAC(dY, 6);
            }
            if (variant == RO_CALIB_SHIFT) {
                const double dQ[3] = {m->scale * p2[0], m->scale * p2[1], m->scale * p2[2]};
                double dY[3];
                matTvec3(R, dQ, dY);
                PROJ_JAC(dY, 8);
            }
            if (variant == RO_SHARED || variant == RO_VARYING) {
                const double dQ[3] = {-b * p2[0] / f2, -b * p2[1] / f2, 0.0};
                double dY[3];
                matTvec3(R, dQ, dY);
                int c2 = variant == RO_SHARED ? 7 : 8;
                o->J21[0][c2] += g * (dY[0] - u0 * dY[2]);
                o->J21[1][c2] += g * (dY[1] - u1 * dY[2]);
                o->J21[0][7] += u0;
                o->J21[1][7] += u1;
            }
#undef PROJ_JAC
        }
    }
    (void)np;
}

RO_API double ro_cost(int variant, const double *x1, const double *x2, const double *d1, const double *d2,
                      size_t n, const ro_model *m, double scale_reproj, double weight_sampson, int loss_type,
                      double loss_scale) {
    lm_ctx c;
    lm_ctx_init(&c, variant, m);
    double cost = 0.0;
    pt_terms o;
    for (size_t k = 0; k < n; ++k) {
        point_terms(&c, x1 + 2 * k, x2 + 2 * k, d1[k], d2[k], 0, &o);
        if (weight_sampson > 0.0) cost += weight_sampson * loss_eval(loss_type, loss_scale, o.rs * o.rs);
        if (scale_reproj > 0.0) {
            if (o.v12) cost += loss_eval(loss_type, loss_scale, scale_reproj * (o.r12[0] * o.r12[0] + o.r12[1] * o.r12[1]));
            if (o.v21) cost += loss_eval(loss_type, loss_scale, scale_reproj * (o.r21[0] * o.r21[0] + o.r21[1] * o.r21[1]));
        }
    }
    return cost;
}

/* JtJ: 9x9 row-major (only the leading np x np block is used), Jtr: 9 */
RO_API void ro_accumulate(int variant, const double *x1, const double *x2, const double *d1, const double *d2,
                          size_t n, const ro_model *m, double scale_reproj, double weight_sampson,
                          int loss_type, double loss_scale, double *JtJ, double *Jtr) {
    lm_ctx c;
    lm_ctx_init(&c, variant, m);
    const int np = n_params(variant);
    memset(JtJ, 0, 81 * sizeof(double));
    memset(Jtr, 0, 9 * sizeof(double));
    pt_terms o;
    for (size_t k = 0; k < n; ++k) {
        point_terms(&c, x1 + 2 * k, x2 + 2 * k, d1[k], d2[k], 1, &o);
        if (weight_sampson > 0.0) {
            /* the binary scales the Sampson residual AND its Jacobian row by weight_sampson before the rank-1
             * update, i.e. the normal equations carry weight_sampson^2 while the cost (ro_cost) carries
             * weight_sampson — verified against refine_monodepth_relpose so@0x261030 with weights 0.25/0.5/2
             * and all losses (one damped step reproduced to 1e-16 only with the square).  Invisible at the
             * default weight 1. */
            /* ... and the two focal accumulators (so@0x2592e0 / so@0x260fa0) evaluate the robust weight at
             * weight_sampson * r^2 where the calibrated one (so@0x261030) uses r^2 (same experiment, CAUCHY /
             * HUBER / TRUNCATED_CAUCHY at weights 0.5 and 2: 1e-15 only with this argument). */
            const double warg = (variant == RO_SHARED || variant == RO_VARYING) ? weight_sampson * (o.rs * o.rs) : o.rs * o.rs;
            double w = weight_sampson * weight_sampson * loss_weight(loss_type, loss_scale, warg);
            if (w != 0.0)
                for (int i = 0; i < np; ++i) {
                    Jtr[i] += w * o.Js[i] * o.rs;
                    for (int j = 0; j <= i; ++j) JtJ[9 * i + j] += w * o.Js[i] * o.Js[j];
                }
        }
        if (scale_reproj > 0.0) {
            if (o.v12) {
                double r2 = scale_reproj * (o.r12[0] * o.r12[0] + o.r12[1] * o.r12[1]);
                double w = scale_reproj * loss_weight(loss_type, loss_scale, r2);
                if (w != 0.0)
                    for (int i = 0; i < np; ++i) {
                        Jtr[i] += w * (o.J12[0][i] * o.r12[0] + o.J12[1][i] * o.r12[1]);
                        for (int j = 0; j <= i; ++j)
                            JtJ[9 * i + j] += w * (o.J12[0][i] * o.J12[0][j] + o.J12[1][i] * o.J12[1][j]);
                    }
            }
            if (o.v21) {
                double r2 = scale_reproj * (o.r21[0] * o.r21[0] + o.r21[1] * o.r21[1]);
                double w = scale_reproj * loss_weight(loss_type, loss_scale, r2);
                if (w != 0.0)
                    for (int i = 0; i < np; ++i) {
                        Jtr[i] += w * (o.J21[0][i] * o.r21[0] + o.J21[1][i] * o.r21[1]);
                        for (int j = 0; j <= i; ++j)
                            JtJ[9 * i + j] += w * (o.J21[0][i] * o.J21[0][j] + o.J21[1][i] * o.J21[1][j]);
                    }
            }
        }
    }
    for (int i = 0; i < np; ++i)
        for (int j = i + 1; j < np; ++j) JtJ[9 * i + j] = JtJ[9 * j + i];
}

/* Cholesky (lower) solve of A x = b on the leading np x np block; NaN on failure (as Eigen LLT). */
static void llt_solve(const double *A9, const double *b, int np, double *x) {
    double L[81];
    memset(L, 0, sizeof(L));
    for (int j = 0; j < np; ++j) {
        double s = A9[9 * j + j];
        for (int k = 0; k < j; ++k) s -= L[9 * j + k] * L[9 * j + k];
        double d = sqrt(s);
        L[9 * j + j] = d;
        for (int i = j + 1; i < np; ++i) {
            double v = A9[9 * i + j];
            for (int k = 0; k < j; ++k) v -= L[9 * i + k] * L[9 * j + k];
            L[9 * i + j] = v / d;
        }
    }
    double y[9];
    for (int i = 0; i < np; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[9 * i + k] * y[k];
        y[i] = v / L[9 * i + i];
    }
    for (int i = np - 1; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < np; ++k) v -= L[9 * k + i] * x[k];
        x[i] = v / L[9 * i + i];
    }
}

static void model_step(int variant, const ro_model *m, const double *dp, ro_model *out) {
    *out = *m;
    double dq[4], qn[4];
    quat_exp(dp, dq);
    quat_multiply(m->q, dq, qn);
    memcpy(out->q, qn, 32);
    out->t[0] += dp[3]; out->t[1] += dp[4]; out->t[2] += dp[5];
    out->scale += dp[6];
    if (variant == RO_CALIB_SHIFT) { out->shift1 += dp[7]; out->shift2 += dp[8]; }
    if (variant == RO_SHARED) { out->f1 += dp[7]; out->f2 = out->f1; }
    if (variant == RO_VARYING) { out->f1 += dp[7]; out->f2 += dp[8]; }
}

RO_API void ro_refine(int variant, const double *x1, const double *x2, const double *d1, const double *d2,
                      size_t n, ro_model *m, double scale_reproj, double weight_sampson,
                      const ro_bundle_opt *opt, ro_bundle_stats *stats) {
    const int np = n_params(variant);
    double JtJ[81], Jtr[9], sol[9];
    ro_bundle_stats st;
    st.cost = ro_cost(variant, x1, x2, d1, d2, n, m, scale_reproj, weight_sampson, opt->loss_type, opt->loss_scale);
    st.initial_cost = st.cost;
    st.grad_norm = -1;
    st.step_norm = -1;
    st.invalid_steps = 0;
    st.lambda = opt->initial_lambda;
    int recompute = 1;
    for (st.iterations = 0; st.iterations < opt->max_iterations; ++st.iterations) {
        if (recompute) {
            ro_accumulate(variant, x1, x2, d1, d2, n, m, scale_reproj, weight_sampson, opt->loss_type,
                          opt->loss_scale, JtJ, Jtr);
            double g = 0;
            for (int i = 0; i < np; ++i) g += Jtr[i] * Jtr[i];
            st.grad_norm = sqrt(g);
            if (st.grad_norm < opt->gradient_tol) break;
        }
        for (int k = 0; k < np; ++k) JtJ[9 * k + k] += st.lambda;
        llt_solve(JtJ, Jtr, np, sol);
        double sn = 0;
        for (int i = 0; i < np; ++i) { sol[i] = -sol[i]; sn += sol[i] * sol[i]; }
        st.step_norm = sqrt(sn);
        if (st.step_norm < opt->step_tol) break;
        ro_model mn;
        model_step(variant, m, sol, &mn);
        double cost_new = ro_cost(variant, x1, x2, d1, d2, n, &mn, scale_reproj, weight_sampson,
                                  opt->loss_type, opt->loss_scale);
        if (cost_new < st.cost) {
            *m = mn;
            st.lambda = fmax(opt->min_lambda, st.lambda / 10);
            st.cost = cost_new;
            recompute = 1;
        } else {
            st.invalid_steps++;
            for (int k = 0; k < np; ++k) JtJ[9 * k + k] -= st.lambda;
            st.lambda = fmin(opt->max_lambda, st.lambda * 10);
            recompute = 0;
        }
    }
    if (stats) *stats = st;
}

/* ------------------------------------------------------------------------- */
/* R1: ransac<Est,Model> so@0x22f030 + score_models<> so@0x22ebc0 (ransac_impl.h),
 * estimators of robust/estimators/relative_pose.cc.            */
RO_API int ro_solve_shared_focal(const double *x1h, const double *x2h, const double *d1, const double *d2,
                                 ro_model *out);

typedef struct {
    int variant;
    size_t n;
    const double *x1, *x2, *d1, *d2;
    const ro_ransac_opt *opt;
    ro_sampler sampler;
    double sq_thr, scale_reproj;
    ro_bundle_opt lo;
} estimator;

static int est_generate(estimator *e, ro_model *models) {
    size_t s[3];
    ro_sampler_generate(&e->sampler, s);
    double x1h[9], x2h[9], d1[3], d2[3];
    for (int i = 0; i < 3; ++i) {
        x1h[3 * i] = e->x1[2 * s[i]]; x1h[3 * i + 1] = e->x1[2 * s[i] + 1]; x1h[3 * i + 2] = 1.0;
        x2h[3 * i] = e->x2[2 * s[i]]; x2h[3 * i + 1] = e->x2[2 * s[i] + 1]; x2h[3 * i + 2] = 1.0;
        d1[i] = e->d1[s[i]];
        d2[i] = e->d2[s[i]];
    }
    switch (e->variant) {
    case RO_CALIB: return ro_solve_calib_scale(x1h, x2h, d1, d2, models);
    case RO_CALIB_SHIFT: return ro_solve_calib_shift(x1h, x2h, d1, d2, models);
    case RO_SHARED: return ro_solve_shared_focal(x1h, x2h, d1, d2, models);
    default: return ro_solve_varying_focal(x1h, x2h, d1, d2, models);
    }
}

static double est_score(const estimator *e, const ro_model *m, size_t *count) {
    if (e->variant == RO_CALIB || e->variant == RO_CALIB_SHIFT)
        return ro_msac_score_pose(m->q, m->t, e->x1, e->x2, e->n, e->sq_thr, count);
    double F[9];
    ro_fundamental_from_model(m, F);
    return ro_msac_score_F(F, e->x1, e->x2, e->n, e->sq_thr, count);
}

static void est_refine(const estimator *e, ro_model *m) {
    ro_refine(e->variant, e->x1, e->x2, e->d1, e->d2, e->n, m, e->scale_reproj, e->opt->weight_sampson,
              &e->lo, NULL);
}

RO_API void ro_ransac(int variant, const double *x1, const double *x2, const double *d1, const double *d2,
                      size_t n, const ro_ransac_opt *opt, ro_model *best, ro_ransac_stats *stats,
                      char *inliers) {
    estimator e;
    e.variant = variant; e.n = n; e.x1 = x1; e.x2 = x2; e.d1 = d1; e.d2 = d2; e.opt = opt;
    e.sq_thr = opt->max_epipolar_error * opt->max_epipolar_error;
    e.scale_reproj = opt->max_reproj_error > 0.0
                         ? (opt->max_epipolar_error * opt->max_epipolar_error) /
                               (opt->max_reproj_error * opt->max_reproj_error)
                         : 0.0;
    e.lo.max_iterations = 25; e.lo.loss_type = RO_LOSS_TRUNCATED;
    /* VaryingFocalMonodepthPoseEstimator::refine_model so@0x4fb0a0 never copies the threshold into its
     * BundleOptions: loss_scale stays at the struct default 1.0 (so@0x5232a0); the calibrated (so@0x4fa550)
     * and shared-focal (so@0x4fad60) estimators use opt.max_epipolar_error.  Reproduced as is. */
    e.lo.loss_scale = variant == RO_VARYING ? 1.0 : opt->max_epipolar_error;
    e.lo.gradient_tol = 1e-10; e.lo.step_tol = 1e-8; e.lo.initial_lambda = 1e-3;
    e.lo.min_lambda = 1e-10; e.lo.max_lambda = 1e10;

    model_init(best);
    stats->refinements = 0; stats->iterations = 0; stats->num_inliers = 0;
    stats->inlier_ratio = 0.0; stats->model_score = DBL_MAX;
    if (inliers) memset(inliers, 0, n);
    if (n < 3) return;
    ro_sampler_init(&e.sampler, n, 3, opt->seed, opt->progressive_sampling, (size_t)opt->max_prosac_iterations);

    size_t best_minimal_inliers = 0;
    double best_minimal_score = DBL_MAX;
    double dyn_max_iter = (double)opt->max_iterations;
    const double log_prob_missing = log(1.0 - opt->success_prob);
    ro_model models[8];
    size_t cnt = 0;
    for (stats->iterations = 0; stats->iterations < opt->max_iterations; stats->iterations++) {
        if (stats->iterations > opt->min_iterations && (double)stats->iterations > dyn_max_iter) break;
        int nm = est_generate(&e, models);
        int best_idx = -1;
        for (int i = 0; i < nm; ++i) {
            double score = est_score(&e, &models[i], &cnt);
            int more = cnt > best_minimal_inliers;
            int better = score < best_minimal_score;
            if (more || better) {
                if (more) best_minimal_inliers = cnt;
                if (better) best_minimal_score = score;
                best_idx = i;
                if (score < stats->model_score) {
                    stats->model_score = score;
                    *best = models[i];
                    stats->num_inliers = (int64_t)cnt;
                }
            }
        }
        if (best_idx == -1) continue;
        ro_model refined = models[best_idx];
        est_refine(&e, &refined);
        stats->refinements++;
        double rs = est_score(&e, &refined, &cnt);
        if (rs < stats->model_score) {
            stats->model_score = rs;
            stats->num_inliers = (int64_t)cnt;
            *best = refined;
        }
        stats->inlier_ratio = (double)stats->num_inliers / (double)n;
        if (stats->inlier_ratio >= 0.9999) dyn_max_iter = (double)opt->min_iterations;
        else if (stats->inlier_ratio <= 0.0001) dyn_max_iter = (double)opt->max_iterations;
        else {
            const double prob_outlier = 1.0 - pow(stats->inlier_ratio, 3.0);
            dyn_max_iter = ceil(log_prob_missing / log(prob_outlier) * opt->dyn_num_trials_mult);
        }
    }
    /* final refinement: unlike the LO inside the loop it does NOT write stats.model_score (the reported score is
     * the one before this step) — seen on the binary: with no minimal model at all model_score stays DBL_MAX
     * while num_inliers / the model are taken from the refined identity, and one-iteration runs report the
     * pre-refinement score. */
    ro_model refined = *best;
    est_refine(&e, &refined);
    stats->refinements++;
    double rs = est_score(&e, &refined, &cnt);
    if (rs < stats->model_score) {
        stats->num_inliers = (int64_t)cnt;
        *best = refined;
    }
    stats->inlier_ratio = (double)stats->num_inliers / (double)n;
    free(e.sampler.growth);
    if (inliers) {
        if (variant == RO_CALIB || variant == RO_CALIB_SHIFT)
            ro_get_inliers_pose(best->q, best->t, x1, x2, n, e.sq_thr, inliers);
        else {
            double F[9];
            ro_fundamental_from_model(best, F);
            ro_get_inliers_F(F, x1, x2, n, e.sq_thr, inliers);
        }
    }
}

/* P1/P2: estimate_monodepth_relative_pose so@0x224170, estimate_shared_focal_… so@0x223300,
 * estimate_varying_focal_… so@0x223a40 (robust.cc).  cam = (fx, fy, cx, cy) for the
 * calibrated variants (SIMPLE_PINHOLE: fx = fy); x in pixels.  Focal variants take
 * principal-point-centred pixels and ignore cam1/cam2.                              */
RO_API void ro_estimate(int variant, const double *x1, const double *x2, const double *d1, const double *d2,
                        size_t n, const double *cam1, const double *cam2, const ro_ransac_opt *ropt,
                        const ro_bundle_opt *bopt, ro_model *best, ro_ransac_stats *stats, char *inliers) {
    double *a = (double *)malloc(sizeof(double) * (4 * n + 4));
    double *b = a + 2 * n;
    ro_ransac_opt ro = *ropt;
    ro_bundle_opt bo = *bopt;
    double nscale = 1.0;
    if (variant == RO_CALIB || variant == RO_CALIB_SHIFT) {
        for (size_t k = 0; k < n; ++k) {
            a[2 * k] = (x1[2 * k] - cam1[2]) / cam1[0];
            a[2 * k + 1] = (x1[2 * k + 1] - cam1[3]) / cam1[1];
            b[2 * k] = (x2[2 * k] - cam2[2]) / cam2[0];
            b[2 * k + 1] = (x2[2 * k + 1] - cam2[3]) / cam2[1];
        }
        const double fo1 = 0.5 * (cam1[0] + cam1[1]), fo2 = 0.5 * (cam2[0] + cam2[1]);
        const double k = 0.5 * (1.0 / fo1 + 1.0 / fo2);
        ro.max_epipolar_error = ropt->max_epipolar_error * k;
        ro.max_reproj_error = ropt->max_reproj_error * k;
        bo.loss_scale = 0.5 * ro.max_epipolar_error;
    } else {
        /* normalize_points so@0x4f6ae0: shared scale, no centroid */
        double s = 0.0;
        for (size_t k = 0; k < n; ++k) {
            s += sqrt(x1[2 * k] * x1[2 * k] + x1[2 * k + 1] * x1[2 * k + 1]);
            s += sqrt(x2[2 * k] * x2[2 * k] + x2[2 * k + 1] * x2[2 * k + 1]);
        }
        nscale = s / (sqrt(2.0) * (double)n);
        for (size_t k = 0; k < 2 * n; ++k) { a[k] = x1[k] / nscale; b[k] = x2[k] / nscale; }
        ro.max_epipolar_error = ropt->max_epipolar_error / nscale;
        ro.max_reproj_error = ropt->max_reproj_error / nscale;
        bo.loss_scale = bopt->loss_scale / nscale;
    }
    ro_ransac(variant, a, b, d1, d2, n, &ro, best, stats, inliers);
    if (stats->num_inliers > 3) {
        size_t m = 0;
        double *ia = (double *)malloc(sizeof(double) * (6 * n));
        double *ib = ia + 2 * n, *id1 = ia + 4 * n, *id2 = ia + 5 * n;
        for (size_t k = 0; k < n; ++k)
            if (inliers[k]) {
                ia[2 * m] = a[2 * k]; ia[2 * m + 1] = a[2 * k + 1];
                ib[2 * m] = b[2 * k]; ib[2 * m + 1] = b[2 * k + 1];
                id1[m] = d1[k]; id2[m] = d2[k];
                ++m;
            }
        const double sr = ro.max_reproj_error > 0.0
                              ? (ro.max_epipolar_error * ro.max_epipolar_error) /
                                    (ro.max_reproj_error * ro.max_reproj_error)
                              : 0.0;
        ro_refine(variant, ia, ib, id1, id2, m, best, sr, ropt->weight_sampson, &bo, NULL);
        free(ia);
    }
    if (variant == RO_SHARED || variant == RO_VARYING) {
        best->f1 *= nscale;
        best->f2 *= nscale;
    }
    free(a);
}

/* ------------------------------------------------------------------------- */
/* S3: relpose_monodepth_3pt_shared_focal so@0x18fdf0.  Unknowns f, scale and the depth
 * of the third point in camera 2: points 0,1 align in 3-D,
 * point 2 only reprojects.  With g = 1/f^2, nu = depth_2/scale and
 *   S1 = g*A1_01 + B1_01,  S2 = g*A2_01 + B2_01,  scale^2 = S1/S2,
 * the difference of the (0,2) and (1,2) distance equations is linear in nu
 * (nu = N/(2 S1 L)), and substituting into the (0,2) equation leaves a quintic in g
 * whose constant term vanishes identically (g = 0 <=> f = inf), i.e. a quartic:
 * <= 4 solutions, as the binary's 4x4 action matrix.  Valid iff g > 0 and nu > 0.
 * The binary orders its solutions by its eigen-solver; here they come in the order of
 * solve_quartic_real (documented deviation; affects only same-iteration RANSAC ties). */
typedef struct { double c[8]; int deg; } poly;
static poly pmake(int deg, const double *c) { poly p; memset(&p, 0, sizeof p); p.deg = deg; for (int i = 0; i <= deg; ++i) p.c[i] = c[i]; return p; }
static poly plin(double c1, double c0) { double c[2] = {c0, c1}; return pmake(1, c); } /* c1*g + c0 */
static poly pmul(poly a, poly b) {
    poly r; memset(&r, 0, sizeof r); r.deg = a.deg + b.deg;
    for (int i = 0; i <= a.deg; ++i) for (int j = 0; j <= b.deg; ++j) r.c[i + j] += a.c[i] * b.c[j];
    return r;
}
static poly padd(poly a, double sa, poly b, double sb) {
    poly r; memset(&r, 0, sizeof r); r.deg = a.deg > b.deg ? a.deg : b.deg;
    for (int i = 0; i <= r.deg; ++i) r.c[i] = sa * (i <= a.deg ? a.c[i] : 0.0) + sb * (i <= b.deg ? b.c[i] : 0.0);
    return r;
}
static double peval(poly p, double x) { double v = 0; for (int i = p.deg; i >= 0; --i) v = v * x + p.c[i]; return v; }

RO_API int ro_solve_shared_focal(const double *x1h, const double *x2h, const double *d1, const double *d2,
                                 ro_model *out) {
    double A1[3], B1[3]; /* pairs 01, 02, 12 */
    static const int PI[3] = {0, 0, 1}, PJ[3] = {1, 2, 2};
    for (int r = 0; r < 3; ++r) {
        int i = PI[r], j = PJ[r];
        double vx = d1[i] * x1h[3 * i] - d1[j] * x1h[3 * j], vy = d1[i] * x1h[3 * i + 1] - d1[j] * x1h[3 * j + 1];
        A1[r] = vx * vx + vy * vy;
        B1[r] = (d1[i] - d1[j]) * (d1[i] - d1[j]);
    }
    double vx = d2[0] * x2h[0] - d2[1] * x2h[3], vy = d2[0] * x2h[1] - d2[1] * x2h[4];
    const double A2 = vx * vx + vy * vy, B2 = (d2[0] - d2[1]) * (d2[0] - d2[1]);
    const double n2 = x2h[6] * x2h[6] + x2h[7] * x2h[7];
    const double c02 = d2[0] * (x2h[0] * x2h[6] + x2h[1] * x2h[7]);
    const double c12 = d2[1] * (x2h[3] * x2h[6] + x2h[4] * x2h[7]);
    const double m0 = d2[0] * d2[0] * (x2h[0] * x2h[0] + x2h[1] * x2h[1]);
    const double m1 = d2[1] * d2[1] * (x2h[3] * x2h[3] + x2h[4] * x2h[4]);
    poly S1 = plin(A1[0], B1[0]), S2 = plin(A2, B2);
    poly T02 = plin(A1[1], B1[1]), T12 = plin(A1[2], B1[2]);
    poly L = plin(c02 - c12, d2[0] - d2[1]);
    poly N = padd(pmul(S1, plin(m0 - m1, d2[0] * d2[0] - d2[1] * d2[1])), 1.0, pmul(S2, padd(T02, 1.0, T12, -1.0)), -1.0);
    poly LL = pmul(L, L);
    poly lhs = pmul(pmul(pmul(S2, T02), S1), LL);                       /* *4 */
    poly r1 = pmul(pmul(N, N), plin(n2, 1.0));
    poly r2 = pmul(pmul(pmul(S1, N), L), plin(c02, d2[0]));             /* *4 */
    poly r3 = pmul(pmul(pmul(S1, S1), LL), plin(m0, d2[0] * d2[0]));    /* *4 */
    poly P = padd(padd(lhs, 4.0, r1, -1.0), 1.0, padd(r2, 4.0, r3, -4.0), 1.0);
    /* P.c[0] == 0 analytically: quartic P.c[5] g^4 + ... + P.c[1] */
    double roots[4];
    int nr = ro_solve_quartic_real(P.c[4] / P.c[5], P.c[3] / P.c[5], P.c[2] / P.c[5], P.c[1] / P.c[5], roots);
    int n = 0;
    for (int ir = 0; ir < nr; ++ir) {
        double g = roots[ir];
        /* Newton polish on the quartic */
        for (int it = 0; it < 3; ++it) {
            double v = (((P.c[5] * g + P.c[4]) * g + P.c[3]) * g + P.c[2]) * g + P.c[1];
            double dv = ((4.0 * P.c[5] * g + 3.0 * P.c[4]) * g + 2.0 * P.c[3]) * g + P.c[2];
            if (dv == 0.0) break;
            g -= v / dv;
        }
        if (!(g > 0.0)) continue;
        const double s1 = peval(S1, g), s2 = peval(S2, g);
        const double sc2 = s1 / s2;
        if (!(sc2 > 0.0)) continue;
        const double nu = peval(N, g) / (2.0 * s1 * peval(L, g));
        if (!(nu > 0.0)) continue;
        const double w = sqrt(g), f = 1.0 / w, scale = sqrt(sc2);
        ro_model *m = &out[n];
        model_init(m);
        m->scale = scale; m->f1 = f; m->f2 = f;
        double X[3][3], Y[3][3];
        for (int i = 0; i < 3; ++i) {
            X[i][0] = d1[i] * x1h[3 * i] * w; X[i][1] = d1[i] * x1h[3 * i + 1] * w; X[i][2] = d1[i];
            double dep = scale * (i < 2 ? d2[i] : nu);
            Y[i][0] = dep * x2h[3 * i] * w; Y[i][1] = dep * x2h[3 * i + 1] * w; Y[i][2] = dep;
        }
        align3(X, Y, m->q, m->t);
        ++n;
    }
    return n;
}
