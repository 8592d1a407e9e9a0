// tc_probe.cu — stand-alone check + micro-benchmark of the tensor-core count tier (mdrp_b200/csrc/rp_tc.cuh).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -I mdrp_b200/csrc \
//        -o tools/tc_probe.bin tools/tc_probe.cu
//   tools/tc_probe.bin [pairs=64] [n=2000] [models_per_pair=6600] [reps=5]
//
// 1. correctness: raw (Cs, Ts) of the first work item against FP64 (error in units of the eps_tc / d_den bounds of
//    the kernel header), and for EVERY model  out <= #{k : r2_k >= thr^2}  (rigor) plus how sharp the count is;
// 2. speed: point-scores per second of the kernel alone (CUDA events).
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "rp_tc.cuh"
#include "rp_score.cuh"

using namespace rp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static Quat quat_from_axis_angle(double ax, double ay, double az, double ang) {
    const double n = std::sqrt(ax * ax + ay * ay + az * az);
    Quat q;
    q.w = std::cos(ang / 2);
    const double s = std::sin(ang / 2) / n;
    q.x = ax * s; q.y = ay * s; q.z = az * s;
    return q;
}

int main(int argc, char **argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 64;
    const int n = argc > 2 ? atoi(argv[2]) : 2000;
    const int MPP = argc > 3 ? atoi(argv[3]) : 6600;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    const int nseg = (MPP + 4 * SEG - 1) / (4 * SEG) < 1 ? 1 : std::max(1, (MPP + 659) / 660);  // ~660 models per segment like cfg2
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> N01(0.0, 1.0);
    std::uniform_real_distribution<double> U01(0.0, 1.0);
    const double f = 800.0, thr = 2.0 * (1.0 / f);

    // ---- scenes: same model as mdrp_b200/synth.py ----
    std::vector<float4> pts32((size_t)P * n);
    std::vector<double> pts64((size_t)P * n * 4);
    std::vector<PairParams> pairs(P);
    std::vector<Model> models((size_t)P * nseg * 4 * SEG);
    std::vector<int> seg_count((size_t)P * nseg, 0), item_prefix(P + 1, 0);
    memset(models.data(), 0, models.size() * sizeof(Model));
    for (int p = 0; p < P; ++p) {
        const Quat q = quat_from_axis_angle(N01(rng), N01(rng), N01(rng), 0.35 * (0.2 + 0.8 * U01(rng)));
        const M3 R = quat_to_rotmat(q);
        double t[3] = {N01(rng), N01(rng), N01(rng)};
        const double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        for (double &v : t) v *= 0.5 / tn;
        double Mmax = 0, mmax = 0;
        for (int k = 0; k < n; ++k) {
            const double x = (U01(rng) * 1280 - 640) / f, y = (U01(rng) * 960 - 480) / f, z = 2 + 6 * U01(rng);
            const V3 X = v3(x * z, y * z, z);
            V3 Y = mul(R, X);
            Y = v3(Y.x + t[0], Y.y + t[1], Y.z + t[2]);
            double x2 = Y.x / Y.z + 0.5 / f * N01(rng), y2 = Y.y / Y.z + 0.5 / f * N01(rng);
            if (U01(rng) < 0.3) { x2 = (U01(rng) * 1280 - 640) / f; y2 = (U01(rng) * 960 - 480) / f; }
            const double x1 = x + 0.5 / f * N01(rng), y1 = y + 0.5 / f * N01(rng);
            double *o = &pts64[((size_t)p * n + k) * 4];
            o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
            pts32[(size_t)p * n + k] = make_float4((float)x1, (float)y1, (float)x2, (float)y2);
            const double m1 = fabs(x1) + fabs(y1) + 1, m2 = fabs(x2) + fabs(y2) + 1;
            Mmax = std::max(Mmax, m1 * m2);
            mmax = std::max(mmax, std::max(m1, m2));
        }
        PairParams pp;
        memset(&pp, 0, sizeof pp);
        pp.off = (long long)p * n; pp.n = n; pp.valid = 1; pp.thr = thr; pp.sq_thr = thr * thr;
        pp.Mmax = Mmax * 1.000001; pp.mmax = mmax * 1.000001;
        pairs[p] = pp;
        // models: the true pose perturbed by a log-uniform amount (good ... useless), a few NaN models
        int left = MPP;
        for (int s = 0; s < nseg; ++s) {
            const int c = std::min(left, (MPP + nseg - 1) / nseg);
            seg_count[(size_t)p * nseg + s] = c;
            left -= c;
            for (int i = 0; i < c; ++i) {
                const double mag = std::pow(10.0, -4.0 + 4.0 * U01(rng));
                Model m = identity_model();
                const Quat dq = quat_from_axis_angle(N01(rng), N01(rng), N01(rng), mag * U01(rng));
                m.q = quat_mul(dq, q);
                m.t = v3(t[0] + mag * N01(rng), t[1] + mag * N01(rng), t[2] + mag * N01(rng));
                if (U01(rng) < 0.01) m.t.x = NAN;
                models[((size_t)p * nseg + s) * 4 * SEG + i] = m;
            }
        }
        item_prefix[p + 1] = item_prefix[p] + (MPP + tc::TILE_MODELS - 1) / tc::TILE_MODELS;
    }
    const int n_items = item_prefix[P];

    // ---- device buffers ----
    float4 *d_pts32, *d_feat;
    PairParams *d_pairs;
    Model *d_models;
    int *d_seg, *d_pref, *d_nitems, *d_out;
    float *d_debug;
    unsigned long long *d_eval;
    const int debug_cols = std::min(n, 256);
    CK(cudaMalloc(&d_pts32, pts32.size() * 16));
    CK(cudaMalloc(&d_feat, pts32.size() * 128 + 128 * 64));
    CK(cudaMalloc(&d_pairs, pairs.size() * sizeof(PairParams)));
    CK(cudaMalloc(&d_models, models.size() * sizeof(Model)));
    CK(cudaMalloc(&d_seg, seg_count.size() * 4));
    CK(cudaMalloc(&d_pref, item_prefix.size() * 4));
    CK(cudaMalloc(&d_nitems, 4));
    CK(cudaMalloc(&d_out, models.size() * 4));
    CK(cudaMalloc(&d_debug, (size_t)tc::TILE_MODELS * 2 * debug_cols * 4));
    CK(cudaMalloc(&d_eval, 8));
    CK(cudaMemcpy(d_pts32, pts32.data(), pts32.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pairs, pairs.data(), pairs.size() * sizeof(PairParams), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_models, models.data(), models.size() * sizeof(Model), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_seg, seg_count.data(), seg_count.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pref, item_prefix.data(), item_prefix.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nitems, &n_items, 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, models.size() * 4));
    CK(cudaMemset(d_debug, 0, (size_t)tc::TILE_MODELS * 2 * debug_cols * 4));
    CK(cudaMemset(d_eval, 0, 8));

    const long long NP = (long long)P * n;
    tc::tc_features_kernel<<<(unsigned)((NP + 255) / 256), 256>>>(NP, d_pts32, d_feat);
    CK(cudaGetLastError());

    // ---- tensor map over the feature rows ----
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {32, (cuuint64_t)NP};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)tc::NT};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_feat, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 2; }

    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(tc::tc_count_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    tc::TcArgs a;
    memset(&a, 0, sizeof a);
    a.n_pairs = P; a.nseg = nseg; a.pairs = d_pairs; a.seg_count = d_seg; a.item_prefix = d_pref; a.n_items = d_nitems;
    a.models = d_models; a.out = d_out; a.pose = 1; a.evaluated = d_eval; a.debug = d_debug; a.debug_cols = debug_cols;
    const int grid = std::min(sms, n_items);
    tc::tc_count_kernel<0><<<grid, tc::THREADS, tc::SMEM_BYTES>>>(tmap, a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("kernel ran: %d items on %d CTAs, nseg %d\n", n_items, grid, nseg);

    // ---- 1a. raw accumulators of work item 0 against FP64 ----
    std::vector<float> dbg((size_t)tc::TILE_MODELS * 2 * debug_cols);
    std::vector<int> out(models.size());
    CK(cudaMemcpy(dbg.data(), d_debug, dbg.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    {
        double worstC = 0, worstT = 0, meanC = 0;
        long cntc = 0;
        int bad_print = 0;
        const PairParams pp = pairs[0];
        for (int r = 0; r < tc::TILE_MODELS && r < seg_count[0]; ++r) {
            const Model m = models[r];
            const M3 E = essential_from_motion(m.q, m.t);
            const double e[9] = {E.r0.x, E.r0.y, E.r0.z, E.r1.x, E.r1.y, E.r1.z, E.r2.x, E.r2.y, E.r2.z};
            double emax = 0;
            bool fin = true;
            for (double v : e) { emax = std::max(emax, fabs(v)); fin = fin && std::isfinite(v); }
            if (!fin) continue;
            const double u = 5.9604644775390625e-08;
            const double eps = tc::EPS_UNITS * u * emax * (pp.Mmax + thr * pp.mmax) * 1.0001;
            const double g = thr * thr * (1.0 + 1e-5), dden = 0.0078125 * emax * emax * pp.mmax * pp.mmax;
            int ex;
            (void)frexp((1.0 + 1.0 / tc::A_SPLIT) * eps * eps, &ex);
            const int sc = (48 - (ex - 1) + 1) / 2;
            for (int k = 0; k < debug_cols; ++k) {
                const double *x = &pts64[(size_t)k * 4];
                const double a0 = e[0] * x[0] + e[1] * x[1] + e[2], a1 = e[3] * x[0] + e[4] * x[1] + e[5], a2 = e[6] * x[0] + e[7] * x[1] + e[8];
                const double b0 = e[0] * x[2] + e[3] * x[3] + e[6], b1 = e[1] * x[2] + e[4] * x[3] + e[7];
                const double C = x[2] * a0 + x[3] * a1 + a2, den = a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1;
                const double Cs = dbg[((size_t)r * 2 + 0) * debug_cols + k], Ts = dbg[((size_t)r * 2 + 1) * debug_cols + k];
                const double errC = fabs(ldexp(Cs, -sc) - C) / eps;
                // Ts = -2^2s [(1+a) g (den~ + dden) + (1+1/a) eps^2] (1 + 2^-9 on the constant part): recover den~ - den in units of dden
                const double den_t = (-ldexp(Ts, -2 * sc) - (1.0 + 1.0 / tc::A_SPLIT) * eps * eps) / ((1.0 + tc::A_SPLIT) * g) - dden;
                const double errT = fabs(den_t - den) / dden;
                worstC = std::max(worstC, errC);
                worstT = std::max(worstT, errT);
                meanC += errC;
                ++cntc;
                if ((errC > 1.0 || errT > 1.5 || !std::isfinite(Cs)) && bad_print < 8) {
                    printf("  BAD row %d col %d: Cs %.9g (2^-s Cs %.9g) C %.9g errC/eps %.3g | den~ %.9g den %.9g errT/dden %.3g\n", r, k, Cs,
                           ldexp(Cs, -sc), C, errC, den_t, den, errT);
                    ++bad_print;
                }
            }
        }
        printf("raw accumulators (item 0, %ld values): max |C~-C|/eps_tc = %.4f (mean %.5f), max |den~-den|/d_den = %.4f  [must be < 1]\n",
               cntc, worstC, meanC / std::max(1L, cntc), worstT);
    }
    // ---- 1b. rigor and sharpness of the counts (first pairs only: the host loop is O(models x points)) ----
    {
        long violations = 0, checked = 0, nan_ok = 0, nan_bad = 0;
        double sum_out = 0, sum_true = 0;
        int printed = 0;
        const int PC = std::min(P, 2);
        for (int p = 0; p < PC; ++p)
            for (int s = 0; s < nseg; ++s)
                for (int i = 0; i < seg_count[(size_t)p * nseg + s]; ++i) {
                    const size_t slot = ((size_t)p * nseg + s) * 4 * SEG + i;
                    const Model m = models[slot];
                    const M3 E = essential_from_motion(m.q, m.t);
                    if (!std::isfinite(E.r0.x + E.r0.y + E.r0.z + E.r1.x + E.r1.y + E.r1.z + E.r2.x + E.r2.y + E.r2.z)) {
                        if (out[slot] == n) ++nan_ok; else ++nan_bad;
                        continue;
                    }
                    int outliers = 0;
                    for (int k = 0; k < n; ++k) {
                        const double *x = &pts64[((size_t)p * n + k) * 4];
                        const double r2 = sampson_r2_exact(E, x[0], x[1], x[2], x[3]);
                        outliers += !(r2 < thr * thr);
                    }
                    ++checked;
                    sum_out += out[slot];
                    sum_true += outliers;
                    if (out[slot] > outliers || out[slot] < 0) {
                        ++violations;
                        if (printed++ < 8) printf("  VIOLATION pair %d seg %d i %d: out %d > true outliers %d\n", p, s, i, out[slot], outliers);
                    }
                }
        printf("counts: %ld models checked, %ld violations (out > true outliers), NaN models %ld ok / %ld bad, sharpness sum(out)/sum(true) = %.5f\n",
               checked, violations, nan_ok, nan_bad, sum_out / std::max(1.0, sum_true));
    }
    // ---- 2. speed ----
    a.debug = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double ps = (double)P * MPP * n;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(e0));
            if (mode == 0) tc::tc_count_kernel<0><<<grid, tc::THREADS, tc::SMEM_BYTES>>>(tmap, a);
            else if (mode == 1) tc::tc_count_kernel<1><<<grid, tc::THREADS, tc::SMEM_BYTES>>>(tmap, a);
            else tc::tc_count_kernel<2><<<grid, tc::THREADS, tc::SMEM_BYTES>>>(tmap, a);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        printf("speed mode %d (%s): %.3f ms for %d pairs x %d models x %d points = %.3e point-scores -> %.3e point-scores/s (%.1f TFLOP/s tensor, 96 flop each)\n",
               mode, mode == 0 ? "full" : (mode == 1 ? "loads only" : "mma only"), best, P, MPP, n, ps, ps / (best * 1e-3), ps * 96 / (best * 1e-3) / 1e12);
    }
    return 0;
}
