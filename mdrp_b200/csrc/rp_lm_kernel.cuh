// rp_lm_kernel.cuh — batched Levenberg-Marquardt kernel (compiled in its own translation unit,
// repose_lm.cu, WITH FMA contraction: this path is not bit-pinned).
#pragma once
#include "rp_types.cuh"
#include "rp_lm.cuh"
#include <type_traits>

#ifndef RP_LM_UNROLL
#define RP_LM_UNROLL 1      // correspondences per thread in flight in the evaluation loop (build knob; measured: 1 is best)
#endif
#ifndef RP_LM_MIN_BLOCKS
#define RP_LM_MIN_BLOCKS 3   // blocks of 128 threads per SM, i.e. 12 warps: 168 registers per thread
#endif

namespace rp {

constexpr int LM_UNROLL = RP_LM_UNROLL;

// ---------------------------------------------------------------------------------------------
// lm: lm_impl<> of PoseLib bundle.cc, one block per problem (persistent over the problem list).
// JtJ + lambda I (lower LLT), accept when the cost decreases (lambda /= 10, recompute J) else
// lambda *= 10; stop on |J^T r| < gradient_tol or |step| < step_tol.
struct LMArgs {
    const int *prob_list;  // indices into models (null: identity)
    const int *n_prob;
    int prob_per_pair;     // pair = index / prob_per_pair
    const PairParams *pairs;
    const Pt64 *pts64;
    const double *d1, *d2;
    const unsigned char *mask;   // optional per-correspondence subset
    const int *enable;           // optional per-pair switch (final refinement: num_inliers > 3)
    Model *models;
    int use_final;               // 0: LO options (25 its, TRUNCATED, lo_loss_scale); 1: user bundle options
    int max_iterations, loss_type;
    double weight_sampson, gradient_tol, step_tol, initial_lambda, min_lambda, max_lambda;
    double loss_scale_override;  // >0: stage entry point passes its own loss scale / scale_reproj
    double scale_reproj_override;
    rp_bundle_stats *stats;      // optional, per problem index
    unsigned long long *lm_iters;
    unsigned long long *lm_flops;   // optional: FP64 flops executed (LM_FLOPS table x device counts)
    int *work_counter;           // zeroed before the launch: blocks take problems dynamically
    int warp_kernel;             // unmasked problems: one warp per problem (lm_warp_kernel) instead of one block
    int warp_single;             // ... also for the one-problem-per-pair launches without a problem list (final LO)
};

constexpr int LM_LIST_CAP = 16384;  // 32 KB of shared memory per block (3 blocks/SM); indices fit 16 bits

// Warp reduction of NV per-lane values as a reduce-scatter: at the step with lane mask O every lane keeps one
// half of its values and sends the other half to its partner, so the number of live values halves each step
// (NV/2 + NV/4 + ... ~ NV shuffles instead of 5 NV for NV butterfly sums).  Afterwards the warp total of
// value `base + j`, j < lim, sits in v[j] of exactly one lane (odd counts leave a zero pad slot in the upper
// half: `lim` keeps it from being mistaken for the neighbouring range's first value).
template <int MAXV, int CNT, int O>
struct ReduceScatter {
    static RP_D void run(double (&v)[MAXV], int lane, int &base, int &lim) {
        constexpr int H = (CNT + 1) / 2;
        const bool up = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double lo = v[i];
            const double hi = (i + H < CNT) ? v[i + H] : 0.0;
            const double send = up ? lo : hi, keep = up ? hi : lo;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        if (up) { base += H; lim = max(lim - H, 0); }
        else lim = min(lim, H);
        ReduceScatter<MAXV, H, O / 2>::run(v, lane, base, lim);
    }
};
template <int MAXV, int CNT>
struct ReduceScatter<MAXV, CNT, 0> {
    static RP_D void run(double (&)[MAXV], int, int &, int &) {}
};
template <int CNT, int STEPS>
struct ReduceScatterLeft {
    static constexpr int value = ReduceScatterLeft<(CNT + 1) / 2, STEPS - 1>::value;
};
template <int CNT>
struct ReduceScatterLeft<CNT, 0> {
    static constexpr int value = CNT;
};

// One block of THREADS threads per problem.  (One warp per problem was measured 1.8-6x slower on B200: twelve
// problems per SM no longer fit their correspondences in L1.)
// LOSS: RP_LOSS_TRUNCATED for the LO refinement (use_final = 0, loss fixed at compile time), -1 for the final
// refinement with the user's bundle options (loss read at run time).
template <int VARIANT, int NP, int THREADS, int LOSS>
__global__ void __launch_bounds__(THREADS, RP_LM_MIN_BLOCKS * 128 / THREADS) lm_kernel(LMArgs a) {
    constexpr int LM_THREADS = THREADS;
    constexpr int LM_WARPS = THREADS / 32;
    constexpr int NA = NP * (NP + 1) / 2;
    __shared__ Model cur, trial;

    __shared__ double red[LM_WARPS][NA + NP];
    __shared__ double sA[NA], sg[NP];
    __shared__ double cred[LM_WARPS];
    __shared__ int stop_s;
    // masked problems (the final refinement on the inlier set): the inlier indices, compacted once per
    // problem so that every lane of the evaluation passes has work (problems with more than LM_LIST_CAP
    // correspondences test the mask on the fly instead)
    __shared__ unsigned short list_s[LM_LIST_CAP];
    __shared__ int wcount[LM_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_prob = *a.n_prob;
    __shared__ int pj_s;
    for (;;) {
        __syncthreads();
        if (tid == 0) pj_s = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int pj = pj_s;
        if (pj >= n_prob) break;
        const int prob = a.prob_list ? a.prob_list[pj] : pj;
        const int pair = prob / a.prob_per_pair;
        const PairParams pp = a.pairs[pair];
        if (!pp.valid) continue;
        if (a.enable && !a.enable[pair]) continue;
        LMParams P;
        P.weight_sampson = a.weight_sampson;
        P.scale_reproj = a.scale_reproj_override >= 0.0 ? a.scale_reproj_override : pp.scale_reproj;
        if (a.use_final) {
            P.loss_type = LOSS >= 0 ? LOSS : a.loss_type;
            P.loss_scale = a.loss_scale_override > 0.0 ? a.loss_scale_override : pp.final_loss_scale;
        } else {
            P.loss_type = RP_LOSS_TRUNCATED;
            P.loss_scale = pp.lo_loss_scale;
        }
        P.inv_t2 = 1.0 / (P.loss_scale * P.loss_scale);
        const int max_it = a.use_final ? a.max_iterations : 25;
        const int n = pp.n;
        const Pt64 *pts = a.pts64 + pp.off;
        const double *d1 = a.d1 + pp.off, *d2 = a.d2 + pp.off;
        const unsigned char *mask = a.mask ? a.mask + pp.off : nullptr;
        __syncthreads();
        if (tid == 0) { cur = a.models[prob]; stop_s = 0; }
        __syncthreads();

        const bool use_list = mask != nullptr && n <= LM_LIST_CAP;
        int m_work = n;
        if (use_list) {
            // ordered compaction: warp w counts, then writes, the inliers of its contiguous range
            const int lo = (int)(((long long)n * wid) / LM_WARPS), hi = (int)(((long long)n * (wid + 1)) / LM_WARPS);
            int cnt = 0;
            for (int b0 = lo; b0 < hi; b0 += 32) {
                const int k = b0 + lane;
                cnt += __popc(__ballot_sync(0xffffffffu, k < hi && mask[k]));
            }
            if (lane == 0) wcount[wid] = cnt;
            __syncthreads();
            int start = 0;
            m_work = 0;
#pragma unroll
            for (int w = 0; w < LM_WARPS; ++w) { if (w < wid) start += wcount[w]; m_work += wcount[w]; }
            for (int b0 = lo; b0 < hi; b0 += 32) {
                const int k = b0 + lane;
                const bool in = k < hi && mask[k];
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) list_s[start + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)k;
                start += __popc(bal);
            }
            __syncthreads();
        }

        unsigned long long rows = 0;   // accumulated rows of this thread (three 21-bit fields, see LM_FLOPS)
        auto block_cost = [&](const Model &m) -> double {
            const LMFrame F = make_frame(m);
            double c = 0.0;
            LogProd lp;
            lp.init();
            for (int i = tid; i < m_work; i += LM_THREADS) {
                const int k = use_list ? (int)list_s[i] : i;
                if (!use_list && mask && !mask[k]) continue;
                const Pt64 p = pts[k];
                c += point_cost<VARIANT, LOSS>(F, P, p.x1_0, p.x1_1, p.x2_0, p.x2_1, d1[k], d2[k], lp);
            }
            if (lm_loss_may_be_cauchy(LOSS)) c += P.loss_scale * P.loss_scale * lp.total(P.weight_sampson);
            c = warp_sum(c);
            __syncthreads();
            if (lane == 0) cred[wid] = c;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < LM_WARPS; ++w) tot += cred[w];
            return tot;
        };

        // one pass over the correspondences: cost of `m` (block total) and, in the caller's registers, the
        // per-thread partial normal equations at `m`
        auto block_eval = [&](const Model &m, NormalEq<NP> &N) -> double {
            const LMFrame F = make_frame(m);
            double c = 0.0;
            LogProd lp;
            lp.init();
            N.clear();
#pragma unroll LM_UNROLL
            for (int i = tid; i < m_work; i += LM_THREADS) {
                const int k = use_list ? (int)list_s[i] : i;
                if (!use_list && mask && !mask[k]) continue;
                const Pt64 p = pts[k];
                c += point_eval<VARIANT, NP, LOSS>(F, P, p.x1_0, p.x1_1, p.x2_0, p.x2_1, d1[k], d2[k], N, rows, lp);
            }
            if (lm_loss_may_be_cauchy(LOSS)) c += P.loss_scale * P.loss_scale * lp.total(P.weight_sampson);
            c = warp_sum(c);
            __syncthreads();
            if (lane == 0) cred[wid] = c;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < LM_WARPS; ++w) tot += cred[w];
            return tot;
        };
        // per-thread partials -> sA / sg (fixed order: deterministic)
        auto reduce_normal = [&](NormalEq<NP> &N) {
            constexpr int NV = NA + NP, LEFT = ReduceScatterLeft<NV, 5>::value;
            double v[NV];
#pragma unroll
            for (int i = 0; i < NA; ++i) v[i] = N.A[i];
#pragma unroll
            for (int i = 0; i < NP; ++i) v[NA + i] = N.g[i];
            int base = 0, lim = NV;
            ReduceScatter<NV, NV, 16>::run(v, lane, base, lim);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < LEFT; ++j)
                if (j < lim) red[wid][base + j] = v[j];
            __syncthreads();
            if (tid < NA + NP) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < LM_WARPS; ++w) v += red[w][tid];
                if (tid < NA) sA[tid] = v; else sg[tid - NA] = v;
            }
            __syncthreads();
        };

        // lm_impl evaluates the cost of a trial step and, if it is accepted, the Jacobians at the new point.
        // Both walk the same residuals, so they are fused: the trial pass accumulates the normal equations
        // speculatively and they are reduced only when the step is accepted (the last iteration, whose
        // Jacobians nobody would read, is a cost-only pass).
        NormalEq<NP> N;
        int passes = 0;                // passes over the correspondences (uniform)
        double cost;
        if (max_it > 0) { cost = block_eval(cur, N); reduce_normal(N); }
        else cost = block_cost(cur);
        ++passes;
        const double initial_cost = cost;
        double lambda = a.initial_lambda;
        double grad_norm = -1.0, step_norm = -1.0;
        long long invalid_steps = 0;
        bool recompute = true;
        int it = 0;
        for (; it < max_it; ++it) {
            if (recompute) {
                double g2 = 0.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) g2 += sg[i] * sg[i];
                grad_norm = sqrt(g2);
                if (grad_norm < a.gradient_tol) break;  // uniform: every thread reads the same sg
            }
            if (tid == 0) {
                double x[NP];
                llt_solve<NP>(sA, lambda, sg, x);
                double sn = 0.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) { x[i] = -x[i]; sn += x[i] * x[i]; }
                cred[0] = sqrt(sn);
                if (sqrt(sn) < a.step_tol) stop_s = 1;
                else trial = model_step<VARIANT>(cur, x);
            }
            __syncthreads();
            step_norm = cred[0];
            if (stop_s) break;
            const bool last = it == max_it - 1;
            const double cost_new = last ? block_cost(trial) : block_eval(trial, N);
            ++passes;
            if (cost_new < cost) {
                __syncthreads();
                if (tid == 0) cur = trial;
                lambda = fmax(a.min_lambda, lambda / 10);
                cost = cost_new;
                if (!last) reduce_normal(N);
                recompute = true;
            } else {
                ++invalid_steps;
                lambda = fmin(a.max_lambda, lambda * 10);
                recompute = false;
            }
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0) {
            a.models[prob] = cur;
            if (a.stats) {
                rp_bundle_stats s;
                s.iterations = it; s.initial_cost = initial_cost; s.cost = cost; s.lambda = lambda;
                s.invalid_steps = invalid_steps; s.step_norm = step_norm; s.grad_norm = grad_norm;
                a.stats[prob] = s;
            }
            if (a.lm_iters) atomicAdd(a.lm_iters, (unsigned long long)it);
        }
        if (a.lm_flops) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);   // fields stay < 2^21 per block
            if (lane == 0) {
                const unsigned long long fl = (rows & 0x1fffffull) * LM_FLOPS[VARIANT][1] + ((rows >> 21) & 0x1fffffull) * LM_FLOPS[VARIANT][2] +
                                              ((rows >> 42) & 0x1fffffull) * LM_FLOPS[VARIANT][3] +
                                              (wid == 0 ? (unsigned long long)passes * m_work * LM_FLOPS[VARIANT][0] : 0ull);
                atomicAdd(a.lm_flops, fl);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lm_warp_kernel: the same LM, ONE WARP per problem (unmasked problems: the LO refinements).
// The block-per-problem kernel above spends a quarter of its issue slots waiting at block barriers (four warps of
// unequal work meet twice per pass, and 127 threads wait for the serial Cholesky) and on the loads of the next
// correspondence.  Here a warp owns its problem: nothing but warp-level synchronisation, the serial solve of one
// problem overlaps the passes of the eleven other warps of the SM, and every lane prefetches its own next two
// correspondences into its own shared-memory slots with cp.async (no registers held by loads in flight; the data of
// twelve different problems per SM does not fit the L1, which is what made the round-1 warp-per-problem
// attempt 1.8-6x slower).  A lane only ever reads the slots it filled itself, so cp.async.wait_group is the only
// synchronisation the ring needs.
#ifndef RP_LMW_WPB
#define RP_LMW_WPB 4, 4, 4, 4      // warps (= problems in flight) per block of lm_warp_kernel, per variant (build knobs)
#endif
#ifndef RP_LMW_BPS
#define RP_LMW_BPS 3, 2, 2, 2      // blocks per SM
#endif
// Resident warps per SM, i.e. registers per thread = 65536 / (32 x warps): the 7-parameter problem fits 168 registers
// (12 warps); the 8- and 9-parameter ones spill there (45 + 9 FP64 accumulators, 30 doubles of frame) and were measured
// faster with 8 warps x 255 registers and no spills than with 12 warps that spill.
constexpr int LMW_WPB[4] = {RP_LMW_WPB};
constexpr int LMW_BPS[4] = {RP_LMW_BPS};
constexpr int LMW_MAX_WPB = 8;
constexpr int LMW_SLOTS = 3;       // ring slots per lane: prefetch distance 2, the refill targets the slot read one iteration earlier

RP_D void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
RP_D void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
RP_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
RP_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VARIANT, int NP, int LOSS>
__global__ void __launch_bounds__(LMW_WPB[VARIANT] * 32, LMW_BPS[VARIANT]) lm_warp_kernel(LMArgs a) {
    constexpr int NA = NP * (NP + 1) / 2;
    struct __align__(16) Slot { double x[4]; double d1, d2; };   // one correspondence: Pt64 + depths
    struct WarpState {
        Model cur, trial;
        double sA[NA], sg[NP];
        double step_norm;
        int stop;
    };
    __shared__ WarpState ws_all[LMW_WPB[VARIANT]];
    __shared__ Slot ring_all[LMW_WPB[VARIANT]][LMW_SLOTS][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpState &W = ws_all[wid];
    Slot (*ring)[32] = ring_all[wid];
    const int n_prob = *a.n_prob;
    for (;;) {
        int pj = 0;
        if (lane == 0) pj = atomicAdd(a.work_counter, 1);
        pj = __shfl_sync(0xffffffffu, pj, 0);
        if (pj >= n_prob) break;
        const int prob = a.prob_list ? a.prob_list[pj] : pj;
        const int pair = prob / a.prob_per_pair;
        const PairParams pp = a.pairs[pair];
        if (!pp.valid) continue;
        if (a.enable && !a.enable[pair]) continue;
        LMParams P;
        P.weight_sampson = a.weight_sampson;
        P.scale_reproj = a.scale_reproj_override >= 0.0 ? a.scale_reproj_override : pp.scale_reproj;
        if (a.use_final) {
            P.loss_type = LOSS >= 0 ? LOSS : a.loss_type;
            P.loss_scale = a.loss_scale_override > 0.0 ? a.loss_scale_override : pp.final_loss_scale;
        } else {
            P.loss_type = RP_LOSS_TRUNCATED;
            P.loss_scale = pp.lo_loss_scale;
        }
        P.inv_t2 = 1.0 / (P.loss_scale * P.loss_scale);
        const int max_it = a.use_final ? a.max_iterations : 25;
        const int n = pp.n;
        const Pt64 *pts = a.pts64 + pp.off;
        const double *d1 = a.d1 + pp.off, *d2 = a.d2 + pp.off;
        __syncwarp();
        if (lane == 0) { W.cur = a.models[prob]; W.stop = 0; }
        __syncwarp();

        // this lane's correspondence i -> its slot s (always commits a group, so the group count is uniform)
        auto fetch = [&](int i, int s) {
            if (i < n) {
                Slot *d = &ring[s][lane];
                const char *src = reinterpret_cast<const char *>(pts + i);
                cp_async16(&d->x[0], src);
                cp_async16(&d->x[2], src + 16);
                cp_async8(&d->d1, d1 + i);
                cp_async8(&d->d2, d2 + i);
            }
            cp_async_commit();
        };

        unsigned long long rows = 0;
        // one pass over the correspondences; JAC: accumulate the normal equations as well
        auto warp_pass = [&](const Model &m, NormalEq<NP> &N, auto jac_tag) -> double {
            constexpr bool jac = decltype(jac_tag)::value;
            const LMFrame F = make_frame(m);
            double c = 0.0;
            LogProd lp;
            lp.init();
            if (jac) N.clear();
            fetch(lane, 0);
            fetch(lane + 32, 1);
            int s = 0;
            for (int i = lane; i < n; i += 32) {
                cp_async_wait<1>();
                const Slot q = ring[s][lane];
                int s2 = s + 2; if (s2 >= LMW_SLOTS) s2 -= LMW_SLOTS;
                fetch(i + 64, s2);
                if (jac) c += point_eval<VARIANT, NP, LOSS>(F, P, q.x[0], q.x[1], q.x[2], q.x[3], q.d1, q.d2, N, rows, lp);
                else c += point_cost<VARIANT, LOSS>(F, P, q.x[0], q.x[1], q.x[2], q.x[3], q.d1, q.d2, lp);
                if (++s == LMW_SLOTS) s = 0;
            }
            cp_async_wait<0>();
            if (lm_loss_may_be_cauchy(LOSS)) c += P.loss_scale * P.loss_scale * lp.total(P.weight_sampson);
            return warp_sum(c);
        };
        auto reduce_normal = [&](NormalEq<NP> &N) {
            constexpr int NV = NA + NP, LEFT = ReduceScatterLeft<NV, 5>::value;
            double v[NV];
#pragma unroll
            for (int i = 0; i < NA; ++i) v[i] = N.A[i];
#pragma unroll
            for (int i = 0; i < NP; ++i) v[NA + i] = N.g[i];
            int base = 0, lim = NV;
            ReduceScatter<NV, NV, 16>::run(v, lane, base, lim);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < LEFT; ++j)
                if (j < lim) { const int e = base + j; if (e < NA) W.sA[e] = v[j]; else W.sg[e - NA] = v[j]; }
            __syncwarp();
        };

        NormalEq<NP> N;
        int passes = 0;
        double cost;
        if (max_it > 0) { cost = warp_pass(W.cur, N, std::true_type{}); reduce_normal(N); }
        else cost = warp_pass(W.cur, N, std::false_type{});
        ++passes;
        const double initial_cost = cost;
        double lambda = a.initial_lambda;
        double grad_norm = -1.0, step_norm = -1.0;
        long long invalid_steps = 0;
        bool recompute = true;
        int it = 0;
        for (; it < max_it; ++it) {
            if (recompute) {
                double g2 = 0.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) g2 += W.sg[i] * W.sg[i];
                grad_norm = sqrt(g2);
                if (grad_norm < a.gradient_tol) break;
            }
            if (lane == 0) {
                double x[NP];
                llt_solve<NP>(W.sA, lambda, W.sg, x);
                double sn = 0.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) { x[i] = -x[i]; sn += x[i] * x[i]; }
                W.step_norm = sqrt(sn);
                if (sqrt(sn) < a.step_tol) W.stop = 1;
                else W.trial = model_step<VARIANT>(W.cur, x);
            }
            __syncwarp();
            step_norm = W.step_norm;
            if (W.stop) break;
            const bool last = it == max_it - 1;
            const double cost_new = last ? warp_pass(W.trial, N, std::false_type{}) : warp_pass(W.trial, N, std::true_type{});
            ++passes;
            if (cost_new < cost) {
                __syncwarp();
                if (lane == 0) W.cur = W.trial;
                lambda = fmax(a.min_lambda, lambda / 10);
                cost = cost_new;
                if (!last) reduce_normal(N);
                recompute = true;
            } else {
                ++invalid_steps;
                lambda = fmin(a.max_lambda, lambda * 10);
                recompute = false;
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) {
            a.models[prob] = W.cur;
            if (a.stats) {
                rp_bundle_stats s;
                s.iterations = it; s.initial_cost = initial_cost; s.cost = cost; s.lambda = lambda;
                s.invalid_steps = invalid_steps; s.step_norm = step_norm; s.grad_norm = grad_norm;
                a.stats[prob] = s;
            }
            if (a.lm_iters) atomicAdd(a.lm_iters, (unsigned long long)it);
        }
        if (a.lm_flops) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
            if (lane == 0) {
                const unsigned long long fl = (rows & 0x1fffffull) * LM_FLOPS[VARIANT][1] + ((rows >> 21) & 0x1fffffull) * LM_FLOPS[VARIANT][2] +
                                              ((rows >> 42) & 0x1fffffull) * LM_FLOPS[VARIANT][3] +
                                              (unsigned long long)passes * n * LM_FLOPS[VARIANT][0];
                atomicAdd(a.lm_flops, fl);
            }
        }
    }
}

// defined in repose_lm.cu; returns the cudaError_t of the launch
int launch_lm_kernel(int sms, int variant, const LMArgs &a, cudaStream_t st);

}  // namespace rp
