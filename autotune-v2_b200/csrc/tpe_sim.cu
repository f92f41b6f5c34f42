// (c) TPE: Parzen-estimator suggestion fused with the Gamma-process simulation and successive halving - the hybrids
// (Hyperband whose first round of every bracket is a TPE run, with optional history transfer between brackets), plus
// stand-alone TPE and random search as degenerate single-bracket plans.
//
// Reference call sites: optimisers/sequential/tpe_optimiser.py:40-89 (hyperopt.fmin with tpe.suggest,
// n_startup_jobs = 10 + injected), hybrid_hyperband_tpe_optimiser.py:41-95,
// hybrid_hyperband_tpe_with_transfer.py:50-209 (transfer predicates, latest-resource bookkeeping).  The TPE arithmetic
// itself lives in third-party hyperopt 0.2.4 (absent from the reference tree and from this image): what is implemented
// here is the published algorithm as restated in oracle/tpe.py - adaptive Parzen estimators over the "good"
// (n_below = min(ceil(gamma sqrt N), 25) lowest losses) and "bad" observations of each hp.uniform hyperparameter
// independently, 24 candidates drawn from the truncated good mixture, the candidate maximising log l(x) - log g(x).
//
// Mapping: one warp = one repetition.  The TPE phase of a bracket is sequential in its arms (suggestion k needs the
// losses of arms 0..k-1 at the bracket's first resource level), so the warp's 32 lanes are spent *inside* a suggestion:
// lanes = candidates for sampling and scoring (pairwise candidate x component kernel evaluations, components
// broadcast from shared memory), lanes = observations for the order statistics (good/bad split, sorted insertion,
// compaction, bandwidths).  FP32 CUDA cores + MUFU (ex2/lg2/erf); not a dense contraction, so no tensor cores.
// Trajectories use exactly the streams and arithmetic of the Hyperband kernel (sim_core.cuh): with TPE disabled the
// results are bit-identical to at2_hb_sim_f32.
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "at2_common.cuh"
#include "halving.cuh"
#include "sim_core.cuh"

namespace {

struct TpeDev {
    int enabled, n_startup, n_cand, lf, transfer, threshold;
    float gamma, prior_weight;
};

struct TpeParams {
    at2_plan plan;
    Fam<float> fam[AT2_MAX_FAMILIES];
    ArmSpace sp;
    TpeDev tpe;
    int descending;
    const float* tables;
    long long n_reps, rep_offset;
    int npad, cap, warp_bytes, scratch_floats;
    float* ofe;
    int* ofe_arm;
    float* best_xy;
    float* round_best;
    int* round_best_arm;
    int* survivors;
    float* ckpt_loss;
    float* arm_xy;      // optional [n_reps][n_arms][2]
    float* obs_loss;    // optional [n_reps][n_arms]: the loss TPE observed (only written for observed arms)
    int* n_injected;    // optional [n_reps][n_brackets]: observations transferred into the bracket's TPE run
};

// warp-private working set
struct TpeWarp {
    float *ck, *ax, *ay, *ox, *oy, *ol;
    float4* comp;                    // mixture components {coef, sqrt(log2(e)/2)/sigma, mu, weight}
    unsigned short *srt0, *srt1, *age;
    unsigned short* gidx;            // observation indices of the "good" members, best loss first (tpe_split); <= 25
    unsigned char *latest, *below;
    int n_obs;
};

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// append observation (x, y, l) and keep the per-dimension sorted index lists sorted (stable: after equal values)
__device__ __forceinline__ void tpe_insert(TpeWarp& w, float x, float y, float l, int lane) {
    const int n = w.n_obs;
    if (lane == 0) w.ox[n] = x, w.oy[n] = y, w.ol[n] = l;
    __syncwarp();
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const float* vals = d ? w.oy : w.ox;
        unsigned short* srt = d ? w.srt1 : w.srt0;
        const float v = d ? y : x;
        int cnt = 0;
        for (int j = lane; j < n; j += 32) cnt += vals[srt[j]] <= v ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        for (int base = ((n - 1) >> 5) << 5; base >= 0 && base + 31 >= cnt; base -= 32) {
            const int j = base + lane;
            const bool mv = j >= cnt && j < n;
            const unsigned short e = mv ? srt[j] : 0;
            __syncwarp();
            if (mv) srt[j + 1] = e;
            __syncwarp();
        }
        if (lane == 0) srt[cnt] = (unsigned short)n;
        __syncwarp();
    }
    w.n_obs = n + 1;
}

// flag the n_below observations with the lowest losses (ties: older first) and rank every observation by age inside
// its own set (tpe.ap_filter_trials keeps both sets in trial order; linear forgetting weighs by that order)
__device__ __forceinline__ void tpe_split(TpeWarp& w, int n_below, int lane) {
    const int n = w.n_obs;
    for (int j = lane; j < n; j += 32) w.below[j] = 0;
    __syncwarp();
    for (int s = 0; s < n_below; ++s) {
        float best = FLT_MAX;
        int bi = INT_MAX;
        for (int j = lane; j < n; j += 32) {
            const float l = w.ol[j];
            if (!w.below[j] && (l < best || (l == best && j < bi))) best = l, bi = j;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        if (lane == 0 && bi != INT_MAX) w.below[bi] = 1, w.gidx[s] = (unsigned short)bi;
        __syncwarp();
    }
    int cb = 0, ca = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        const bool valid = j < n;
        const bool fb = valid && w.below[j];
        const unsigned mb = __ballot_sync(0xffffffffu, fb), ma = __ballot_sync(0xffffffffu, valid && !fb);
        if (valid) w.age[j] = (unsigned short)(fb ? cb + __popc(mb & lanemask_lt(lane)) : ca + __popc(ma & lanemask_lt(lane)));
        cb += __popc(mb);
        ca += __popc(ma);
    }
    __syncwarp();
}

__device__ __forceinline__ float lf_weight(int age, int n, int lf) {
    // np.concatenate([np.linspace(1/n, 1, n - lf), np.ones(lf)])[age]   (tpe.linear_forgetting_weights)
    const int num = n - lf;
    if (age >= num) return 1.0f;
    const float lo = 1.0f / (float)n;
    return num == 1 ? lo : lo + (float)age * ((1.0f - lo) / (float)(num - 1));
}

__device__ __forceinline__ float normal_cdf(float x, float mu, float sigma) {
    return 0.5f * (1.0f + erff((x - mu) / fmaxf(1.41421356237f * sigma, 1e-12f)));
}

// Adaptive Parzen estimator (tpe.adaptive_parzen_normal) over the observations of dimension d whose `below` flag equals
// want_below.  Leaves comp[0..m] = {coef, sqrt(log2(e)/2)/sigma, mu, weight} in sorted order with the prior inserted
// and returns m + 1; *sum_w receives the sum of the weights.
//   FULL = true  (stand-alone scoring entry): weights normalised, coef = weight / (sigma * p_accept) with the truncation
//                mass p_accept of tpe.GMM1_lpdf (the common 1/sqrt(2 pi) cancels in l/g).
//   FULL = false (suggestions inside the simulation kernel): weights left un-normalised and coef = weight / sigma.
//                log(sum_w) and log(p_accept) shift the scores of all candidates of a model alike, so the arg-max over the
//                candidates (all a suggestion keeps, tpe.broadcast_best) does not depend on them - and the two erf per
//                component they cost are most of the model build; the sampler scales its uniform by sum_w instead.
//   n_list > 0: the members are the observations w.gidx[0 .. n_list) (the good set as tpe_split lists it, at most 25): one
//                lane per member, sorted by counting - instead of scanning all n sorted observations for the few members.
template <bool FULL>
__device__ __forceinline__ int tpe_build_model(TpeWarp& w, int d, bool want_below, float lo, float hi, const TpeDev& cfg, int lane,
                                               float* sum_w_out, int n_list = 0) {
    const int n = w.n_obs;
    const float* vals = d ? w.oy : w.ox;
    const unsigned short* srt = d ? w.srt1 : w.srt0;
    const float prior_mu = 0.5f * (hi + lo), prior_sigma = hi - lo;
    // pass 1: members in sorted order straight into comp[].z (centre) / comp[].w (age rank inside the set); the prior goes
    // to position pp = np.searchsorted(sorted members, prior_mu) = #members < prior_mu, members >= prior_mu shift by one
    int m = 0, ge = 0;
    if (n_list > 0) {
        // rank among the listed members by (value, observation index): the order of the sorted lists, which keep equal
        // values in insertion order (tpe_insert); ages are not needed (linear forgetting starts beyond 25 members)
        const int id = lane < n_list ? (int)w.gidx[lane] : INT_MAX;
        const float v = lane < n_list ? vals[id] : 0.0f;
        int rank = 0;
        for (int q = 0; q < n_list; ++q) {
            const float vq = __shfl_sync(0xffffffffu, v, q);
            const int iq = __shfl_sync(0xffffffffu, id, q);
            rank += (vq < v || (vq == v && iq < id)) ? 1 : 0;
        }
        const bool g = lane < n_list && !(v < prior_mu);
        if (lane < n_list) {
            float4* c = w.comp + (rank + (g ? 1 : 0));
            c->z = v, c->w = 0.0f;
        }
        m = n_list, ge = __popc(__ballot_sync(0xffffffffu, g));
    } else {
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            const int idx = j < n ? srt[j] : 0;
            const float v = vals[idx];
            const bool f = j < n && (w.below[idx] != 0) == want_below;
            const bool g = f && !(v < prior_mu);
            const unsigned mask = __ballot_sync(0xffffffffu, f), gmask = __ballot_sync(0xffffffffu, g);
            if (f) {
                float4* c = w.comp + (m + __popc(mask & lanemask_lt(lane)) + (g ? 1 : 0));
                c->z = v, c->w = (float)w.age[idx];
            }
            m += __popc(mask), ge += __popc(gmask);
        }
    }
    const int pp = m - ge, nc = m + 1;
    if (lane == 0) w.comp[pp].z = prior_mu;
    __syncwarp();
    const float minsigma = prior_sigma / fminf(100.0f, 1.0f + (float)nc), maxsigma = prior_sigma;
    const bool forget = cfg.lf > 0 && cfg.lf < m;
    float sum_w = 0.0f, sum_acc = 0.0f;
    // pass 2: bandwidths from the neighbouring centres, weights.  (Reads of .z of the neighbours race with nobody: this
    // pass only writes .x/.y/.w of its own component.)
    for (int p = lane; p < nc; p += 32) {
        const float mu = w.comp[p].z;
        float sigma, wt;
        if (p == pp) {
            sigma = prior_sigma, wt = cfg.prior_weight;
        } else {
            if (m == 1)
                sigma = 0.5f * prior_sigma;
            else if (p == 0)
                sigma = w.comp[1].z - mu;
            else if (p == nc - 1)
                sigma = mu - w.comp[p - 1].z;
            else
                sigma = fmaxf(mu - w.comp[p - 1].z, w.comp[p + 1].z - mu);
            sigma = fminf(fmaxf(sigma, minsigma), maxsigma);
            wt = forget ? lf_weight((int)w.comp[p].w, m, cfg.lf) : 1.0f;
        }
        sum_w += wt;
        if constexpr (FULL) {
            w.comp[p].x = sigma, w.comp[p].w = wt;
            sum_acc += wt * (normal_cdf(hi, mu, sigma) - normal_cdf(lo, mu, sigma));
        } else {
            const float inv_sigma = fast_rcp(sigma);
            float4* c = w.comp + p;
            c->x = wt * inv_sigma, c->y = 0.84932180028801904f * inv_sigma, c->w = wt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_w += __shfl_xor_sync(0xffffffffu, sum_w, o);
        if constexpr (FULL) sum_acc += __shfl_xor_sync(0xffffffffu, sum_acc, o);
    }
    __syncwarp();
    if constexpr (FULL) {
        const float inv_w = 1.0f / sum_w;
        const float inv_acc = sum_w / sum_acc;                   // 1 / p_accept
        for (int p = lane; p < nc; p += 32) {
            const float4 c = w.comp[p];
            const float wn = c.w * inv_w;
            w.comp[p] = make_float4(wn * inv_acc / c.x, 0.84932180028801904f / c.x, c.z, wn);
        }
        __syncwarp();
        sum_w = 1.0f;
    }
    *sum_w_out = sum_w;
    return nc;
}

// sum_p coef_p * exp(-((x - mu_p)/sigma_p)^2 / 2) of the model currently in comp[]
__device__ __forceinline__ float tpe_mixture(const TpeWarp& w, int nc, float x) {
    float s = 0.0f;
    for (int p = 0; p < nc; ++p) {
        const float4 c = w.comp[p];
        const float z = (x - c.z) * c.y;
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z));
        s = fmaf(c.x, e, s);
    }
    return s;
}

// Candidates from the mixture in comp[0 .. nc) (layout of tpe_build_model<false>) truncated to [lo, hi): lane j < n_cand
// returns candidate j, drawn from the two random words (wa, wb) it is given; the other lanes return the centre.
// tpe.GMM1 draws (component, normal) pairs until the value falls inside the bounds.  The accepted pair has component i with
// probability proportional to w_i a_i (a_i: mass of component i inside the bounds) and then a normal truncated to the
// bounds - which is sampled here directly: one lane per component computes a_i, every candidate lane picks a component
// and inverts the truncated CDF.  All 32 lanes must call it; nc <= 32; uses comp[nc .. 2 nc) as scratch.
__device__ __forceinline__ float tpe_sample_truncated(TpeWarp& w, int nc, float lo, float hi, int n_cand, uint32_t wa, uint32_t wb,
                                                      int lane) {
    float sigma = 1.0f, f_lo = 0.0f, mass = 0.0f, eff = 0.0f;
    if (lane < nc) {
        const float4 c = w.comp[lane];
        sigma = 0.84932180028801904f * fast_rcp(c.y);
        const float zs = 0.70710678118654752f * fast_rcp(sigma);              // Phi(z) = (1 + erf(z / sqrt 2)) / 2
        f_lo = fmaf(0.5f, erff((lo - c.z) * zs), 0.5f);
        mass = fmaf(0.5f, erff((hi - c.z) * zs), 0.5f) - f_lo;
        eff = c.w * mass;
    }
    float cum = eff;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += t;
    }
    const float total = __shfl_sync(0xffffffffu, cum, nc - 1);
    if (lane < nc) w.comp[nc + lane] = make_float4(cum, f_lo, mass, sigma);   // beyond the model: free space
    __syncwarp();
    float cand = 0.5f * (lo + hi);
    if (lane < n_cand) {
        const float u = (float)(wa >> 8) * 5.9604644775390625e-8f * total;                      // [0, total)
        const float t = fmaf((float)(wb >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);  // (0, 1)
        int pick = 0;
        while (pick < nc - 1 && u >= w.comp[nc + pick].x) ++pick;
        const float4 g = w.comp[nc + pick];
        const float v = fmaf(g.w, normcdfinvf(fmaf(t, g.z, g.y)), w.comp[pick].z);
        cand = fminf(fmaxf(v, lo), nextafterf(hi, lo));                                          // [lo, hi)
    }
    return cand;
}

// One hyperparameter of one suggestion: candidates from the good model, argmax of log l - log g
// (tpe.build_posterior, ap_uniform_sampler, GMM1, GMM1_lpdf, broadcast_best).
// Candidates come from tpe_sample_truncated instead of a rejection loop: the wide prior component alone rejects 62 % of
// its draws, so with 24 lanes the loop ran three to four Philox blocks deep on almost every suggestion.
// r0: the lane's Philox block of this suggestion; words (2 d, 2 d + 1) serve dimension d.
__device__ __forceinline__ float tpe_suggest_dim(TpeWarp& w, int d, float lo, float hi, const TpeDev& cfg, int n_good,
                                                 const At2U4& r0, int lane) {
    float cand = 0.5f * (lo + hi);
    float lik[2] = {1.0f, 1.0f};
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {                   // 0: good model (sample from it), 1: bad model
        float sum_w;
        const int nc = tpe_build_model<false>(w, d, which == 0, lo, hi, cfg, lane, &sum_w, which == 0 ? n_good : 0);
        if (which == 0) cand = tpe_sample_truncated(w, nc, lo, hi, cfg.n_cand, d ? r0.z : r0.x, d ? r0.w : r0.y, lane);
        __syncwarp();
        lik[which] = tpe_mixture(w, nc, cand);
        __syncwarp();
    }
    const float l_good = lik[0], l_bad = lik[1];
    float score = lane < cfg.n_cand ? fast_lg2(l_good) - fast_lg2(l_bad) : -FLT_MAX;
    if (score != score) score = -FLT_MAX;
    int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                          // np.argmax: first maximum
        const float os = __shfl_xor_sync(0xffffffffu, score, o);
        const int ow = __shfl_xor_sync(0xffffffffu, who, o);
        if (os > score || (os == score && ow < who)) score = os, who = ow;
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, cand, who);
}

__device__ __forceinline__ bool transferable(int mode, int r_arm, double r_bracket, int R, int threshold) {
    switch (mode) {   // hybrid_hyperband_tpe_with_transfer.py:161-209
        case AT2_TRANSFER_LONGEST: return r_arm >= R;
        case AT2_TRANSFER_ALL: return (double)r_arm >= r_bracket;
        case AT2_TRANSFER_SAME: return (double)r_arm == r_bracket;
        case AT2_TRANSFER_THRESHOLD: return r_arm >= threshold && (double)r_arm >= r_bracket;
        default: return false;
    }
}

template <bool HAS_SMOOTH, int CK>
#ifndef AT2_TPE_MAX_WARPS
#define AT2_TPE_MAX_WARPS 24
#endif
// one CTA per SM, up to 24 warps (80 registers per thread): the kernel is a long chain of small dependent warp-wide
// steps, so resident warps are what hides its latency
__global__ void __launch_bounds__(AT2_TPE_MAX_WARPS * 32, 1) tpe_sim_kernel(const __grid_constant__ TpeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = p.plan.R, A = p.plan.n_arms, C = p.plan.n_ckpts, F = p.sp.F, cap = p.cap;
    const At2TableLayout L{R, p.sp.n_slots, HAS_SMOOTH ? 1 : 0, at2_sg_stride(C)};
    const int tab_elems = (L.total_philox() + 3) & ~3;
    float* s_tab = reinterpret_cast<float*>(smem_raw);
    int* s_ckpos = reinterpret_cast<int*>(s_tab + tab_elems);       // [AT2_MAX_CKPTS + 4]
    int* s_sglast = s_ckpos + AT2_MAX_CKPTS + 4;                     // [AT2_MAX_CKPTS + 4] last position a smooth read-out needs
    unsigned char* warp_base = reinterpret_cast<unsigned char*>(s_sglast + AT2_MAX_CKPTS + 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < L.total_philox(); i += blockDim.x) s_tab[i] = p.tables[i];
    if (tid < AT2_MAX_CKPTS + 4) s_ckpos[tid] = tid < C ? p.plan.ckpt_time[tid] - 1 : INT_MAX;
    __syncthreads();
    const float* s_rec = s_tab + L.off_rec();
    const float* s_sg = s_tab + L.off_sg();
    if (tid < AT2_MAX_CKPTS + 4) {
        int last = tid < C ? p.plan.ckpt_time[tid] - 1 : 0;
        if (HAS_SMOOTH && tid < C)
            for (int t = 0; t < R; ++t)
                if (s_sg[t * L.sg_stride + tid] != 0.0f) last = max(last, t);
        s_sglast[tid] = last;
    }
    __syncthreads();
    const SimTables tb{(uint32_t)__cvta_generic_to_shared(s_rec), (uint32_t)__cvta_generic_to_shared(s_sg), s_ckpos, R,
                       L.sg_stride};

    // carve the warp's working set (the sort scratch of the halving rounds aliases the mixture components: a bracket's
    // TPE phase is over before its halving starts)
    unsigned char* mine = warp_base + (size_t)warp * p.warp_bytes;
    TpeWarp w;
    float* fp = reinterpret_cast<float*>(mine);
    w.ck = fp, fp += (size_t)C * cap;
    w.ax = fp, fp += cap;
    w.ay = fp, fp += cap;
    w.ox = fp, fp += cap;
    w.oy = fp, fp += cap;
    w.ol = fp, fp += cap;
    w.comp = reinterpret_cast<float4*>(fp);
    Key32* keys = reinterpret_cast<Key32*>(fp);
    unsigned short* ids_a = reinterpret_cast<unsigned short*>(keys + p.npad);
    unsigned short* ids_b = ids_a + p.npad;
    fp += p.scratch_floats;
    unsigned short* up = reinterpret_cast<unsigned short*>(fp);
    w.srt0 = up, up += cap;
    w.srt1 = up, up += cap;
    w.age = up, up += cap;
    w.gidx = up, up += 32;
    unsigned char* bp = reinterpret_cast<unsigned char*>(up);
    w.latest = bp, bp += cap;
    w.below = bp, bp += cap;
    w.n_obs = 0;

    const bool desc = p.descending != 0;
    const float sign = desc ? -1.0f : 1.0f;
    const float x_lo = p.sp.x_min, x_hi = p.sp.x_min + p.sp.x_rng, y_lo = p.sp.y_min, y_hi = p.sp.y_min + p.sp.y_rng;
    const HalvingOut<float> ho{p.ofe, p.ofe_arm, p.round_best, p.round_best_arm, p.survivors};

    for (long long rep = (long long)blockIdx.x * nwarps + warp; rep < p.n_reps; rep += (long long)gridDim.x * nwarps) {
        HalvingState<float> st{0.0f, -1, 0};
        const unsigned long long gid0 = (unsigned long long)(p.rep_offset + rep) * (unsigned long long)A;
        for (int b = 0; b < p.plan.n_brackets; ++b) {
            const int n_b = p.plan.bracket_n[b], first = p.plan.bracket_first_arm[b];
            const int rd0 = p.plan.bracket_first_round[b];
            const int c0 = p.plan.round_ckpt[rd0];
            const double r0 = p.plan.bracket_r0[b];   // the un-truncated float the reference compares against
            // (0) one arm per lane: the start-up hyperparameters of every arm of the bracket and - when the bracket runs
            //     TPE - the offset of its loss at the bracket's first resource level over f(x, y), which depends only on
            //     the arm's own streams (sim_core.cuh philox_affine_offset), parked in row 0 of ck until (3) fills it;
            // (1) arms of earlier brackets that may be injected as TPE history
            //     (hybrid_hyperband_tpe_with_transfer.py:116-158);
            // (2) the bracket's own arms one after the other: suggest -> loss = f(x, y) + offset -> observe;
            // (3) the full trajectories, one arm per lane, exactly as the Hyperband kernel runs them.
            // Each big routine (insert, suggest, trajectory) is instantiated exactly once: with them duplicated the kernel
            // was instruction-fetch bound (ncu: 49 % of stall samples in no_instructions).
            w.n_obs = 0;
            const bool use_tpe = p.tpe.enabled && n_b > p.tpe.n_startup;
            const bool inject = use_tpe && b > 0 && p.tpe.transfer != AT2_TRANSFER_NONE;
            const int n_prev = inject ? first : 0;
            int n_startup = p.tpe.n_startup;                          // + injected, tpe_optimiser.py:52
            int n_inj = 0;
#pragma unroll 1
            for (int base = 0; base < n_b; base += 32) {
                const bool act = base + lane < n_b;
                const int arm = first + (act ? base + lane : 0);
                const ArmDraw d = draw_arm(p.sp, gid0 + arm, arm);
                float off = 0.0f;
                if (use_tpe) {
                    const Fam<float> fam = p.fam[d.fam];
                    const bool sm = HAS_SMOOTH && fam.smooth != 0;
                    off = philox_affine_offset<HAS_SMOOTH>(tb, act, sm, fam.up, fam.pw_off, gid0 + arm, p.sp.rk,
                                                           (sm ? s_sglast[c0] : s_ckpos[c0]) + 1, c0,
                                                           p.sp.noise * d.z - fam.s0, fam.s1);
                }
                if (act) w.ax[arm] = d.x, w.ay[arm] = d.y, w.ck[arm] = off;
            }
            __syncwarp();
#pragma unroll 1
            for (int step = -n_prev; step < (use_tpe ? n_b : 0); ++step) {
                bool do_insert = false;
                float ix = 0.0f, iy = 0.0f, il = 0.0f;
                if (step < 0) {                                      // (1) candidate for injection
                    const int a = first + step;
                    const int lc = w.latest[a] & 0x7f;
                    // the resource count the reference recorded for the arm: int(r_i), which is 0 (not R) where the read
                    // wrapped to the last point (R = 729, eta = 3)
                    const int r_arm = (w.latest[a] & 0x80) ? 0 : p.plan.ckpt_time[lc];
                    if (transferable(p.tpe.transfer, r_arm, r0, R, p.tpe.threshold))
                        do_insert = true, ++n_inj, ix = w.ax[a], iy = w.ay[a], il = sign * w.ck[lc * cap + a];
                } else {                                             // (2) next arm of the bracket
                    if (step == 0) n_startup += w.n_obs;
                    const int arm = first + step;
                    float xy[2] = {w.ax[arm], w.ay[arm]};
                    if (w.n_obs >= n_startup) {
                        const int n_good = (int)fminf(ceilf(p.tpe.gamma * sqrtf((float)w.n_obs)), 25.0f);
                        tpe_split(w, n_good, lane);
                        const unsigned long long gid = gid0 + arm;
                        const At2U4 r0 = at2_philox10_rk(At2U4{(uint32_t)lane, (uint32_t)gid, (uint32_t)(gid >> 32), AT2_STREAM_TPE},
                                                         p.sp.rk);
#pragma unroll 1
                        for (int dd = 0; dd < 2; ++dd)
                            xy[dd] = tpe_suggest_dim(w, dd, dd ? y_lo : x_lo, dd ? y_hi : x_hi, p.tpe, n_good, r0, lane);
                        if (lane == 0) w.ax[arm] = xy[0], w.ay[arm] = xy[1];
                    }
                    if (step + 1 < n_b) {                            // the loss TPE sees for this arm
                        const float loss = base_function(p.sp.func, xy[0], xy[1]) + w.ck[arm];
                        do_insert = true, ix = xy[0], iy = xy[1], il = sign * loss;
                        if (lane == 0 && p.obs_loss != nullptr) p.obs_loss[rep * A + arm] = loss;
                    }
                }
                __syncwarp();
                if (do_insert) tpe_insert(w, ix, iy, il, lane);
            }
            __syncwarp();
            if (lane == 0 && p.n_injected != nullptr) p.n_injected[rep * p.plan.n_brackets + b] = n_inj;
#pragma unroll 1
            for (int base = 0; base < n_b; base += 32) {              // (3) full trajectories
                const bool act = base + lane < n_b;
                const int arm = first + (act ? base + lane : 0);
                // family and initial noise come back from the arm's own stream, the base function from the (possibly
                // suggested) hyperparameters
                const ArmDraw d = draw_arm(p.sp, gid0 + arm, arm);
                const Fam<float> fam = p.fam[d.fam];
                const float f = base_function(p.sp.func, w.ax[arm], w.ay[arm]);
                __syncwarp();                                         // row 0 of ck is overwritten below
                philox_trajectory<float, HAS_SMOOTH, CK>(tb, act, HAS_SMOOTH && fam.smooth != 0, (f + p.sp.noise * d.z) - fam.s0,
                                                         f - fam.s1, fam.up, fam.pw_off, gid0 + arm, p.sp.rk, R,
                                                         w.ck + arm, cap, C);
            }
            __syncwarp();
            warp_halving_bracket<float>(p.plan, b, w.ck, cap, keys, ids_a, ids_b, desc, lane, rep, ho, st, w.latest);
        }
        halving_finish<float>(ho, st, lane, rep);
        if (p.ckpt_loss != nullptr)
            for (int i = lane; i < A * C; i += 32) p.ckpt_loss[rep * A * C + i] = w.ck[(i % C) * cap + i / C];
        if (p.arm_xy != nullptr)
            for (int a = lane; a < A; a += 32) p.arm_xy[(rep * A + a) * 2] = w.ax[a], p.arm_xy[(rep * A + a) * 2 + 1] = w.ay[a];
        if (lane == 0 && p.best_xy != nullptr) p.best_xy[2 * rep] = w.ax[st.best_arm], p.best_xy[2 * rep + 1] = w.ay[st.best_arm];
        __syncwarp();
    }
}

// stand-alone scoring entry (tests): build both models from given observations, score given candidates
__global__ void tpe_score_kernel(const float* __restrict__ obs, const float* __restrict__ loss, int n_obs, float lo, float hi,
                                 TpeDev cfg, const float* __restrict__ cand, int n_cand, float* __restrict__ score_out,
                                 float* __restrict__ model_out, float* __restrict__ lean_out, float* __restrict__ sample_out,
                                 int cap, int n_prob) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t warp_bytes = (size_t)cap * (3 * 4 + 3 * 2 + 1) + 16 * (cap + 4) + 128;
    unsigned char* mine = smem_raw + (size_t)warp * ((warp_bytes + 15) & ~(size_t)15);
    TpeWarp w = {};
    float* fp = reinterpret_cast<float*>(mine);
    w.comp = reinterpret_cast<float4*>(fp), fp += 4 * (cap + 4);
    w.ox = fp, fp += cap;
    w.oy = fp, fp += cap;
    w.ol = fp, fp += cap;
    unsigned short* up = reinterpret_cast<unsigned short*>(fp);
    w.srt0 = up, up += cap;
    w.srt1 = up, up += cap;
    w.age = up, up += cap;
    w.gidx = up, up += 32;
    w.below = reinterpret_cast<unsigned char*>(up);
    for (int prob = blockIdx.x * nwarps + warp; prob < n_prob; prob += gridDim.x * nwarps) {
        w.n_obs = 0;
        for (int j = 0; j < n_obs; ++j)
            tpe_insert(w, obs[(size_t)prob * n_obs + j], 0.0f, loss[(size_t)prob * n_obs + j], lane);
        const int n_good = (int)fminf(ceilf(cfg.gamma * sqrtf((float)n_obs)), 25.0f);
        tpe_split(w, n_good, lane);
        float unused;
        const int nb = tpe_build_model<true>(w, 0, true, lo, hi, cfg, lane, &unused);
        if (model_out != nullptr)
            for (int q = lane; q < nb; q += 32) {
                const float4 c = w.comp[q];
                float* m = model_out + ((size_t)prob * 2 * (cap + 1) + q) * 3;
                m[0] = c.w, m[1] = c.z, m[2] = 0.84932180028801904f / c.y;
            }
        __syncwarp();
        float lg[8];
        for (int q = 0; q < 8; ++q) {
            const int j = q * 32 + lane;
            lg[q] = j < n_cand ? tpe_mixture(w, nb, cand[(size_t)prob * n_cand + j]) : 1.0f;
        }
        __syncwarp();
        const int na = tpe_build_model<true>(w, 0, false, lo, hi, cfg, lane, &unused);
        if (model_out != nullptr)
            for (int q = lane; q < na; q += 32) {
                const float4 c = w.comp[q];
                float* m = model_out + ((size_t)prob * 2 * (cap + 1) + (cap + 1) + q) * 3;
                m[0] = c.w, m[1] = c.z, m[2] = 0.84932180028801904f / c.y;
            }
        __syncwarp();
        for (int q = 0; q < 8; ++q) {
            const int j = q * 32 + lane;
            if (j < n_cand)
                score_out[(size_t)prob * n_cand + j] =
                    0.69314718055994531f * (fast_lg2(lg[q]) - fast_lg2(tpe_mixture(w, na, cand[(size_t)prob * n_cand + j])));
        }
        __syncwarp();
        if (lean_out != nullptr) {      // the models as a suggestion inside the simulation kernel builds them
            const int nb2 = tpe_build_model<false>(w, 0, true, lo, hi, cfg, lane, &unused, n_good);   // the list path
            if (sample_out != nullptr) {   // 32 draws from the good model, the way a suggestion draws its candidates
                const At2U4 r = at2_philox10(At2U4{(uint32_t)lane, (uint32_t)prob, 0u, AT2_STREAM_TPE}, 0xC0FFEEu, 0x5EEDu);
                sample_out[(size_t)prob * 32 + lane] = tpe_sample_truncated(w, nb2, lo, hi, 32, r.x, r.y, lane);
                __syncwarp();
            }
            for (int q = 0; q < 8; ++q) {
                const int j = q * 32 + lane;
                lg[q] = j < n_cand ? tpe_mixture(w, nb2, cand[(size_t)prob * n_cand + j]) : 1.0f;
            }
            __syncwarp();
            const int na2 = tpe_build_model<false>(w, 0, false, lo, hi, cfg, lane, &unused);
            for (int q = 0; q < 8; ++q) {
                const int j = q * 32 + lane;
                if (j < n_cand)
                    lean_out[(size_t)prob * n_cand + j] =
                        0.69314718055994531f * (fast_lg2(lg[q]) - fast_lg2(tpe_mixture(w, na2, cand[(size_t)prob * n_cand + j])));
            }
            __syncwarp();
        }
    }
}

int fill_tpe(const at2_tpe_config* t, TpeDev* d) {
    AT2_REQUIRE(t != nullptr, "NULL tpe config");
    AT2_REQUIRE(t->n_candidates >= 1 && t->n_candidates <= 32, "n_candidates %d outside [1,32]", t->n_candidates);
    AT2_REQUIRE(t->n_startup >= 0 && t->lf >= 0 && t->gamma > 0 && t->prior_weight > 0, "bad TPE parameters");
    AT2_REQUIRE(t->transfer >= AT2_TRANSFER_NONE && t->transfer <= AT2_TRANSFER_THRESHOLD, "unknown transfer mode %d", t->transfer);
    d->enabled = t->enabled, d->n_startup = t->n_startup, d->n_cand = t->n_candidates, d->lf = t->lf;
    d->transfer = t->transfer, d->threshold = t->threshold, d->gamma = (float)t->gamma, d->prior_weight = (float)t->prior_weight;
    return AT2_OK;
}

template <bool HS, int CK>
int launch_tpe(const TpeParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = tpe_sim_kernel<HS, CK>;
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // persistent grid: never more CTAs than are resident at once (registers can bind before shared memory does)
    int by_regs = 0, dev = 0, sms = 0;
    if (at2_resident_by_regs(kern, block, &by_regs) != AT2_OK) return AT2_ECUDA;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if ((long long)by_regs * sms < grid) grid = by_regs * sms;
    kern<<<grid, block, smem, st>>>(p);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}

}  // namespace

extern "C" int at2_single_round_plan(int32_t R, int32_t n_evals, int32_t n_resources, at2_plan* plan) {
    AT2_REQUIRE(plan != nullptr, "at2_single_round_plan: plan is NULL");
    AT2_REQUIRE(R >= 2 && n_evals >= 1 && n_evals < 65536, "at2_single_round_plan: need R >= 2 and 1 <= n_evals < 65536");
    memset(plan, 0, sizeof(*plan));
    // time = int(n_resources); python's fs[time-1] wraps for time <= 0 and raises beyond R
    int t = n_resources >= 1 ? n_resources : R + n_resources;
    AT2_REQUIRE(t >= 1 && t <= R, "list index out of range");
    plan->R = R, plan->eta = 1, plan->n_brackets = 1, plan->n_rounds = 1, plan->n_arms = n_evals, plan->n_ckpts = 1;
    plan->n_surv = 0;
    plan->bracket_n[0] = n_evals, plan->bracket_first_arm[0] = 0, plan->bracket_first_round[0] = 0;
    plan->bracket_first_round[1] = 1;
    plan->round_n_eval[0] = n_evals, plan->round_keep[0] = 0, plan->round_time[0] = t, plan->round_ckpt[0] = 0;
    plan->round_res[0] = n_resources;
    plan->ckpt_time[0] = t;
    plan->bracket_r0[0] = (double)n_resources;
    return AT2_OK;
}

extern "C" int at2_tpe_sim_f32(const at2_plan* plan, const at2_sim_config* cfg, const at2_tpe_config* tpe,
                               const void* d_tables, int64_t n_reps, int64_t rep_offset, uint64_t seed,
                               const at2_hb_outputs* out, float* d_arm_xy, float* d_obs_loss, int32_t* d_n_injected,
                               void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(plan && cfg && d_tables && out, "at2_tpe_sim_f32: NULL plan/config/tables/outputs");
    AT2_REQUIRE(n_reps > 0 && rep_offset >= 0, "at2_tpe_sim_f32: n_reps must be positive and rep_offset >= 0");
    AT2_REQUIRE(plan->n_arms > 0 && plan->n_arms < 65536 && plan->n_rounds > 0 && plan->n_ckpts > 0 &&
                    plan->n_ckpts <= AT2_MAX_CKPTS, "at2_tpe_sim_f32: plan is empty or malformed");
    AT2_REQUIRE(cfg->n_families >= 1 && cfg->n_families <= AT2_MAX_FAMILIES, "at2_tpe_sim_f32: n_families %d outside [1,%d]",
                cfg->n_families, AT2_MAX_FAMILIES);
    AT2_REQUIRE(cfg->func >= 0 && cfg->func <= AT2_FUNC_RASTRIGIN,
                "at2_tpe_sim_f32: func id %d is not one of the reference's 2-D test functions (TPE suggestions are 2-D)", cfg->func);
    const int has_smooth = at2_cfg_has_smooth(cfg);
    AT2_REQUIRE(!has_smooth || at2_savgol_window(plan->R) <= plan->R,
                "If mode is 'interp', window_length must be less than or equal to the size of x.");
    TpeParams p = {};
    int rc = fill_tpe(tpe, &p.tpe);
    if (rc != AT2_OK) return rc;
    p.plan = *plan;
    int slot[AT2_MAX_FAMILIES];
    p.sp.n_slots = at2_family_slot_map(cfg, slot);
    for (int f = 0; f < cfg->n_families; ++f) {
        p.fam[f].ml = (float)(cfg->fam[f].ml_agg / 100.0);
        p.fam[f].up = (float)cfg->fam[f].up_spik;
        p.fam[f].s0 = (float)cfg->fam[f].start_shift;
        p.fam[f].s1 = (float)cfg->fam[f].end_shift;
        p.fam[f].smooth = cfg->fam[f].is_smooth;
        p.fam[f].pw_off = slot[f] * at2_rec_rows(plan->R);
    }
    p.sp.func = cfg->func, p.sp.F = cfg->n_families, p.sp.scheduler = cfg->scheduler, p.descending = cfg->descending;
    p.sp.x_min = (float)cfg->x_min, p.sp.x_rng = (float)(cfg->x_max - cfg->x_min);
    p.sp.y_min = (float)cfg->y_min, p.sp.y_rng = (float)(cfg->y_max - cfg->y_min);
    p.sp.noise = (float)cfg->init_noise;
    p.sp.key0 = (uint32_t)seed, p.sp.key1 = (uint32_t)(seed >> 32);
    at2_round_keys(seed, p.sp.rk);
    p.tables = (const float*)d_tables;
    p.n_reps = n_reps, p.rep_offset = rep_offset;
    p.ofe = (float*)out->d_ofe, p.ofe_arm = out->d_ofe_arm, p.best_xy = (float*)out->d_best_xy;
    p.round_best = (float*)out->d_round_best, p.round_best_arm = out->d_round_best_arm;
    p.survivors = out->d_survivors, p.ckpt_loss = (float*)out->d_ckpt_loss, p.arm_xy = d_arm_xy, p.obs_loss = d_obs_loss;
    p.n_injected = d_n_injected;
    p.npad = halving_npad(plan);
    p.cap = (plan->n_arms + 31) & ~31;
    const int C = plan->n_ckpts;
    // mixture components [cap + 4] float4, aliased by the halving sort scratch (npad keys + 2 npad ids)
    size_t scratch = 16 * (size_t)(p.cap + 4);
    if (scratch < (size_t)p.npad * (8 + 4)) scratch = (size_t)p.npad * (8 + 4);
    scratch = (scratch + 15) & ~(size_t)15;
    p.scratch_floats = (int)(scratch / 4);
    size_t wb = (size_t)p.cap * ((C + 5) * 4 + 3 * 2 + 2) + 32 * 2 + scratch;
    p.warp_bytes = (int)((wb + 15) & ~(size_t)15);
    At2TableLayout L{plan->R, p.sp.n_slots, has_smooth, at2_sg_stride(C)};
    const size_t fixed = (size_t)((L.total_philox() + 3) & ~3) * 4 + 2 * (AT2_MAX_CKPTS + 4) * sizeof(int);
    int dev = 0, max_smem = 0, sms = 0;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int nwarps = AT2_TPE_MAX_WARPS;
    {
        const int w = at2_debug().tpe_warps_per_cta;
        if (w > 0 && w <= AT2_TPE_MAX_WARPS) nwarps = w;
    }
    while (nwarps > 1 && fixed + (size_t)nwarps * p.warp_bytes > (size_t)max_smem) --nwarps;
    if ((long long)nwarps > n_reps) nwarps = (int)n_reps;
    const size_t smem = fixed + (size_t)nwarps * p.warp_bytes;
    if (smem > (size_t)max_smem) {
        at2_set_error("at2_tpe_sim_f32 needs %zu B of shared memory per CTA (%d arms); device limit %d B", smem, plan->n_arms, max_smem);
        return AT2_ELIMIT;
    }
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * nwarps > 64) per_sm = 64 / nwarps > 0 ? 64 / nwarps : 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (n_reps + nwarps - 1) / nwarps;
    if (grid > need) grid = need;
    cudaStream_t st = (cudaStream_t)stream;
    if (!has_smooth) return launch_tpe<false, 1>(p, (int)grid, nwarps * 32, smem, st);
    if (C <= 4) return launch_tpe<true, 4>(p, (int)grid, nwarps * 32, smem, st);
    if (C == 5) return launch_tpe<true, 5>(p, (int)grid, nwarps * 32, smem, st);   // R = 81, eta = 3: scalar fifth weight (sim_core.cuh)
    if (C <= 8) return launch_tpe<true, 8>(p, (int)grid, nwarps * 32, smem, st);
    return launch_tpe<true, 16>(p, (int)grid, nwarps * 32, smem, st);
}

extern "C" int at2_tpe_score_f32(const float* d_obs, const float* d_loss, int32_t n_obs, int32_t n_problems, double low,
                                 double high, const at2_tpe_config* tpe, const float* d_candidates, int32_t n_candidates,
                                 float* d_scores, float* d_models, float* d_lean_scores, float* d_samples, void* stream) {
    at2_reset_launch_count();
    TpeDev cfg;
    const int rc = fill_tpe(tpe, &cfg);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(d_obs && d_loss && d_candidates && d_scores, "at2_tpe_score_f32: NULL pointer");
    AT2_REQUIRE(n_obs >= 1 && n_obs <= 4096 && n_problems >= 1 && n_candidates >= 1 && n_candidates <= 256 && high > low,
                "at2_tpe_score_f32: bad sizes (1 <= n_obs <= 4096, 1 <= n_candidates <= 256, low < high)");
    const int cap = (n_obs + 31) & ~31;
    const size_t warp_bytes = (((size_t)cap * (3 * 4 + 3 * 2 + 1) + 16 * (size_t)(cap + 4) + 128) + 15) & ~(size_t)15;
    const int nwarps = 4;
    const size_t smem = warp_bytes * nwarps;
    AT2_CUDA_CHECK(cudaFuncSetAttribute(tpe_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (n_problems + nwarps - 1) / nwarps;
    tpe_score_kernel<<<grid, nwarps * 32, smem, (cudaStream_t)stream>>>(d_obs, d_loss, n_obs, (float)low, (float)high, cfg,
                                                                         d_candidates, n_candidates, d_scores, d_models,
                                                                         d_lean_scores, d_samples, cap, n_problems);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}
