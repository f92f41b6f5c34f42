// (a)+(b) with early stopping: the same Hyperband repetitions as hb_sim.cu, but an arm is only simulated as far as the
// round that reads it needs - arms eliminated at r = 1 never leave their first point, the survivor of a bracket is the
// only arm simulated to R.  The reference simulates every arm to R on its first evaluate()
// (benchmarks/opt_function_simulation_problem.py:67-69) although it only ever reads the checkpoints
// ("results are unaffected by lazy simulation"); with counter-based streams per arm the values read are
// bit-identical to the full simulation's, so OFEs, survivor sets and round-bests are the same - at ~1/12 of the epochs.
//
// Mapping: a warp owns a tile of G <= 32 repetitions and walks the plan round by round (G comes from a cost model of the
// job, see the entry point).  A round is worked through in chunks of repetitions: the (repetition, alive arm) pairs of a
// chunk are spread flat over the lanes, each lane re-simulates its arm from its stream's start up to the position the
// round reads (raw arms) or the end of the Savitzky-Golay window of that read-out (smoothed arms - listed behind the raw
// ones so that only their passes run long), and the chunk's losses land in shared memory; then the warp runs that round
// of successive halving - repetition by repetition for wide rounds, 32 / n repetitions side by side for rounds of
// n <= 32 arms.  Nothing is carried between rounds except the survivor ids.
#include <cfloat>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <mutex>

#include "at2_common.cuh"
#include "halving.cuh"
#include "sim_core.cuh"

namespace {

struct LazyParams {
    at2_plan plan;
    Fam<float> fam[AT2_MAX_FAMILIES];
    ArmSpace sp;
    int descending;
    const float* tables;
    long long n_reps, rep_offset, n_tiles;
    int G, npad, widest, warp_bytes, dims;
    int km_a, km_b;   // entries of the two survivor-id buffers of a repetition (see the kernel)
    int cur_cap;      // floats in a warp's loss buffer (>= widest bracket)
    int mixed;        // the families hold both raw and smoothed ones
    // set-up cache (global scratch, one slab per warp of the grid): start value, target and family of every arm of the tile
    // at hand, written when a bracket's first round draws the arm, read by the later rounds that simulate it again
    float* ws_f1;
    float* ws_fn;
    unsigned char* ws_fam;
    long long ws_stride;   // entries per warp (>= G * n_arms)
    float* ofe;
    int* ofe_arm;
    float* best_xy;
    float* round_best;
    int* round_best_arm;
    int* survivors;
    // fused all-gather of the OFEs (see hb_sim.cu): repetition i also lands in element gather_base + i of every gather[r]
    float* gather[AT2_MAX_PEERS];
    int n_gather, gather_multicast;
    long long gather_base;
};

__host__ __device__ inline size_t lazy_keys_offset(int cur_cap, int G) { return ((size_t)cur_cap * 4 + (size_t)G * 12 + 7) & ~(size_t)7; }

// CACHE: the launch has a slab of the set-up cache per warp (see LazyParams); without one (allocation refused, or the
// debug option lazy_no_cache) every round redraws its arms - same results, ~10 % more work on the six-family job.
template <bool HAS_SMOOTH, bool GTAB, bool CACHE>
__global__ void __launch_bounds__(256, 3) hb_lazy_kernel(const __grid_constant__ LazyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = p.plan.R, A = p.plan.n_arms, C = p.plan.n_ckpts, F = p.sp.F, G = p.G;
    const At2TableLayout L{R, p.sp.n_slots, HAS_SMOOTH ? 1 : 0, at2_sg_stride(C)};
    const int tab_elems = GTAB ? 0 : (L.total_philox() + 3) & ~3;   // GTAB: tables stay in global memory (sim_core.cuh tab128)
    float* s_tab = reinterpret_cast<float*>(smem_raw);
    int* s_ckpos = reinterpret_cast<int*>(s_tab + tab_elems);
    int* s_sglast = s_ckpos + AT2_MAX_CKPTS + 4;
    unsigned char* warp_base = reinterpret_cast<unsigned char*>(s_sglast + AT2_MAX_CKPTS + 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if constexpr (!GTAB)
        for (int i = tid; i < L.total_philox(); i += blockDim.x) s_tab[i] = p.tables[i];
    if (tid < AT2_MAX_CKPTS + 4) s_ckpos[tid] = tid < C ? p.plan.ckpt_time[tid] - 1 : INT_MAX;
    __syncthreads();
    const float* s_rec = (GTAB ? p.tables : s_tab) + L.off_rec();
    const float* s_sg = (GTAB ? p.tables : s_tab) + L.off_sg();
    if (tid < AT2_MAX_CKPTS + 4) {
        int last = tid < C ? p.plan.ckpt_time[tid] - 1 : 0;
        if (HAS_SMOOTH && tid < C)
            for (int t = 0; t < R; ++t)
                if (s_sg[t * L.sg_stride + tid] != 0.0f) last = max(last, t);
        s_sglast[tid] = last;
    }
    __syncthreads();
    const SimTables tb{GTAB ? (uint32_t)(L.off_rec() * 4) : (uint32_t)__cvta_generic_to_shared(s_rec),
                       GTAB ? (uint32_t)(L.off_sg() * 4) : (uint32_t)__cvta_generic_to_shared(s_sg), s_ckpos, R, L.sg_stride,
                       reinterpret_cast<const unsigned char*>(p.tables)};
    // warp-private: losses of the current round chunk [cur_cap], survivor ids [G][km_a + km_b], per-repetition state, sort scratch
    unsigned char* mine = warp_base + (size_t)warp * p.warp_bytes;
    float* cur = reinterpret_cast<float*>(mine);                    // [cur_cap] losses of the round chunk at hand
    float* best = cur + p.cur_cap;                                  // [G] running OFE
    int* best_arm = reinterpret_cast<int*>(best + G);               // [G]
    int* surv_w = best_arm + G;                                     // [G] survivors written so far
    Key32* keys = reinterpret_cast<Key32*>(mine + lazy_keys_offset(p.cur_cap, G));   // [npad]
    // Survivor ids of a repetition alternate between two buffers: round i of a bracket reads the one round i - 1 wrote.
    // Buffer A receives the survivors of a bracket's rounds 0, 2, 4, ..., buffer B those of rounds 1, 3, ... - a third of
    // A's at eta = 3 - so they are sized separately (km_a, km_b entries).
    unsigned short* ids = reinterpret_cast<unsigned short*>(keys + p.npad);   // [G][km_a + km_b]
    const int kms = p.km_a + p.km_b;
    auto idbuf = [&](int g, int fl) { return ids + g * kms + (fl ? 0 : p.km_a); };   // fl = 1: buffer A
    const bool desc = p.descending != 0;
    constexpr bool cache = CACHE;
    const size_t wbase = (size_t)((long long)blockIdx.x * nwarps + warp) * (size_t)p.ws_stride;
    // an arm's start value (before the read-at-R shortcut), target and family: drawn from its stream + base function, or -
    // in the rounds after the one that drew it - read back from the warp's slab of the set-up cache
    auto draw_item = [&](bool active, int g, int arm, unsigned long long gid, bool store, float& f1, float& fn, int& fam_idx) {
        const ArmDraw d = draw_arm(p.sp, gid, arm);
        const Fam<float> fam = p.fam[d.fam];
        const float f = base_function(p.sp.func, d.x, d.y, d.x2, d.y2);
        fn = f - fam.s1;
        f1 = (f + p.sp.noise * d.z) - fam.s0;
        fam_idx = d.fam;
        if (store && active) {
            const size_t o = wbase + (size_t)g * A + arm;
            __stcg(p.ws_f1 + o, f1), __stcg(p.ws_fn + o, fn), p.ws_fam[o] = (unsigned char)d.fam;
        }
    };
    auto load_item = [&](bool active, int g, int arm, float& f1, float& fn, int& fam_idx) {
        const size_t o = wbase + (size_t)(active ? g : 0) * A + (active ? arm : 0);
        f1 = __ldcg(p.ws_f1 + o), fn = __ldcg(p.ws_fn + o), fam_idx = active ? (int)__ldcg(p.ws_fam + o) : 0;
    };

    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < p.n_tiles; tile += (long long)gridDim.x * nwarps) {
        const long long rep0 = tile * G;
        const int g_eff = (int)min((long long)G, p.n_reps - rep0);
        for (int g = lane; g < G; g += 32) best_arm[g] = -1, surv_w[g] = 0, best[g] = 0.0f;
        __syncwarp();
        for (int b = 0; b < p.plan.n_brackets; ++b) {
            const int first = p.plan.bracket_first_arm[b];
            int n_cur = p.plan.bracket_n[b];
            int flip = 0;
            for (int rd = p.plan.bracket_first_round[b]; rd < p.plan.bracket_first_round[b + 1]; ++rd) {
                const int c = p.plan.round_ckpt[rd];
                const bool first_round = rd == p.plan.bracket_first_round[b];
                const int keep = min(p.plan.round_keep[rd], n_cur);
                // ---- simulate the alive arms of the whole tile up to what this round reads ----
                // A smoothed arm is read through its Savitzky-Golay window, so it has to run further than a raw arm read at
                // the same checkpoint (R = 81, window 19: 19 steps against 0 at r = 1).  When the families mix both kinds
                // and this round's two stopping points differ, the warp first lists a chunk's items raw-first (the family of
                // an arm is one Philox block away) and runs the all-raw passes through the lean loop, to the raw stopping
                // point; only the smoothed tail runs long.  The list aliases the sort scratch (idle until the halving below).
                // A raw arm read at R stands on its target f_n by construction (P = P11 = 1 at the last step,
                // core/simulation_evaluator.py:105-107,113-115; sim_core.cuh mt_attempt lands on it exactly): no run at all.
                const bool raw_at_target = s_ckpos[c] == R - 1;
                const int stop_raw = raw_at_target ? 1 : s_ckpos[c] + 1, stop_sm = HAS_SMOOTH ? s_sglast[c] + 1 : s_ckpos[c] + 1;
                const bool split = HAS_SMOOTH && p.mixed && stop_sm > stop_raw;
                unsigned short* order = reinterpret_cast<unsigned short*>(keys);
                const int cap = p.npad * 4;                  // list entries; cur_cap <= cap
                // The round is worked through in chunks of as many repetitions as the loss buffer holds (cur_cap floats: a
                // few repetitions of a bracket's wide first round, the whole tile of a narrow late round - which is what
                // lets a tile be 32 repetitions wide, so that the long single-arm rounds still fill the warp).
                const int rpc = min(g_eff, p.cur_cap / n_cur);
                // q / n_cur for item numbers q < 65536 by one multiply-high (exact: q * n_cur < 2^32)
                const uint32_t div_magic = n_cur > 1 ? 0xffffffffu / (uint32_t)n_cur + 1u : 0u;
                auto div_n = [&](int q) { return n_cur > 1 ? (int)__umulhi((uint32_t)q, div_magic) : q; };
#pragma unroll 1
                for (int g0 = 0; g0 < g_eff; g0 += rpc) {
                    const int gc = min(rpc, g_eff - g0);     // repetitions g0 .. g0 + gc - 1; losses at cur[(g - g0) * n_cur + j]
                    const int n_items = gc * n_cur;
                    int n_raw = 0;
                    if (split) {
                        int n_sm = 0;
#pragma unroll 1
                        for (int base = 0; base < n_items; base += 32) {
                            const int q = base + lane;
                            const bool act = q < n_items;
                            const int gl = act ? div_n(q) : 0, j = act ? q - gl * n_cur : 0, g = g0 + gl;
                            const int arm = first_round ? first + j : (int)idbuf(g, flip)[j];
                            const unsigned long long gid = (unsigned long long)(p.rep_offset + rep0 + g) * (unsigned long long)A + arm;
                            int fam_idx;
                            if (cache) {        // the whole set-up happens here (first round) or comes from the cache
                                float f1c, fnc;
                                if (first_round)
                                    draw_item(act, g, arm, gid, true, f1c, fnc, fam_idx);
                                else      // the list only needs the family
                                    fam_idx = act ? (int)__ldcg(p.ws_fam + wbase + (size_t)g * A + arm) : 0;
                            } else {
                                fam_idx = arm_family(p.sp, gid, arm);
                            }
                            const bool sm = act && p.fam[fam_idx].smooth != 0;
                            const unsigned m_raw = __ballot_sync(0xffffffffu, act && !sm), m_sm = __ballot_sync(0xffffffffu, sm);
                            const unsigned below = (1u << lane) - 1u;
                            if (act) order[sm ? cap - 1 - (n_sm + __popc(m_sm & below)) : n_raw + __popc(m_raw & below)] = (unsigned short)q;
                            n_raw += __popc(m_raw), n_sm += __popc(m_sm);
                        }
                        __syncwarp();
                    }
#pragma unroll 1
                    for (int base = 0; base < n_items; base += 32) {
                        const int qq = base + lane;
                        const bool active = qq < n_items;
                        const int q = !active ? 0 : (split ? (int)order[qq < n_raw ? qq : cap - 1 - (qq - n_raw)] : qq);
                        const int gl = div_n(q), j = q - gl * n_cur, g = g0 + gl;
                        const int arm = first_round ? first + j : (int)idbuf(g, flip)[j];
                        const unsigned long long gid = (unsigned long long)(p.rep_offset + rep0 + g) * (unsigned long long)A + arm;
                        float f1d, fn;
                        int fam_idx;
                        if (cache && (split || !first_round)) {     // first round + split: the listing pass above wrote the slab
                            load_item(active, g, arm, f1d, fn, fam_idx);
                        } else {
                            draw_item(active, g, arm, gid, cache, f1d, fn, fam_idx);
                        }
                        const Fam<float> fam = p.fam[fam_idx];
                        const bool smooth = HAS_SMOOTH && fam.smooth != 0;
                        const float f1 = (raw_at_target && !smooth) ? fn : f1d;
                        float* dst = cur + q;
                        if (!HAS_SMOOTH || (split && base + 32 <= n_raw))
                            philox_trajectory<float, false, 1, true, GTAB>(tb, active, false, f1, fn, fam.up, fam.pw_off, gid,
                                                                           p.sp.rk, stop_raw, dst, 0, C, c);
                        else
                            philox_trajectory<float, HAS_SMOOTH, 1, true, GTAB>(tb, active, smooth, f1, fn, fam.up, fam.pw_off,
                                                                                gid, p.sp.rk, smooth ? stop_sm : stop_raw, dst, 0, C, c);
                    }
                    __syncwarp();
                    // ---- this round of successive halving for the chunk's repetitions ----
                    if (n_cur <= 32) {
                        // narrow rounds: 32 / n_cur repetitions side by side in one warp, each segment of n_cur lanes ranks
                        // its arms by counting (same stable order as warp_rank_by_counting); the lane that ranks first IS the
                        // round-best of its repetition and does that repetition's bookkeeping
                        const int S = 32 / n_cur, seg = div_n(lane), j = lane - seg * n_cur;
                        const unsigned seg_mask = (n_cur == 32 ? 0xffffffffu : (1u << n_cur) - 1u) << (seg < S ? seg * n_cur : 0);
#pragma unroll 1
                        for (int h0 = 0; h0 < gc; h0 += S) {
                            const bool act = seg < S && h0 + seg < gc;
                            const int gl = act ? h0 + seg : 0, g = g0 + gl;
                            const long long rep_g = rep0 + g;
                            unsigned short* out = idbuf(g, flip ^ 1);
                            const int my_id = first_round ? first + j : (int)idbuf(g, flip)[j];
                            const float my_loss = cur[gl * n_cur + j];
                            const uint32_t ord = act ? orderable(my_loss, desc) : 0xffffffffu;
                            int rank = 0;
#pragma unroll 3
                            for (int k = 0; k < n_cur; ++k) rank += __shfl_sync(0xffffffffu, ord, seg * n_cur + k) < ord ? 1 : 0;
                            rank += __popc(__match_any_sync(0xffffffffu, ord) & seg_mask & ((1u << lane) - 1u));
                            const int sw = surv_w[g];
                            __syncwarp();
                            if (act && rank < keep) {
                                out[rank] = (unsigned short)my_id;
                                if (p.survivors != nullptr) p.survivors[rep_g * p.plan.n_surv + sw + rank] = my_id;
                            }
                            if (act && rank == 0) {
                                surv_w[g] = sw + keep;
                                if (p.round_best != nullptr) p.round_best[rep_g * p.plan.n_rounds + rd] = my_loss;
                                if (p.round_best_arm != nullptr) p.round_best_arm[rep_g * p.plan.n_rounds + rd] = my_id;
                                const int ba = best_arm[g];
                                const float bl = best[g];
                                if (ba < 0 || (desc ? my_loss > bl : my_loss < bl)) best[g] = my_loss, best_arm[g] = my_id;
                            }
                            __syncwarp();
                        }
                    } else {
#pragma unroll 1
                        for (int gl = 0; gl < gc; ++gl) {
                            const int g = g0 + gl;
                            const long long rep_g = rep0 + g;
                            const float* loss = cur + gl * n_cur;
                            const unsigned short* in = idbuf(g, flip);
                            unsigned short* out = idbuf(g, flip ^ 1);
                            float rb_loss;
                            int rb_arm;
                            warp_halving_round_f32(n_cur, keep, first_round, first, in, out, keys, desc, lane,
                                                   [&](int j, int) { return loss[j]; }, rb_loss, rb_arm);
                            __syncwarp();
                            const int sw = surv_w[g];
                            if (p.survivors != nullptr)
                                for (int j = lane; j < keep; j += 32) p.survivors[rep_g * p.plan.n_surv + sw + j] = (int)out[j];
                            if (lane == 0) {
                                surv_w[g] = sw + keep;
                                if (p.round_best != nullptr) p.round_best[rep_g * p.plan.n_rounds + rd] = rb_loss;
                                if (p.round_best_arm != nullptr) p.round_best_arm[rep_g * p.plan.n_rounds + rd] = rb_arm;
                                const int ba = best_arm[g];
                                const float bl = best[g];
                                if (ba < 0 || (desc ? rb_loss > bl : rb_loss < bl)) best[g] = rb_loss, best_arm[g] = rb_arm;
                            }
                            __syncwarp();
                        }
                    }
                }
                flip ^= 1;
                n_cur = keep;
            }
        }
        for (int g = lane; g < g_eff; g += 32) {
            const long long rep_g = rep0 + g;
            if (p.ofe != nullptr) p.ofe[rep_g] = best[g];
            if (p.ofe_arm != nullptr) p.ofe_arm[rep_g] = best_arm[g];
            if (p.gather_multicast)
                multimem_st_f32(p.gather[0] + p.gather_base + rep_g, best[g]);
            else
                for (int r = 0; r < p.n_gather; ++r) p.gather[r][p.gather_base + rep_g] = best[g];
            if (p.best_xy != nullptr) {
                const unsigned long long gid = (unsigned long long)(p.rep_offset + rep_g) * (unsigned long long)A + best_arm[g];
                const ArmDraw d = draw_arm(p.sp, gid, best_arm[g]);
                float* bxy = p.best_xy + (size_t)p.dims * rep_g;
                bxy[0] = d.x, bxy[1] = d.y;
                if (p.dims == 4) bxy[2] = d.x2, bxy[3] = d.y2;
            }
        }
        __syncwarp();
    }
}

// Stream-ordered scratch for the set-up cache: a pool of the library's own per device that keeps what it has handed out
// (release threshold = everything), so a launch's cudaMallocFromPoolAsync / cudaFreeAsync pair costs microseconds after the
// first call and launches on different streams never share a slab.
static void* lazy_scratch_alloc(size_t bytes, cudaStream_t st) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (pools[dev] == nullptr) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool = nullptr;
            if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pools[dev] = pool;
        }
    }
    void* ptr = nullptr;
    if (cudaMallocFromPoolAsync(&ptr, bytes, pools[dev], st) != cudaSuccess) {
        cudaGetLastError();      // no scratch: the kernel redraws the arms instead (same results)
        return nullptr;
    }
    return ptr;
}

template <bool HS, bool GTAB = false>
int launch_lazy(const LazyParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    auto kern = hb_lazy_kernel<HS, GTAB, true>;
    auto kern_nc = hb_lazy_kernel<HS, GTAB, false>;
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern_nc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern_nc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // persistent grid: never more CTAs than are resident at once (registers can bind before shared memory does)
    int by_regs = 0, dev = 0, sms = 0;
    if (at2_resident_by_regs(kern, block, &by_regs) != AT2_OK) return AT2_ECUDA;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if ((long long)by_regs * sms < grid) grid = by_regs * sms;
    // set-up cache: {f1, fn} as floats + the family as a byte per (repetition of the tile, arm), one slab per warp of the grid
    LazyParams q = p;
    void* ws = nullptr;
    int rounds_max = 0;
    for (int b = 0; b < p.plan.n_brackets; ++b)
        rounds_max = max(rounds_max, p.plan.bracket_first_round[b + 1] - p.plan.bracket_first_round[b]);
    if (rounds_max > 1 && !at2_debug().lazy_no_cache) {
        const long long warps = (long long)grid * (block / 32);
        const long long stride = (((long long)p.G * p.plan.n_arms) + 15) & ~15ll;
        ws = lazy_scratch_alloc((size_t)(warps * stride) * 9, st);
        if (ws != nullptr) {
            q.ws_stride = stride;
            q.ws_f1 = reinterpret_cast<float*>(ws);
            q.ws_fn = q.ws_f1 + warps * stride;
            q.ws_fam = reinterpret_cast<unsigned char*>(q.ws_fn + warps * stride);
        }
    }
    if (ws != nullptr)
        kern<<<grid, block, smem, st>>>(q);
    else
        kern_nc<<<grid, block, smem, st>>>(q);
    const cudaError_t launch_rc = cudaGetLastError();
    if (ws != nullptr) cudaFreeAsync(ws, st);     // stream-ordered: after the kernel
    AT2_CUDA_CHECK(launch_rc);
    at2_count_launch(1);
    return AT2_OK;
}

}  // namespace

static int early_stop_impl(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                           int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, const at2_gather* gather,
                           void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(plan && cfg && d_tables && out, "at2_hb_sim_early_stop_f32: NULL plan/config/tables/outputs");
    AT2_REQUIRE(n_reps > 0 && rep_offset >= 0, "at2_hb_sim_early_stop_f32: n_reps must be positive and rep_offset >= 0");
    AT2_REQUIRE(plan->n_arms > 0 && plan->n_arms < 65536 && plan->n_rounds > 0 && plan->n_ckpts > 0 &&
                    plan->n_ckpts <= AT2_MAX_CKPTS, "at2_hb_sim_early_stop_f32: plan is empty or malformed");
    AT2_REQUIRE(cfg->n_families >= 1 && cfg->n_families <= AT2_MAX_FAMILIES, "at2_hb_sim_early_stop_f32: n_families %d outside [1,%d]",
                cfg->n_families, AT2_MAX_FAMILIES);
    AT2_REQUIRE(at2_func_dims(cfg->func) > 0, "at2_hb_sim_early_stop_f32: unknown func id %d", cfg->func);
    AT2_REQUIRE(out->d_ckpt_loss == nullptr, "at2_hb_sim_early_stop_f32: checkpoint losses of unread arms do not exist in this mode");
    const int has_smooth = at2_cfg_has_smooth(cfg);
    AT2_REQUIRE(!has_smooth || at2_savgol_window(plan->R) <= plan->R,
                "If mode is 'interp', window_length must be less than or equal to the size of x.");
    LazyParams p = {};
    if (gather != nullptr) {
        const int grc = at2_check_gather(gather, "at2_hb_sim_early_stop_gather_f32");
        if (grc != AT2_OK) return grc;
        for (int r = 0; r < gather->n_targets; ++r) p.gather[r] = gather->d_target[r];
        p.n_gather = gather->n_targets, p.gather_multicast = gather->is_multicast, p.gather_base = gather->base;
    }
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
    p.dims = at2_func_dims(cfg->func);
    {
        int n_sm = 0;
        for (int f = 0; f < cfg->n_families; ++f) n_sm += cfg->fam[f].is_smooth != 0;
        p.mixed = n_sm > 0 && n_sm < cfg->n_families && !at2_debug().hb_no_split;
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
    p.round_best = (float*)out->d_round_best, p.round_best_arm = out->d_round_best_arm, p.survivors = out->d_survivors;
    p.npad = halving_npad(plan);
    int widest = 0, km_a = 1, km_b = 1;
    for (int b = 0; b < plan->n_brackets; ++b) {
        widest = plan->bracket_n[b] > widest ? plan->bracket_n[b] : widest;
        for (int rd = plan->bracket_first_round[b]; rd < plan->bracket_first_round[b + 1]; ++rd) {
            int& km = ((rd - plan->bracket_first_round[b]) & 1) ? km_b : km_a;
            km = plan->round_keep[rd] > km ? plan->round_keep[rd] : km;
        }
    }
    p.widest = widest, p.km_a = km_a, p.km_b = km_b;
    int dev = 0, max_smem = 0, sms = 0;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // Loss buffer of a warp: the widest bracket has to fit; beyond that, 512 floats keep every chunk of a round at 15+
    // warp passes (and it never exceeds the 4 * npad entries of the raw-first list that shares the sort scratch).
    p.cur_cap = widest > 512 ? widest : 512;
    if (p.cur_cap > 4 * p.npad) p.cur_cap = 4 * p.npad;
    const int C = plan->n_ckpts;
    At2TableLayout L{plan->R, p.sp.n_slots, has_smooth, at2_sg_stride(C)};
    const size_t tab_bytes = (size_t)((L.total_philox() + 3) & ~3) * 4, fixed_small = 2 * (AT2_MAX_CKPTS + 4) * sizeof(int);
    const bool force_gtab = at2_debug().force_gtab != 0;
    // launch geometry of tile width g: a warp's tile state (loss buffer, per-repetition state, sort scratch, survivor ids),
    // tables in shared or global memory, warps per CTA, CTAs per SM
    struct Geo {
        size_t warp_bytes, smem;
        bool gtab;
        int nwarps, per_sm;
    };
    auto geometry = [&](int g) {
        Geo q;
        q.warp_bytes = (lazy_keys_offset(p.cur_cap, g) + (size_t)p.npad * 8 + (size_t)g * (km_a + km_b) * 2 + 15) & ~(size_t)15;
        // plans whose tables do not fit beside two warps' state read them from global memory
        q.gtab = tab_bytes + fixed_small + 2 * q.warp_bytes > (size_t)max_smem || force_gtab;
        const size_t fixed = q.gtab ? fixed_small : tab_bytes + fixed_small;
        q.nwarps = 8;
        while (q.nwarps > 1 && fixed + (size_t)q.nwarps * q.warp_bytes > (size_t)max_smem) q.nwarps >>= 1;
        q.smem = fixed + (size_t)q.nwarps * q.warp_bytes;
        q.per_sm = (int)((size_t)(227 * 1024) / (q.smem + 1024));
        if (q.per_sm < 1) q.per_sm = 1;
        if (q.per_sm * q.nwarps > 24) q.per_sm = 24 / q.nwarps > 0 ? 24 / q.nwarps : 1;   // registers: 24 warps of 80
        return q;
    };
    // Tile width: up to 32 repetitions (the single-arm last rounds then fill the warp).  The width whose busiest warp
    // finishes first under a coarse cost model - warp passes x steps read + per-pass set-up + halving operations, scaled by
    // what a warp's share of its scheduler is worth at that width's residency - so a small job is cut into narrow tiles
    // that cover the machine, a large one into wide ones, and a plan with wide brackets (whose survivor lists eat shared
    // memory) trades resident warps for fuller lanes only where that pays.
    int G = 32;
    if ((long long)G > n_reps) G = (int)n_reps;
    if (at2_debug().lazy_reps_per_tile > 0) {
        G = at2_debug().lazy_reps_per_tile;
        if ((long long)G > n_reps) G = (int)n_reps;
        while (G > 1 && geometry(G).warp_bytes > 40 * 1024) G = (G + 1) / 2;
    } else {
        double best_cost = -1.0;
        int best_g = 1;
        for (int g = G; g >= 1; --g) {
            const Geo q = geometry(g);
            if (g > 1 && q.warp_bytes > 40 * 1024) continue;
            double cost = 0.0;
            for (int b = 0; b < plan->n_brackets; ++b) {
                int n_cur = plan->bracket_n[b];
                for (int rd = plan->bracket_first_round[b]; rd < plan->bracket_first_round[b + 1]; ++rd) {
                    const int keep = plan->round_keep[rd] < n_cur ? plan->round_keep[rd] : n_cur;
                    const int steps = plan->ckpt_time[plan->round_ckpt[rd]] - 1;
                    cost += (double)((g * n_cur + 31) / 32) * (steps + 6);                     // trajectories + arm set-up
                    cost += n_cur <= 32 ? (double)((g + 32 / n_cur - 1) / (32 / n_cur)) * (0.5 + n_cur / 16.0)
                                        : (double)g * (n_cur <= 128 ? 15.0 : 40.0);              // halving
                    n_cur = keep;
                }
            }
            const long long warps = (long long)sms * q.per_sm * q.nwarps;
            const long long tiles = (n_reps + g - 1) / g, rounds = (tiles + warps - 1) / warps;
            // a warp alone on its scheduler issues ~0.18 instructions per cycle, w of them together 1 - 0.82^w (six: 0.7,
            // as ncu shows): a few wide tiles run faster per warp than many narrow ones
            const double per_sched = (double)(tiles < warps ? tiles : warps) / (4.0 * sms);
            cost *= (double)rounds * per_sched / (1.0 - pow(0.82, per_sched));
            if (best_cost < 0.0 || cost < best_cost) best_cost = cost, best_g = g;
        }
        G = best_g;
    }
    const Geo geo = geometry(G);
    p.G = G;
    p.warp_bytes = (int)geo.warp_bytes;
    p.n_tiles = (n_reps + G - 1) / G;
    const bool gtab = geo.gtab;
    const int nwarps = geo.nwarps;
    int per_sm = geo.per_sm;
    const size_t smem = geo.smem;
    if (smem > (size_t)max_smem) {
        at2_set_error("at2_hb_sim_early_stop_f32 needs %zu B of shared memory per CTA; device limit %d B", smem, max_smem);
        return AT2_ELIMIT;
    }
    long long grid = (long long)sms * per_sm;
    const long long need = (p.n_tiles + nwarps - 1) / nwarps;
    if (grid > need) grid = need;
    cudaStream_t st = (cudaStream_t)stream;
    if (gtab)
        return has_smooth ? launch_lazy<true, true>(p, (int)grid, nwarps * 32, smem, st)
                          : launch_lazy<false, true>(p, (int)grid, nwarps * 32, smem, st);
    return has_smooth ? launch_lazy<true>(p, (int)grid, nwarps * 32, smem, st)
                      : launch_lazy<false>(p, (int)grid, nwarps * 32, smem, st);
}

extern "C" int at2_hb_sim_early_stop_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                                         int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, void* stream) {
    return early_stop_impl(plan, cfg, d_tables, n_reps, rep_offset, seed, out, nullptr, stream);
}

extern "C" int at2_hb_sim_early_stop_gather_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables,
                                                int64_t n_reps, int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out,
                                                const at2_gather* gather, void* stream) {
    AT2_REQUIRE(gather != nullptr, "at2_hb_sim_early_stop_gather_f32: gather is NULL");
    return early_stop_impl(plan, cfg, d_tables, n_reps, rep_offset, seed, out, gather, stream);
}
