// (a)+(b) fused: Gamma-process trajectory simulation + Hyperband successive halving, one launch for all repetitions.
//
// Work decomposition (B200: 148 SMs, 4 SMSPs each, no HBM traffic to speak of - the kernel is bound by instruction issue,
// its smoothed passes by shared-memory wavefronts):
//   * a warp (or a team of warps) owns a *tile* of G consecutive repetitions; its G*A arm trajectories are spread flat over the lanes
//     (one lane = one arm at a time) so lane occupancy does not depend on A being a multiple of 32;
//   * every lane runs its own Philox4x32-10 stream -> Box-Muller -> Marsaglia-Tsang attempt loop.  A rejected attempt
//     simply does not advance that lane's epoch, so lanes drift apart by a few epochs instead of the whole warp
//     re-drawing whenever any lane rejects; per-epoch constants come from shared-memory tables indexed per lane;
//   * only the losses at the Hyperband checkpoints are kept (shared memory, C values per arm); smooth families
//     accumulate the Savitzky-Golay dot products of those C read-outs on the fly, so no trajectory is ever stored;
//   * when the families mix raw and smoothed arms, a warp first lists its tile's arms raw-first and runs the all-raw
//     warp passes through the lean trajectory loop (no dot products);
//   * successive halving then runs per repetition inside one warp, in registers (halving.cuh): rounds of <= 32 arms are
//     ranked by counting, wider rounds sorted by a bitonic network over 4 or 8 (loss, position) keys per lane with 64-bit
//     shuffles; the float64 parity path and rounds wider than 256 use a shared-memory bitonic network.
//
// Reference behaviour implemented (paths relative to the reference root):
//   core/simulation_evaluator.py:36-48,84-116   Gamma draw + recurrence
//   core/simulation_evaluator.py:76-82          .fs read-out (raw or Savitzky-Golay)
//   benchmarks/opt_function_simulation_problem.py:50-86,91-118   f_1 / f_n, int(n_resources) read-out
//   benchmarks/opt_function_problem.py:49-147   test functions;  core/arm.py:44-55, core/params.py:36 arm draw
//   core/shape_family_scheduler.py:50-77        family choice
//   optimisers/sequential/hyperband_optimiser.py:46-57,74-106    stable top-k, round-best, OFE (core/optimiser.py:67-68)
#include <cfloat>
#include <climits>
#include <atomic>
#include <cstdlib>

#include "at2_common.cuh"
#include "halving.cuh"
#include "sim_core.cuh"

namespace {

template <typename T>
struct Params {
    at2_plan plan;
    Fam<T> fam[AT2_MAX_FAMILIES];
    ArmSpace sp;
    int descending;
    int dims;         // hyperparameters per arm (2; 4 for the egg4 extension)
    const T* tables;
    long long n_reps, rep_offset;
    int G;            // repetitions per tile
    long long n_tiles;
    long long n_tiles_full;   // tiles [0, n_tiles_full) hold G repetitions; the rest of the job is cut into 1-repetition tiles
    int team_warps;   // warps cooperating on one tile (1 = warp-private tiles, no block-level barriers at all)
    int dyn_slot;     // >= 0: warp-private tiles are handed out through the device counter pair g_sched[dyn_slot]
    int tile_halving; // float path, warp-private tiles: the tile's repetitions are halved together (halving.cuh tile_halving_f32)
    int km_a, km_b;   // entries of a repetition's two survivor-id buffers in that mode
    int split_cap;    // > 0: raw and smooth arms of a tile run in separate warp passes; entries of a warp's order list
    int sort_warps;   // warps of a team that ever run halving = min(team_warps, G); only they own sort scratch
    int n_teams;      // teams per CTA
    int npad;         // power of two >= widest bracket
    int stride;       // G * n_arms rounded up to 32
    int team_bytes;   // shared-memory bytes per team
    // pre-drawn inputs
    const T* f1;
    const T* fn;
    const int* family;
    const T* gamma;
    const T* xy;
    // fused all-gather of the OFEs: repetition i of this call also lands in element gather_base + i of every gather[r]
    // (the peers' symmetric buffers, written over NVLink), or of the one NVSwitch multicast address that fans out to all
    T* gather[AT2_MAX_PEERS];
    int n_gather, gather_multicast;
    long long gather_base;
    // outputs
    T* ofe;
    int* ofe_arm;
    T* best_xy;
    T* round_best;
    int* round_best_arm;
    int* survivors;
    T* ckpt_loss;
};

// Dynamic tile schedule (warp-private tiles).  The warp scheduler's arbitration is not fair (ncu: 4.8 of 6 warps per
// scheduler resident on average with a static round-robin schedule - the favoured warps finish their share early and the
// SM runs its tail half empty), so a warp takes its next tile from a device counter instead: all warps finish within one
// tile of each other.  Tile ids are handed out in increasing order - the G-repetition tiles first, the single repetitions
// (launch_one) last.  Which warp runs a tile does not matter: streams are keyed by the repetition id.  A launch owns one
// slot of the counter array (round-robin on the host; launches that overlap in time must hold different slots, 64 slots);
// the last warp to run dry resets its slot, so a slot is zero whenever it is handed out.
struct SchedSlot {
    unsigned long long next;   // next tile id
    unsigned int done;         // warps that have run dry
    unsigned int pad;
};
#define AT2_SCHED_SLOTS 64
__device__ SchedSlot g_sched[AT2_SCHED_SLOTS];
static std::atomic<unsigned> g_launch_seq{0};   // one sequence for every kernel variant: consecutive launches, distinct slots

// Register cap 80 (65536 / 768): shared memory (a 3-repetition tile per warp) holds about 24 warps per SM anyway - three
// 8-warp CTAs when the constant tables are small, ONE 24-warp CTA when they are large (six families: 18 KB per CTA), see
// choose_launch.  Measured -4 % against a 64-register cap on the configurations with smoothed families.
// capture-free second half of the trajectory loop (sim_core.cuh philox_trajectory SPLIT) per loop variant
#ifndef AT2_SPLIT_LEAN
#define AT2_SPLIT_LEAN true
#endif
#ifndef AT2_SPLIT_SMOOTH
#define AT2_SPLIT_SMOOTH true
#endif
#ifndef AT2_HB_MAX_WARPS
#define AT2_HB_MAX_WARPS 24
#endif
template <typename T, int MODE, bool HAS_SMOOTH, int CK, bool GTAB>
__global__ void __launch_bounds__(AT2_HB_MAX_WARPS * 32, 1) hb_sim_kernel(const __grid_constant__ Params<T> p) {
    typedef typename KeyOf<T>::type K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = p.plan.R, A = p.plan.n_arms, C = p.plan.n_ckpts, F = p.sp.F;
    const At2TableLayout L{R, p.sp.n_slots, HAS_SMOOTH ? 1 : 0, at2_sg_stride(C)};
    // GTAB: the tables stay in global memory (plans too large for shared memory), see sim_core.cuh tab128
    // the Philox kernels stage records + Savitzky-Golay rows; the pre-drawn paths also the folded-step table (pw)
    const int tab_total = MODE == MODE_PHILOX ? L.total_philox() : L.total();
    const int tab_elems = GTAB ? 0 : (tab_total + 3) & ~3;
    T* s_tab = reinterpret_cast<T*>(smem_raw);
    int* s_ckpos = reinterpret_cast<int*>(s_tab + tab_elems);       // [AT2_MAX_CKPTS + 4] positions, then INT_MAX
    unsigned char* team_base = reinterpret_cast<unsigned char*>(s_ckpos + AT2_MAX_CKPTS + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = warp / p.team_warps, wteam = warp - team * p.team_warps;
    if constexpr (!GTAB)
        for (int i = tid; i < tab_total; i += blockDim.x) s_tab[i] = p.tables[i];
    if (tid < AT2_MAX_CKPTS + 4) s_ckpos[tid] = tid < C ? p.plan.ckpt_time[tid] - 1 : INT_MAX;
    __syncthreads();
    const T* s_rec = (GTAB ? p.tables : s_tab) + L.off_rec();
    const T* s_sg = (GTAB ? p.tables : s_tab) + L.off_sg();
    const T* s_pw = (GTAB ? p.tables : s_tab) + L.off_pw();
    const SimTables tb{GTAB ? (uint32_t)(L.off_rec() * sizeof(T)) : (uint32_t)__cvta_generic_to_shared(s_rec),
                       GTAB ? (uint32_t)(L.off_sg() * sizeof(T)) : (uint32_t)__cvta_generic_to_shared(s_sg), s_ckpos, R,
                       L.sg_stride, reinterpret_cast<const unsigned char*>(p.tables),
                       // the end value is the last checkpoint: everything after the second-to-last one runs capture-free
                       p.plan.ckpt_time[C - 1] == R ? (C >= 2 ? p.plan.ckpt_time[C - 2] : 1) : INT_MAX};

    unsigned char* my_team = team_base + (size_t)team * p.team_bytes;
    T* s_ck = reinterpret_cast<T*>(my_team);                                          // [C][stride]
    K* keys = reinterpret_cast<K*>(s_ck + (size_t)C * p.stride) + (size_t)wteam * p.npad;   // [sort_warps][npad]
    unsigned short* ids_a = reinterpret_cast<unsigned short*>(reinterpret_cast<K*>(s_ck + (size_t)C * p.stride) +
                                                              (size_t)p.sort_warps * p.npad) + (size_t)wteam * 2 * p.npad;
    unsigned short* ids_b = ids_a + p.npad;
    const bool desc = p.descending != 0;

    const long long team_global = (long long)blockIdx.x * p.n_teams + team;
    const long long team_stride = (long long)gridDim.x * p.n_teams;
    const bool dynamic = p.dyn_slot >= 0;        // host: only with warp-private tiles
    auto next_tile = [&](long long cur) -> long long {
        if (!dynamic) return cur < 0 ? team_global : cur + team_stride;
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&g_sched[p.dyn_slot].next, 1ull);
        return (long long)__shfl_sync(0xffffffffu, t, 0);
    };
    for (long long tile = next_tile(-1); tile < p.n_tiles; tile = next_tile(tile)) {
        // whole rounds of G-repetition tiles, then the remainder of the job as single repetitions: the teams' shares then
        // differ by one repetition at most, not by one tile (see launch_one)
        const bool full = tile < p.n_tiles_full;
        const long long rep0 = full ? tile * p.G : p.n_tiles_full * p.G + (tile - p.n_tiles_full);
        const int g_eff = full ? p.G : 1;
        const int n_items = g_eff * A;

        // ---------------- phase 1: trajectories -> checkpoint losses in shared memory ----------------
        if constexpr (MODE == MODE_PHILOX) {
            // Families that mix raw and Savitzky-Golay-smoothed arms: a smoothed arm accumulates C dot products along its
            // trajectory and in a mixed warp pass every lane pays for them.  So this warp first lists its items with the raw
            // arms in front and the smoothed ones at the back (the family of an arm is one Philox block away), then runs the
            // passes that hold only raw arms through the lean loop.  The list aliases the halving sort scratch.  Every arm is
            // simulated by the same arithmetic either way: outputs do not change by a bit.  Without the list the warp's items
            // run in their natural order.  Either way there is ONE call site per loop variant (lean / smoothed): the
            // trajectory loops are most of the kernel's code and instruction fetch is a measured limiter (sim_core.cuh).
            const bool listed = HAS_SMOOTH && p.split_cap > 0;
            const int cap = p.split_cap;
            unsigned short* order = reinterpret_cast<unsigned short*>(reinterpret_cast<K*>(s_ck + (size_t)C * p.stride)) +
                                    (size_t)wteam * cap;
            int n_raw = 0, n_sm = 0;
            if (listed) {
                for (int base = wteam * 32; base < n_items; base += p.team_warps * 32) {
                    const int i = base + lane;
                    const bool act = i < n_items;
                    const int ii = act ? i : 0;
                    const int rep_l = ii / A, arm = ii - rep_l * A;
                    const unsigned long long gid = (unsigned long long)(p.rep_offset + rep0 + rep_l) * (unsigned long long)A + arm;
                    const bool sm = act && p.fam[arm_family(p.sp, gid, arm)].smooth != 0;
                    const unsigned m_raw = __ballot_sync(0xffffffffu, act && !sm), m_sm = __ballot_sync(0xffffffffu, sm);
                    const unsigned below = (1u << lane) - 1u;
                    if (act) order[sm ? cap - 1 - (n_sm + __popc(m_sm & below)) : n_raw + __popc(m_raw & below)] = (unsigned short)i;
                    n_raw += __popc(m_raw), n_sm += __popc(m_sm);
                }
                __syncwarp();
            }
            // this warp's share of the tile: `n_mine` items, the q-th of which is item_of(q)
            const int my_groups = n_items > wteam * 32 ? (n_items - wteam * 32 + p.team_warps * 32 - 1) / (p.team_warps * 32) : 0;
            const int n_mine = listed ? n_raw + n_sm : my_groups * 32;
            for (int k = 0; k < n_mine; k += 32) {
                const int q = k + lane;
                int i;
                if (listed)
                    i = q < n_mine ? (int)order[q < n_raw ? q : cap - 1 - (q - n_raw)] : n_items;
                else
                    i = (k / 32) * p.team_warps * 32 + wteam * 32 + lane;
                const bool active = i < n_items;
                const int ii = active ? i : 0;
                const int rep_l = ii / A, arm = ii - rep_l * A;
                const unsigned long long gid = (unsigned long long)(p.rep_offset + rep0 + rep_l) * (unsigned long long)A + arm;
                const ArmDraw d = draw_arm(p.sp, gid, arm);
                const Fam<T> fam = p.fam[d.fam];
                const float f = base_function(p.sp.func, d.x, d.y, d.x2, d.y2);
                const float fn = f - (float)fam.s1;                          // opt_function_simulation_problem.py:63
                const float f1 = (f + p.sp.noise * d.z) - (float)fam.s0;     // :62,:64
                if (!HAS_SMOOTH || (listed && k + 32 <= n_raw))
                    philox_trajectory<T, false, 1, false, GTAB, AT2_SPLIT_LEAN>(tb, active, false, f1, fn, (float)fam.up,
                                                                                fam.pw_off, gid, p.sp.rk, R, s_ck + ii,
                                                                                p.stride, C);
                else
                    philox_trajectory<T, HAS_SMOOTH, CK, false, GTAB, AT2_SPLIT_SMOOTH>(tb, active, HAS_SMOOTH && fam.smooth != 0,
                                                                                        f1, fn, (float)fam.up, fam.pw_off, gid,
                                                                                        p.sp.rk, R, s_ck + ii, p.stride, C);
            }
        } else {
            for (int base = wteam * 32; base < n_items; base += p.team_warps * 32) {
                const int i = base + lane;
                const bool active = i < n_items;
                const int ii = active ? i : 0;
                const int rep_l = ii / A, arm = ii - rep_l * A;
                const long long rep_g = rep0 + rep_l;                   // index into per-call arrays
                const long long a_g = rep_g * A + arm;
                const T f1 = active ? p.f1[a_g] : (T)0;
                const T fn = active ? p.fn[a_g] : (T)0;
                const Fam<T> fam = p.fam[active ? p.family[a_g] : 0];
                const bool smooth = HAS_SMOOTH && fam.smooth != 0;
                T f = f1;
                int nc = 0;               // next checkpoint to capture (raw arms; smooth arms never capture)
                int next_pos = (smooth || !active) ? INT_MAX : s_ckpos[0];
                T acc[CK];
#pragma unroll
                for (int c = 0; c < CK; ++c) acc[c] = HAS_SMOOTH ? s_sg[c] * f : (T)0;  // sg[0][c] * fs[0]
                if (active && next_pos == 0) {
                    s_ck[ii] = f;
                    nc = 1;
                    next_pos = s_ckpos[1];
                }
                const T* grow = p.gamma + (rep_g * A + arm) * (long long)(R - 1);
                for (int tt = 1; tt < R; ++tt) {
                    if (!active) break;
                    const T g = grow[tt - 1];
                    if constexpr (sizeof(T) == 8) {
                        const T P = s_rec[AT2_REC * (fam.pw_off + tt) + 4], P11 = s_rec[AT2_REC * (fam.pw_off + tt) + 5];
                        f = (T)step_exact((double)f, (double)g, tt, R, (double)fn, (double)fam.ml, (double)fam.up,
                                          (double)P, (double)P11);
                    } else {
                        // the production arithmetic (sim_core.cuh mt_attempt) on the given Gamma draw
                        const T* pw = s_pw + AT2_PW * (fam.pw_off + tt);
                        f = (T)step_folded((float)f, (float)g, (float)fn, (float)fam.up, (float)pw[0], (float)pw[1],
                                           (float)pw[2]);
                    }
                    if (HAS_SMOOTH && smooth) {
#pragma unroll
                        for (int c = 0; c < CK; ++c) acc[c] = fma(s_sg[tt * L.sg_stride + c], f, acc[c]);
                    }
                    if (tt == next_pos) {
                        s_ck[nc * p.stride + ii] = f;
                        ++nc;
                        next_pos = s_ckpos[nc];
                    }
                }
                if (HAS_SMOOTH) {
                    if (active && smooth) {
#pragma unroll
                        for (int c = 0; c < CK; ++c)
                            if (c < C) s_ck[c * p.stride + ii] = acc[c];
                    }
                }
            }
        }
        team_barrier(team, p.team_warps);

        if (p.ckpt_loss != nullptr) {
            for (int i = wteam * 32 + lane; i < n_items * C; i += p.team_warps * 32) {
                const int a = i / C, c = i - a * C;
                p.ckpt_loss[(rep0 * A + a) * C + c] = s_ck[c * p.stride + a];
            }
        }

        // ---------------- phase 2: successive halving ----------------
        // warp-private float tiles: the tile's repetitions together, narrow rounds side by side (the per-repetition state
        // lives where the per-repetition path keeps its two id lists); otherwise one warp per repetition
        float* t_best = nullptr;
        int* t_best_arm = nullptr;
        if constexpr (sizeof(T) == 4) {
            if (p.tile_halving) {
                const int n_ids = (p.G * (p.km_a + p.km_b) + 1) & ~1;
                t_best = reinterpret_cast<float*>(ids_a + n_ids);
                t_best_arm = reinterpret_cast<int*>(t_best + p.G);
                HalvingOut<float> ho{p.ofe, p.ofe_arm, p.round_best, p.round_best_arm, p.survivors};
                tile_halving_f32(p.plan, s_ck, p.stride, A, g_eff, keys, ids_a, p.km_a, p.km_b, t_best, t_best_arm,
                                 t_best_arm + p.G, desc, lane, rep0, ho);
            }
        }
        for (int rep_l = wteam; rep_l < g_eff; rep_l += p.team_warps) {
            const long long rep_g = rep0 + rep_l;
            T ofe_value;
            int best_arm;
            if (t_best != nullptr) {
                ofe_value = (T)t_best[rep_l], best_arm = t_best_arm[rep_l];
            } else {
                HalvingOut<T> ho{p.ofe, p.ofe_arm, p.round_best, p.round_best_arm, p.survivors};
                best_arm = warp_halving<T>(p.plan, s_ck + rep_l * A, p.stride, keys, ids_a, ids_b, desc, lane, rep_g, ho,
                                           &ofe_value);
            }
            if constexpr (sizeof(T) == 4) {
                if (p.gather_multicast) {       // one multimem store, replicated by the switch
                    if (lane == 0) multimem_st_f32(reinterpret_cast<float*>(p.gather[0]) + p.gather_base + rep_g, (float)ofe_value);
                } else if (lane < p.n_gather) {
                    p.gather[lane][p.gather_base + rep_g] = ofe_value;   // one peer per lane
                }
            }
            if (lane == 0 && p.best_xy != nullptr) {
                if constexpr (MODE == MODE_PHILOX) {   // re-draw the winner's hyperparameters from its stream
                    const unsigned long long gid = (unsigned long long)(p.rep_offset + rep_g) * (unsigned long long)A + best_arm;
                    const ArmDraw d = draw_arm(p.sp, gid, best_arm);
                    T* bxy = p.best_xy + (size_t)p.dims * rep_g;
                    bxy[0] = (T)d.x, bxy[1] = (T)d.y;
                    if (p.dims == 4) bxy[2] = (T)d.x2, bxy[3] = (T)d.y2;
                } else {
                    const long long a_g = rep_g * A + best_arm;
                    p.best_xy[2 * rep_g] = p.xy != nullptr ? p.xy[2 * a_g] : (T)0;
                    p.best_xy[2 * rep_g + 1] = p.xy != nullptr ? p.xy[2 * a_g + 1] : (T)0;
                }
            }
        }
        team_barrier(team, p.team_warps);
    }
    if (dynamic && lane == 0) {
        // this warp has run dry; the last one to do so hands the slot back zeroed
        __threadfence();
        if (atomicAdd(&g_sched[p.dyn_slot].done, 1u) == (unsigned)team_stride - 1u) {
            g_sched[p.dyn_slot].next = 0ull;
            g_sched[p.dyn_slot].done = 0u;
            __threadfence();
        }
    }
}

// ---- launch plumbing ---------------------------------------------------------------------------------------------------
template <typename T>
struct LaunchCfg {
    int G, team_warps, n_teams, block, grid, npad, stride, team_bytes;
    int sms;
    int gtab;   // tables read from global memory instead of being staged into shared memory
    size_t smem;
};

static int opt_int(int v, int dflt) { return v > 0 ? v : dflt; }   // a debug option, or the library's own choice

template <typename T>
int choose_launch(const at2_plan* plan, int has_smooth, int F /* record blocks */, long long n_reps, int mode, LaunchCfg<T>* out) {
    typedef typename KeyOf<T>::type K;
    const int A = plan->n_arms, C = plan->n_ckpts, R = plan->R;
    int widest = 0;
    for (int b = 0; b < plan->n_brackets; ++b) widest = plan->bracket_n[b] > widest ? plan->bracket_n[b] : widest;
    int npad = 64;
    while (npad < widest) npad <<= 1;
    At2TableLayout L{R, F, has_smooth, at2_sg_stride(C)};
    size_t fixed = (size_t)(((mode == MODE_PHILOX ? L.total_philox() : L.total()) + 3) & ~3) * sizeof(T) +
                   (AT2_MAX_CKPTS + 4) * sizeof(int);
    const size_t sort_per_warp = (size_t)npad * (sizeof(K) + 2 * sizeof(unsigned short));
    int dev = 0, max_smem = 0, sms = 0;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // A team of `tw` warps owns tiles of G repetitions.  Pick (tw, G): every warp of the team gets the same number
    // of 32-arm groups, lanes are well packed, and the team's shared memory stays small enough for >= ~32 resident
    // warps per SM.  Small problems end up with tw = 1: warp-private tiles and no block-level barrier at all.
    const size_t per_warp_budget = 9 * 1024;
    int best_tw = 1, best_g = 1;
    double best_score = -1.0;
    int km_a = 1, km_b = 1;
    halving_id_buffer_sizes(plan, &km_a, &km_b);
    // can one warp hold a whole repetition at all?  (R = 243: 369 arms x 6 checkpoints do not fit the per-warp budget)
    const bool solo_fits = (size_t)C * ((A + 31) & ~31) * sizeof(T) + sort_per_warp <= per_warp_budget;
    for (int tw = 1; tw <= 8; tw *= 2) {
        for (int g = 1; g <= 64; ++g) {
            const int items = g * A;
            const int groups = (items + 31) / 32;
            const int per_warp = (groups + tw - 1) / tw;
            const double packing = (double)items / ((double)per_warp * tw * 32);
            const int stride = (items + 31) & ~31;
            // halving is one warp per repetition: only min(tw, g) warps of a team ever sort
            const size_t team_bytes = (size_t)C * stride * sizeof(T) + (tw < g ? tw : g) * sort_per_warp;
            if (team_bytes > per_warp_budget * tw && !(tw == 8 && g == 1)) continue;
            // halving wants at least one repetition per warp of the team
            const double sh_balance = (double)g / ((double)((g + tw - 1) / tw) * tw);
            // a team of several warps gives up the dynamic tile schedule, the tile-wide halving and the barrier-free
            // trajectory phase: where one warp can hold a repetition, a team is worth it only for a clear packing gain
            // (R = 27: team of 2 x 16 repetitions 0.89 ms, one warp x 7 repetitions 0.75 ms); where it cannot, the team
            // shapes compete among themselves as before (R = 243: 4 warps x 3 repetitions)
            double score = packing * (0.9 + 0.1 * sh_balance) - (solo_fits ? 0.03 * (tw - 1) : 0.002 * tw);
            // warp-private float tiles halve their repetitions together only while the tile's survivor lists fit the sort
            // scratch (fill_params: tile_halving); one repetition more than that falls back to per-repetition halving,
            // which costs more than a slightly emptier last warp pass (measured on R = 27: 8 repetitions 0.89 ms, 7: 0.75)
            if (sizeof(T) == 4 && tw == 1 && (size_t)(((g * (km_a + km_b) + 1) & ~1) * 2 + g * 12) > (size_t)npad * 4) score *= 0.85;
            if (score > best_score + 1e-9) best_score = score, best_tw = tw, best_g = g;
        }
    }
    if (best_score < 0) best_tw = 8, best_g = 1;
    best_tw = opt_int(at2_debug().hb_team_warps, best_tw);
    best_g = opt_int(at2_debug().hb_reps_per_tile, best_g);
    if ((long long)best_g > n_reps) best_g = (int)n_reps;
    const int stride = (best_g * A + 31) & ~31;
    size_t team_bytes = (size_t)C * stride * sizeof(T) + (best_tw < best_g ? best_tw : best_g) * sort_per_warp;
    team_bytes = (team_bytes + 15) & ~(size_t)15;
    // plans whose tables do not fit beside one team's state (R ~ 1000 with six families) read them from global memory
    out->gtab = (fixed + team_bytes > (size_t)max_smem || at2_debug().force_gtab != 0) ? 1 : 0;
    if (out->gtab) fixed = (AT2_MAX_CKPTS + 4) * sizeof(int);
    // Warps per CTA: every CTA carries its own copy of the constant tables (`fixed`), so with large tables fewer, bigger
    // CTAs keep more warps resident.  Take the size with the most resident warps per SM (shared memory, 80 registers per
    // thread, 2048 threads), the larger one on ties (the tile schedule is dynamic, so small CTAs balance nothing; measured:
    // one 24-warp CTA per SM against three 8-warp ones is even on R = 81 and 7 % faster on R = 27).
    const long long n_tiles = (n_reps + best_g - 1) / best_g;
    int warps_per_cta = 8 < best_tw ? best_tw : 8, best_resident = -1;
    for (int wpc = warps_per_cta; wpc <= AT2_HB_MAX_WARPS; wpc += 4) {
        if (wpc % best_tw != 0 || (best_tw > 1 && wpc / best_tw > 15)) continue;     // named barriers 1..15 for the teams
        const size_t need = fixed + (size_t)(wpc / best_tw) * team_bytes;
        if (need > (size_t)max_smem) break;
        int ctas = (int)((size_t)(227 * 1024) / (need + 1024));
        const int by_regs = 65536 / (80 * 32 * wpc), by_threads = 2048 / (wpc * 32);
        ctas = ctas < by_regs ? ctas : by_regs;
        ctas = ctas < by_threads ? ctas : by_threads;
        if (ctas * wpc >= best_resident) best_resident = ctas * wpc, warps_per_cta = wpc;
    }
    warps_per_cta = opt_int(at2_debug().hb_warps_per_cta, warps_per_cta);
    if (warps_per_cta < best_tw) warps_per_cta = best_tw;
    if (warps_per_cta > AT2_HB_MAX_WARPS) warps_per_cta = AT2_HB_MAX_WARPS;
    int n_teams = warps_per_cta / best_tw;
    while (n_teams > 1 && (fixed + n_teams * team_bytes > (size_t)max_smem || (long long)(n_teams - 1) * sms >= n_tiles)) --n_teams;
    const size_t smem = fixed + n_teams * team_bytes;
    if (smem > (size_t)max_smem) {
        at2_set_error("hyperband simulation needs %zu B of shared memory per CTA (R=%d, %d arms, %d checkpoints); device limit %d B",
                      smem, R, A, C, max_smem);
        return AT2_ELIMIT;
    }
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int max_by_threads = 2048 / (n_teams * best_tw * 32);
    if (per_sm > max_by_threads) per_sm = max_by_threads;
    long long grid = (long long)sms * per_sm;
    const long long ctas_needed = (n_tiles + n_teams - 1) / n_teams;
    if (grid > ctas_needed) grid = ctas_needed;
    out->sms = sms;
    out->G = best_g, out->team_warps = best_tw, out->n_teams = n_teams, out->block = n_teams * best_tw * 32;
    out->grid = (int)grid, out->npad = npad, out->stride = stride, out->team_bytes = (int)team_bytes, out->smem = smem;
    return AT2_OK;
}

template <typename T, int MODE, bool HS, int CK, bool GTAB = false>
int launch_one(const Params<T>& p, const LaunchCfg<T>& lc, cudaStream_t st) {
    auto kern = hb_sim_kernel<T, MODE, HS, CK, GTAB>;
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // persistent grid: exactly the CTAs that are resident at once (registers can bind before shared memory does), so
    // every CTA walks the same number of tiles and there is no half-empty second wave
    int by_regs = 0;
    if (at2_resident_by_regs(kern, lc.block, &by_regs) != AT2_OK) return AT2_ECUDA;
    int grid = lc.grid;
    if ((long long)by_regs * lc.sms < grid) grid = by_regs * lc.sms;
    // Tile schedule.  A static round-robin over equal tiles leaves some teams one whole tile (G repetitions, ~1/9 of their
    // work at 10^5 repetitions) more than others.  So only whole rounds are dealt out as G-repetition tiles; what is left
    // (< teams * G + G repetitions) goes out as single repetitions.  Repetition ids do not change, results neither.
    Params<T> q = p;
    const long long teams = (long long)grid * q.n_teams, full_tiles = q.n_reps / q.G;
    q.dyn_slot = (q.team_warps == 1 && !at2_debug().hb_static_schedule) ? (int)(g_launch_seq.fetch_add(1) % AT2_SCHED_SLOTS) : -1;
    if (q.dyn_slot >= 0) {
        // dynamic: G-repetition tiles, then `singles` single repetitions per warp to even out the finish
        const int singles = opt_int(at2_debug().hb_singles, 1);
        const long long body = q.n_reps - teams * singles;
        q.n_tiles_full = body > 0 ? body / q.G : 0;
        if (q.n_tiles_full > full_tiles) q.n_tiles_full = full_tiles;
    } else {
        q.n_tiles_full = at2_debug().hb_equal_tiles ? full_tiles : (full_tiles / teams) * teams;
    }
    q.n_tiles = q.n_tiles_full + (q.n_reps - q.n_tiles_full * q.G);
    kern<<<grid, lc.block, lc.smem, st>>>(q);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}

template <typename T, int MODE>
int launch_dispatch(const Params<T>& p, const LaunchCfg<T>& lc, int has_smooth, cudaStream_t st) {
    if (lc.gtab) return has_smooth ? launch_one<T, MODE, true, 16, true>(p, lc, st) : launch_one<T, MODE, false, 1, true>(p, lc, st);
    if (!has_smooth) return launch_one<T, MODE, false, 1>(p, lc, st);
    const int C = p.plan.n_ckpts;
    if (C <= 4) return launch_one<T, MODE, true, 4>(p, lc, st);
    if (C == 5) return launch_one<T, MODE, true, 5>(p, lc, st);
    if (C <= 6) return launch_one<T, MODE, true, 6>(p, lc, st);
    if (C <= 8) return launch_one<T, MODE, true, 8>(p, lc, st);
    return launch_one<T, MODE, true, 16>(p, lc, st);
}

template <typename T>
int fill_params(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, long long n_reps,
                const at2_hb_outputs* out, int mode, Params<T>* p, LaunchCfg<T>* lc) {
    AT2_REQUIRE(plan && cfg && d_tables && out, "at2_hb_sim: NULL plan/config/tables/outputs");
    AT2_REQUIRE(n_reps > 0, "at2_hb_sim: n_reps must be positive (got %lld)", n_reps);
    AT2_REQUIRE(plan->n_arms > 0 && plan->n_rounds > 0 && plan->n_ckpts > 0 && plan->n_ckpts <= AT2_MAX_CKPTS,
                "at2_hb_sim: plan is empty or malformed");
    AT2_REQUIRE(plan->n_arms < 65536, "at2_hb_sim: more than 65535 arms per repetition");
    AT2_REQUIRE(cfg->n_families >= 1 && cfg->n_families <= AT2_MAX_FAMILIES, "at2_hb_sim: n_families %d outside [1,%d]",
                cfg->n_families, AT2_MAX_FAMILIES);
    AT2_REQUIRE(at2_func_dims(cfg->func) > 0, "at2_hb_sim: unknown func id %d", cfg->func);
    const int has_smooth = at2_cfg_has_smooth(cfg);
    AT2_REQUIRE(!has_smooth || at2_savgol_window(plan->R) <= plan->R,
                "If mode is 'interp', window_length must be less than or equal to the size of x.");
    p->plan = *plan;
    int slot[AT2_MAX_FAMILIES];
    p->sp.n_slots = at2_family_slot_map(cfg, slot);
    for (int f = 0; f < cfg->n_families; ++f) {
        p->fam[f].ml = sizeof(T) == 8 ? (T)cfg->fam[f].ml_agg : (T)(cfg->fam[f].ml_agg / 100.0);
        p->fam[f].up = (T)cfg->fam[f].up_spik;
        p->fam[f].s0 = (T)cfg->fam[f].start_shift;
        p->fam[f].s1 = (T)cfg->fam[f].end_shift;
        p->fam[f].smooth = cfg->fam[f].is_smooth;
        p->fam[f].pw_off = slot[f] * at2_rec_rows(plan->R);
    }
    p->dims = at2_func_dims(cfg->func);
    p->sp.func = cfg->func, p->sp.F = cfg->n_families, p->sp.scheduler = cfg->scheduler, p->descending = cfg->descending;
    p->sp.x_min = (float)cfg->x_min, p->sp.x_rng = (float)(cfg->x_max - cfg->x_min);
    p->sp.y_min = (float)cfg->y_min, p->sp.y_rng = (float)(cfg->y_max - cfg->y_min);
    p->sp.noise = (float)cfg->init_noise;
    p->tables = (const T*)d_tables;
    p->n_reps = n_reps;
    p->ofe = (T*)out->d_ofe, p->ofe_arm = out->d_ofe_arm, p->best_xy = (T*)out->d_best_xy;
    p->round_best = (T*)out->d_round_best, p->round_best_arm = out->d_round_best_arm;
    p->survivors = out->d_survivors, p->ckpt_loss = (T*)out->d_ckpt_loss;
    const int rc = choose_launch<T>(plan, has_smooth, p->sp.n_slots, n_reps, mode, lc);
    if (rc != AT2_OK) return rc;
    p->G = lc->G, p->npad = lc->npad, p->stride = lc->stride;
    p->team_warps = lc->team_warps, p->n_teams = lc->n_teams, p->team_bytes = lc->team_bytes;
    p->sort_warps = lc->team_warps < lc->G ? lc->team_warps : lc->G;
    // raw / smoothed arms in separate warp passes when the families mix both and a warp's list and noted classes fit in the
    // team's sort scratch, which they alias (see the kernel); debug option hb_no_split disables it
    const int cap = ((lc->stride / 32 + lc->team_warps - 1) / lc->team_warps) * 32;
    const size_t scratch = (size_t)p->sort_warps * lc->npad * (sizeof(typename KeyOf<T>::type) + 2 * sizeof(unsigned short));
    int n_smooth = 0;
    for (int f = 0; f < cfg->n_families; ++f) n_smooth += cfg->fam[f].is_smooth ? 1 : 0;
    p->split_cap = (mode == MODE_PHILOX && n_smooth > 0 && n_smooth < cfg->n_families &&
                    (size_t)cap * 2 * lc->team_warps <= scratch && lc->stride < 65536 && !at2_debug().hb_no_split) ? cap : 0;
    p->n_tiles = (n_reps + lc->G - 1) / lc->G;
    // float path with warp-private tiles: the repetitions of a tile are halved together when their survivor-id buffers and
    // running state fit where the per-repetition path keeps its two id lists (2 * npad entries);
    // debug option hb_no_tile_halving disables it (tests)
    halving_id_buffer_sizes(plan, &p->km_a, &p->km_b);
    const size_t tile_state = (size_t)((lc->G * (p->km_a + p->km_b) + 1) & ~1) * 2 + (size_t)lc->G * 12;
    p->tile_halving = (sizeof(T) == 4 && lc->team_warps == 1 && tile_state <= (size_t)lc->npad * 4 &&
                       !at2_debug().hb_no_tile_halving) ? 1 : 0;
    return AT2_OK;
}

}  // namespace

extern "C" int at2_hb_sim_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                              int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, void* stream) {
    at2_reset_launch_count();
    Params<float> p = {};
    LaunchCfg<float> lc;
    const int rc = fill_params<float>(plan, cfg, d_tables, n_reps, out, MODE_PHILOX, &p, &lc);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(rep_offset >= 0, "at2_hb_sim_f32: rep_offset must be >= 0");
    p.rep_offset = rep_offset;
    p.sp.key0 = (uint32_t)seed, p.sp.key1 = (uint32_t)(seed >> 32);
    at2_round_keys(seed, p.sp.rk);
    return launch_dispatch<float, MODE_PHILOX>(p, lc, at2_cfg_has_smooth(cfg), (cudaStream_t)stream);
}

extern "C" int at2_hb_sim_gather_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                                     int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, const at2_gather* gather,
                                     void* stream) {
    at2_reset_launch_count();
    Params<float> p = {};
    LaunchCfg<float> lc;
    const int rc = fill_params<float>(plan, cfg, d_tables, n_reps, out, MODE_PHILOX, &p, &lc);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(rep_offset >= 0, "at2_hb_sim_gather_f32: rep_offset must be >= 0");
    const int grc = at2_check_gather(gather, "at2_hb_sim_gather_f32");
    if (grc != AT2_OK) return grc;
    for (int r = 0; r < gather->n_targets; ++r) p.gather[r] = gather->d_target[r];
    p.n_gather = gather->n_targets, p.gather_multicast = gather->is_multicast, p.gather_base = gather->base;
    p.rep_offset = rep_offset;
    p.sp.key0 = (uint32_t)seed, p.sp.key1 = (uint32_t)(seed >> 32);
    at2_round_keys(seed, p.sp.rk);
    return launch_dispatch<float, MODE_PHILOX>(p, lc, at2_cfg_has_smooth(cfg), (cudaStream_t)stream);
}

extern "C" int at2_hb_sim_predrawn_f64(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables,
                                       int64_t n_reps, const double* d_f1, const double* d_fn, const int32_t* d_family,
                                       const double* d_gamma, const double* d_xy, const at2_hb_outputs* out,
                                       void* stream) {
    at2_reset_launch_count();
    Params<double> p = {};
    LaunchCfg<double> lc;
    const int rc = fill_params<double>(plan, cfg, d_tables, n_reps, out, MODE_PREDRAWN, &p, &lc);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(d_f1 && d_fn && d_family && d_gamma, "at2_hb_sim_predrawn_f64: NULL input array");
    p.f1 = d_f1, p.fn = d_fn, p.family = d_family, p.gamma = d_gamma, p.xy = d_xy;
    return launch_dispatch<double, MODE_PREDRAWN>(p, lc, at2_cfg_has_smooth(cfg), (cudaStream_t)stream);
}

extern "C" int at2_hb_sim_predrawn_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables,
                                       int64_t n_reps, const float* d_f1, const float* d_fn, const int32_t* d_family,
                                       const float* d_gamma, const float* d_xy, const at2_hb_outputs* out,
                                       void* stream) {
    at2_reset_launch_count();
    Params<float> p = {};
    LaunchCfg<float> lc;
    const int rc = fill_params<float>(plan, cfg, d_tables, n_reps, out, MODE_PREDRAWN, &p, &lc);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(d_f1 && d_fn && d_family && d_gamma, "at2_hb_sim_predrawn_f32: NULL input array");
    p.f1 = d_f1, p.fn = d_fn, p.family = d_family, p.gamma = d_gamma, p.xy = d_xy;
    return launch_dispatch<float, MODE_PREDRAWN>(p, lc, at2_cfg_has_smooth(cfg), (cudaStream_t)stream);
}

// ---- FP32 FMA micro-benchmark (roofline denominator measured in the same run) -----------------------------------------
// 16 independent accumulator chains per thread; variant 0: a = a*m + c with m, c in registers (3 distinct sources),
// variant 1: a = a*m + a (2 distinct sources), variant 2: a = a*1.0000001f + 1e-9f (immediate operands).  The caller
// takes the best of the variants as the FFMA peak.
template <int VARIANT>
__global__ void fma_probe_kernel(long long iters, float* sink) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = threadIdx.x * 1e-3f + k;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (VARIANT == 0) a[k] = fmaf(a[k], m, c);
            if (VARIANT == 1) a[k] = fmaf(a[k], m, a[k]);
            if (VARIANT == 2) a[k] = fmaf(a[k], 0.99999988f, 1e-9f);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456f) sink[0] = s;  // never true; keeps the chains alive
}

extern "C" int at2_fma_peak_probe(int32_t blocks, int32_t threads, int64_t iters, int32_t variant, float* d_sink, void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && d_sink, "at2_fma_peak_probe: bad argument");
    AT2_REQUIRE(variant >= 0 && variant <= 2, "at2_fma_peak_probe: variant must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 0) fma_probe_kernel<0><<<blocks, threads, 0, st>>>(iters, d_sink);
    if (variant == 1) fma_probe_kernel<1><<<blocks, threads, 0, st>>>(iters, d_sink);
    if (variant == 2) fma_probe_kernel<2><<<blocks, threads, 0, st>>>(iters, d_sink);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}
