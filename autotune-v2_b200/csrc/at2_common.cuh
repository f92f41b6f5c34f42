// Shared definitions for libat2_b200 (sm_100a).  See include/at2.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "at2.h"

// ---- error plumbing (thread-local message, no exceptions across the ABI) -------------------------------------------
void at2_set_error(const char* fmt, ...);
void at2_count_launch(int n);
void at2_reset_launch_count();

#define AT2_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            at2_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AT2_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

#define AT2_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            at2_set_error(__VA_ARGS__); \
            return AT2_EINVAL;          \
        }                               \
    } while (0)

// ---- debug options (include/at2.h at2_debug_option): tile / CTA shapes and code-path switches used by tests and
// measurements.  Process-wide, all zero by default (= the library decides); never read from the environment.
struct At2Debug {
    int hb_reps_per_tile, hb_team_warps, hb_warps_per_cta;   // > 0: override choose_launch (hb_sim.cu)
    int hb_no_split, hb_equal_tiles, hb_no_tile_halving, hb_static_schedule, hb_singles;
    int force_gtab;                                           // constant tables from global memory on any plan
    int lazy_reps_per_tile;                                   // hb_lazy.cu tile width
    int lazy_no_cache;                                        // hb_lazy.cu: redraw the arms every round (no set-up cache)
    int tpe_warps_per_cta;                                    // tpe_sim.cu
    int kf_warps_per_cta, kf_reps_per_tile;                   // nn_lookup.cu hb_knownfn_kernel
};
const At2Debug& at2_debug();

// CTAs of `kern` resident per SM as far as registers and threads go (shared memory is budgeted by the caller): a
// persistent grid must not exceed what is resident at once, or the surplus CTAs form a half-empty second wave.
template <typename Kern>
static int at2_resident_by_regs(Kern kern, int block, int* out) {
    cudaFuncAttributes fa;
    AT2_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
    const int warps = (block + 31) / 32;
    const int regs_per_warp = ((fa.numRegs * 32 + 511) / 512) * 512;
    int by_regs = regs_per_warp > 0 ? 65536 / (regs_per_warp * warps) : 32;
    const int by_threads = 2048 / (warps * 32);
    if (by_regs > by_threads) by_regs = by_threads;
    *out = by_regs < 1 ? 1 : by_regs;
    return AT2_OK;
}

// ---- constant-table layout shared by host builder and kernels (units: reals) --------------------------------------
// rec: one 12-real record per (family block, step) - families with equal (ml_agg, nec_agg) share a block, see
//      at2_family_slot_map; F below counts blocks: row r = f * rows + t with t in [0,R] and rows = at2_rec_rows(R) (R + 1
//      made odd, see below).  Row t = 0 of a family is unused; row R is a sentinel whose c is NaN, so the acceptance test
//      of a finished lane can never pass and no bounds check is needed in the inner loop.
//   float tables (Philox paths): two 128-bit loads from one address (+0, +16 bytes), 16 bytes spare:
//          {d, c, d*log2(e), d*log2(e) + 32 | +-w, K1w, K0, P11 | 0, 0, 0, 0}
//          d, c: Marsaglia-Tsang constants of the Gamma draw of step t; w = d * scale (Gamma draw g = w v^3), its SIGN BIT
//          flags the steps whose position is a checkpoint of the plan the table was built for (the trajectory loop stores
//          the loss of a flagged step, sim_core.cuh); K1w = w ml (1 - P), K0 = P - 2 ml (1 - P), ml = ml_agg / 100,
//          P = (t/(R-1))**nec, P11 = (t/(R-1))**(1.1*nec): the folded down step k = v^3 K1w + K0 (sim_core.cuh mt_attempt).
//          Why 48-byte records: a 128-bit shared-memory load is served per quarter warp, conflict-free when its distinct
//          16-byte words fall into distinct word slots modulo 8.  With 32-byte records each of the two loads used only
//          every other slot (ncu, six families: the first use of a record was the top stall); a stride of three slots
//          walks through all eight, so lanes spread over up to 8 consecutive steps of one family never collide, and the
//          odd `rows` puts the same step of different families into different slots.
//   double tables (pre-drawn parity path): {d, c, d*log2(e), d*scale, P, P11, 1-P, 0, 0, 0, 0, 0}
// sg[t][c]  t in [0,R], c in [0,sg_stride): Savitzky-Golay weight of raw[t] in the smoothed value at checkpoint c
//          (present only when some family is smooth; row R is zero).  sg_stride = 12 (the record stride, so that the row of
//          a lane's step is its record offset away from sg[0]) up to 12 checkpoints, 24 above.
// pwt[r] 4 reals {K1 = ml (1 - P), K0, P11, 0}: the folded step on a given Gamma draw (pre-drawn float path only; the
//          Philox kernels do not stage it).
#define AT2_REC 12
#define AT2_PW 4
__host__ __device__ inline int at2_sg_stride(int n_ckpts) { return n_ckpts <= 12 ? 12 : 24; }
__host__ __device__ inline int at2_rec_rows(int R) { return (R + 1) | 1; }

struct At2TableLayout {
    int R, F, has_smooth, sg_stride;
    __host__ __device__ int rows() const { return at2_rec_rows(R); }
    __host__ __device__ int off_rec() const { return 0; }
    __host__ __device__ int off_sg() const { return AT2_REC * F * rows(); }    // 16-byte aligned
    __host__ __device__ int off_pw() const { return off_sg() + (has_smooth ? sg_stride * (R + 1) : 0); }
    __host__ __device__ int total_philox() const { return off_pw(); }          // what a Philox kernel stages
    __host__ __device__ int total() const { return off_pw() + AT2_PW * F * rows(); }
};

static inline int at2_check_gather(const at2_gather* g, const char* who) {
    AT2_REQUIRE(g != nullptr && g->n_targets >= 1 && g->n_targets <= AT2_MAX_PEERS, "%s: need 1..%d gather targets", who, AT2_MAX_PEERS);
    AT2_REQUIRE(g->base >= 0, "%s: gather base must be >= 0", who);
    AT2_REQUIRE(!g->is_multicast || g->n_targets == 1, "%s: a multicast gather has exactly one target address", who);
    for (int r = 0; r < g->n_targets; ++r) AT2_REQUIRE(g->d_target[r] != nullptr, "%s: gather target %d is NULL", who, r);
    return AT2_OK;
}

// Families that share (ml_agg, nec_agg) have identical per-step records (the other family attributes - up_spik, the shifts,
// is_smooth - are per-arm scalars, not table entries): they share one block of the table.  slot[f] = index of family f's
// block; returns the number of blocks.  (The reference's six general families are two copies of three shapes with
// different shifts, run_simulation.py:41-47: 3 blocks, 12 KB less shared memory per CTA.)
static inline int at2_family_slot_map(const at2_sim_config* cfg, int* slot) {
    int n = 0;
    for (int f = 0; f < cfg->n_families; ++f) {
        slot[f] = -1;
        for (int g = 0; g < f; ++g)
            if (cfg->fam[g].ml_agg == cfg->fam[f].ml_agg && cfg->fam[g].nec_agg == cfg->fam[f].nec_agg) {
                slot[f] = slot[g];
                break;
            }
        if (slot[f] < 0) slot[f] = n++;
    }
    return n;
}

static inline int at2_cfg_has_smooth(const at2_sim_config* cfg) {
    for (int f = 0; f < cfg->n_families; ++f)
        if (cfg->fam[f].is_smooth) return 1;
    return 0;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based -------------------------------------------------------------
// Round count: 10 is the standard Philox4x32-10 (Random123, cuRAND).  Salmon et al. report that 7 rounds already pass
// BigCrush; building with -DAT2_PHILOX_ROUNDS=7 is a measured what-if, not a supported configuration: the
// known-answer tests and every seeded expectation are for 10 rounds.
#ifndef AT2_PHILOX_ROUNDS
#define AT2_PHILOX_ROUNDS 10
#endif
#define AT2_PHILOX_M0 0xD2511F53u
#define AT2_PHILOX_M1 0xCD9E8D57u
#define AT2_PHILOX_W0 0x9E3779B9u
#define AT2_PHILOX_W1 0xBB67AE85u

struct At2U4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ At2U4 at2_philox10(At2U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < AT2_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)AT2_PHILOX_M0 * c.x;
        const uint64_t p1 = (uint64_t)AT2_PHILOX_M1 * c.z;
        At2U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += AT2_PHILOX_W0;
        k1 += AT2_PHILOX_W1;
    }
    return c;
}

// Same function with the key schedule precomputed on the host (rk[2 r], rk[2 r + 1] = keys of round r).  When rk points
// into a __grid_constant__ kernel parameter the round keys are constant-bank operands of the XORs: no key-bump
// instructions and no registers for the schedule.
__device__ __forceinline__ At2U4 at2_philox10_rk(At2U4 c, const uint32_t* __restrict__ rk) {
#pragma unroll
    for (int r = 0; r < AT2_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)AT2_PHILOX_M0 * c.x;
        const uint64_t p1 = (uint64_t)AT2_PHILOX_M1 * c.z;
        At2U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ rk[2 * r];
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ rk[2 * r + 1];
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}
static inline void at2_round_keys(uint64_t seed, uint32_t* rk) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk[2 * r] = k0, rk[2 * r + 1] = k1;
        k0 += AT2_PHILOX_W0, k1 += AT2_PHILOX_W1;
    }
}

// One 4-byte store to an NVSwitch multicast (multimem) address: the switch replicates it to every GPU of the group.  Only
// multimem.* instructions may touch such an address (PTX ISA 8.1+: any other access is undefined behaviour).
__device__ __forceinline__ void multimem_st_f32(float* mc_addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

__device__ __forceinline__ void multimem_st_f64(double* mc_addr, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc_addr), "d"(v) : "memory");
}

// RNG stream tags (4th counter word)
#define AT2_STREAM_ARM 0u   // arm initialisation: x, y, family, z
#define AT2_STREAM_TRAJ 1u  // trajectory attempts
#define AT2_STREAM_TPE 2u   // TPE candidate sampling
