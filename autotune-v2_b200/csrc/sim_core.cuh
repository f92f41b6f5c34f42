// Device-side core of the Gamma-process simulation, shared by the Hyperband kernel (hb_sim.cu) and the TPE / hybrid
// kernel (tpe_sim.cu).  Reference: core/simulation_evaluator.py:36-48,84-116 (Gamma draw + recurrence),
// benchmarks/opt_function_problem.py:49-138 (test functions), core/arm.py:44-55 + core/params.py:36 (arm draw),
// core/shape_family_scheduler.py:50-77 (family choice).
#pragma once
#include <climits>
#include <type_traits>

#include "at2_common.cuh"

namespace {

enum { MODE_PHILOX = 0, MODE_PREDRAWN = 1 };

// ---- arithmetic policies ---------------------------------------------------------------------------------------------
// double: the reference's float64 operation order with explicitly un-fused IEEE ops (bit-exact parity path).
// float : production arithmetic, contraction to FFMA welcome.
template <typename T>
struct Fam {
    T ml, up, s0, s1;  // float: ml = ml_agg/100 folded; double: raw ml_agg
    int smooth, pw_off;
};

__device__ __forceinline__ double step_exact(double f, double g, int t, int R, double fn, double ml_agg, double up_spik,
                                             double P, double P11) {
    if (g == 2.0) return f;                                                        // :100-101
    if (g > 2.0) {                                                                 // :102-107
        const double debt = __dsub_rn(fn, f);
        const double ml = __dadd_rn(f, __ddiv_rn(__dmul_rn(__dmul_rn(__dsub_rn(g, 2.0), ml_agg), debt), 100.0));
        return __dadd_rn(ml, __dmul_rn(__dsub_rn(fn, ml), P));
    }
    const double up = __dadd_rn(f, __ddiv_rn(__dmul_rn(up_spik, 1.0), __dadd_rn(1.0, g)));  // :108-115
    double nxt = __dadd_rn(up, __dmul_rn(__dsub_rn(fn, up), P11));
    if (R - t == 1) nxt = fn;
    return nxt;
}

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sin(float x) {
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_cos(float x) {
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float step_fast(float f, float g, float fn, float mlc, float m2c, float upc, float P, float P11,
                                           float Q) {
    // production arithmetic of simulation_evaluator.py:99-115.  Down branch (:103-107) in fused form:
    //   ml = f + a (fn - f), a = (g - 2) ml_agg / 100;  f' = ml + (fn - ml) P = f + (fn - f) (a (1 - P) + P),  Q = 1 - P.
    // At the last step P = P11 = 1, so both branches land on f_n to within one rounding (the reference's float64 path has
    // the same property on its down branch).
    const float a = fmaf(g, mlc, m2c);                      // m2c = -2 mlc, hoisted by the callers
    const float dn = fmaf(fn - f, fmaf(a, Q, P), f);
    const float up = fmaf(upc, fast_rcp(1.0f + g), f);
    const float un = fmaf(fn - up, P11, up);
    return g > 2.0f ? dn : (g < 2.0f ? un : f);
}

// ---- test functions in FP32 (benchmarks/opt_function_problem.py:49-138) ------------------------------------------------
__device__ __forceinline__ float egg_pair(float a, float b) {
    return -(b + 47.0f) * sinf(sqrtf(fabsf(b + 0.5f * a + 47.0f))) - a * sinf(sqrtf(fabsf(a - b - 47.0f)));
}
// (x2, y2): third and fourth coordinate, used by the 4-D extension only
__device__ __noinline__ float base_function(int func, float x, float y, float x2 = 0.0f, float y2 = 0.0f) {
    switch (func) {
        case AT2_FUNC_EGG4:   // extension (include/at2.h): n-dimensional Egg-holder over consecutive coordinate pairs
            return (egg_pair(x, y) + egg_pair(y, x2)) + egg_pair(x2, y2);
        case AT2_FUNC_EGG:
            return -(y + 47.0f) * sinf(sqrtf(fabsf(y + 0.5f * x + 47.0f))) - x * sinf(sqrtf(fabsf(x - y - 47.0f)));
        case AT2_FUNC_BRANIN: {
            const float b = 5.1f / (4.0f * 9.8696044010893586f), c = 5.0f / 3.14159265358979f;
            const float tt = 1.0f / (8.0f * 3.14159265358979f);
            const float q = y - b * x * x + c * x - 6.0f;
            return q * q + 10.0f * (1.0f - tt) * cosf(x) + 10.0f;
        }
        case AT2_FUNC_CAMEL: {
            const float x2 = x * x, y2 = y * y;
            return (4.0f - 2.1f * x2 + x2 * x2 / 3.0f) * x2 + x * y + (-4.0f + 4.0f * y2) * y2;
        }
        case AT2_FUNC_WAVE: {
            const float r2 = x * x + y * y;
            return -(1.0f + cosf(12.0f * sqrtf(r2))) / (0.5f * r2 + 2.0f);
        }
        default:  // AT2_FUNC_RASTRIGIN
            return 20.0f + ((x * x - 10.0f * cospif(2.0f * x)) + (y * y - 10.0f * cospif(2.0f * y)));
    }
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// Table fetch of the trajectory loop.  GTAB = false: the tables were staged into shared memory and `addr` is a shared-space
// address.  GTAB = true (plans whose tables do not fit beside the per-warp state: R ~ 1000 with six families): `addr` is a
// byte offset into the table in global memory, read through the read-only path (L1/L2 resident: <= 300 KB per launch).
template <bool GTAB>
__device__ __forceinline__ float4 tab128(const unsigned char* gbase, uint32_t addr) {
    if constexpr (GTAB)
        return __ldg(reinterpret_cast<const float4*>(gbase + addr));
    else
        return lds128(addr);
}
// st.shared under a predicate (ptxas turns an `if` around a store into a branch)
__device__ __forceinline__ void sts32_if(uint32_t addr, float v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}" ::"r"(addr), "f"(v), "r"((int)p) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <bool GTAB>
__device__ __forceinline__ float tab32(const unsigned char* gbase, uint32_t addr) {
    if constexpr (GTAB)
        return __ldg(reinterpret_cast<const float*>(gbase + addr));
    else
        return lds32(addr);
}
__device__ __forceinline__ void team_barrier(int team, int team_warps) {
    if (team_warps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_warps * 32) : "memory");
}


// Where arms live and how they are drawn (Philox path).
struct ArmSpace {
    int func, F, scheduler;
    int n_slots;       // blocks of per-step records in the table (families with equal shapes share one, at2_family_slot_map)
    float x_min, x_rng, y_min, y_rng, noise;
    uint32_t key0, key1;
    uint32_t rk[20];   // Philox key schedule of (key0, key1), at2_round_keys
};

// arm draw of the Philox path: x, y uniform on the domain (core/params.py:36), family
// (core/shape_family_scheduler.py:61,75), z ~ N(0,1) for the initial noise (benchmarks/opt_function_problem.py:138)
struct ArmDraw {
    float x, y, z;
    int fam;
    float x2, y2;   // AT2_FUNC_EGG4 only
};
__device__ __noinline__ ArmDraw draw_arm(const ArmSpace& sp, unsigned long long gid, int arm) {
    const At2U4 r = at2_philox10(At2U4{0u, (uint32_t)gid, (uint32_t)(gid >> 32), AT2_STREAM_ARM}, sp.key0, sp.key1);
    ArmDraw d;
    d.x = fmaf((float)(r.x >> 8) * 5.9604644775390625e-8f, sp.x_rng, sp.x_min);
    d.y = fmaf((float)(r.y >> 8) * 5.9604644775390625e-8f, sp.y_rng, sp.y_min);
    d.fam = sp.scheduler == AT2_SCHED_UNIFORM ? (int)__umulhi(r.z, (uint32_t)sp.F) : arm % sp.F;
    const float u1 = fmaf((float)(r.w >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);
    d.z = d.x2 = d.y2 = 0.0f;
    if (sp.noise != 0.0f) {
        // Box-Muller with the approximate MUFU forms (relative error ~1e-6, far below the FP32 tolerance tier; the full and
        // the early-stopping kernels share this function, so they stay bit-identical to each other).  The angle's leading
        // 24 bits are the low bytes of the x, u1 and y words - bits their own conversions (>> 8) never look at - so a noisy
        // 2-D problem draws an arm from ONE Philox block.  (The fourth byte repeats x's: 2^-24 of a turn, immaterial.)
        const uint32_t ang = __byte_perm(__byte_perm(r.x, r.y, 0x0040), r.w, 0x0410);   // bytes (x.0, w.0, y.0, x.0), MSB first
        d.z = fast_sqrt(-1.3862943611198906f * fast_lg2(u1)) * fast_cos((float)ang * 1.4629180792671596e-9f);   // 2 pi / 2^32
    }
    // the second block feeds the extra coordinates of the 4-D extension only
    if (sp.func == AT2_FUNC_EGG4) {
        const At2U4 r2 = at2_philox10(At2U4{1u, (uint32_t)gid, (uint32_t)(gid >> 32), AT2_STREAM_ARM}, sp.key0, sp.key1);
        d.x2 = fmaf((float)(r2.y >> 8) * 5.9604644775390625e-8f, sp.x_rng, sp.x_min);
        d.y2 = fmaf((float)(r2.z >> 8) * 5.9604644775390625e-8f, sp.y_rng, sp.y_min);
    }
    return d;
}

// the family draw_arm would give this arm, without the rest of the draw
__device__ __noinline__ int arm_family(const ArmSpace& sp, unsigned long long gid, int arm) {
    if (sp.scheduler != AT2_SCHED_UNIFORM) return arm % sp.F;
    const At2U4 r = at2_philox10(At2U4{0u, (uint32_t)gid, (uint32_t)(gid >> 32), AT2_STREAM_ARM}, sp.key0, sp.key1);
    return (int)__umulhi(r.z, (uint32_t)sp.F);
}

// Shared-memory tables of one CTA as the trajectory loop sees them.
struct SimTables {
    uint32_t rec_addr;     // shared-space address of rec[0]  (GTAB: byte offset of rec in the global table)
    uint32_t sg_addr;      // shared-space address of sg[0][0]   (GTAB: byte offset of sg)
    const int* ckpos;      // checkpoint positions (ckpt_time - 1), INT_MAX-terminated, in shared memory
    int R, sg_stride;
    const unsigned char* gbase = nullptr;   // GTAB only: the table in global memory
    int lean_from = INT_MAX;   // first step (1-based record row) from which no checkpoint but the end value follows, see
                               // philox_trajectory SPLIT; INT_MAX: the plan's last checkpoint is not the last position
};

// ---- the Philox word stream of a trajectory ---------------------------------------------------------------------------
// Attempt pair j of a lane (attempts 2j, 2j+1) owns words 2j, 2j+1 of the lane's flat Philox4x32-10 word stream (word i =
// component i % 4 of the block with counter i / 4, stream tag AT2_STREAM_TRAJ) - ONE word per attempt:
//   Box-Muller: radius from the first word, angle from the second -> two normals.  The conversions keep the words' top 24
//               bits (I2F rounds the rest away);
//   accept uniforms, 32 bits each: top byte = the LOW byte of the pair's own word (wa for attempt 2j, wb for 2j+1) - eight
//               fresh bits the normals never see - and below it the partner word's bytes 1, 2, 3 in that order, i.e. its
//               least significant used byte on top.  Given the normal, those 24 bits act as a dither: byte 1 of a word moves
//               the radius / angle by 2^-16 relative, so it is uniform and independent of the acceptance threshold to
//               O(1e-4 / 256), and a dither that is uniform on [0, 1) makes the 8-bit comparison unbiased:
//               P(F + delta < 256 A) = A exactly.  Checked at 10^6 draws per shape against the Gamma law
//               (tests/test_parity_tiers_gpu.py: KS D = 6e-4).
// Two blocks (A, B) therefore feed four pairs = eight attempts:
//   pair 0: A.x A.y     pair 1: A.z A.w     pair 2: B.x B.y     pair 3: B.z B.w
// The mapping attempt -> words is a pure function of the attempt index: a partial run (early stopping, affine offsets,
// full-trajectory output) walks exactly the prefix of the stream the full run walks.
__device__ __forceinline__ void box_muller_pair(uint32_t wa, uint32_t wb, float& x0, float& x1, float& ku0, float& ku1) {
    const float u1 = fmaf((float)wa, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (0,1]
    const float rad = fast_sqrt(-1.3862943611198906f * fast_lg2(u1));                      // sqrt(-2 ln u1)
    const float ang = (float)(int32_t)wb * 1.4629180792671596e-9f;                        // [-pi, pi)
    x0 = rad * fast_cos(ang);
    x1 = rad * fast_sin(ang);
    ku0 = (float)__byte_perm(wa, wb, 0x0567);    // bytes (wa.0, wb.1, wb.2, wb.3), most significant first: uniform x 2^32
    ku1 = (float)__byte_perm(wa, wb, 0x4123);    // bytes (wb.0, wa.1, wa.2, wa.3)
}

// One Marsaglia-Tsang attempt at the step whose record sits at (gm, pw) - Marsaglia & Tsang 2000, full log test
//   log2 U < log2e (x^2/2 + d - d v^3) + d log2 v^3,  with log2 U = log2(ku) - 32 folded into the record (gm.w = d log2e + 32)
// - and the step of the recurrence it feeds (simulation_evaluator.py:99-115) in folded form:
//   g = w v^3 (Gamma draw), g1 = 1 + g
//   g > 2:  f' = k fn + (1 - k) f,          k = a (1 - P) + P, a = (g - 2) ml_agg / 100  ==  v^3 K1w + K0   (:103-107)
//   g <= 2: up = f + up_spik / (1 + g),  f' = P11 fn + (1 - P11) up                                           (:108-115)
// K1w = w ml (1 - P) and K0 = P - 2 ml (1 - P) are per-(family, step) constants of the table.  At the last step
// P = P11 = 1, so both branches land on fn (:113-115).  g == 2 exactly (:100-101, f unchanged) has probability ~2^-24 per
// draw with float g and is not singled out here; the pre-drawn paths, which are compared with the reference, keep it.
// v <= 0 needs no test of its own: lg2 of a non-positive v^3 is NaN / -inf and the comparison fails; the sentinel record
// of a finished lane has c = NaN and fails the same way.
struct AttemptOut {
    bool ok;
    float fnew;
};
__device__ __forceinline__ AttemptOut mt_attempt(const float4 gm, const float4 pw, float x, float ku, float ff, float fn,
                                                 float upc) {
    const float v = fmaf(gm.y, x, 1.0f);
    const float v3 = v * v * v;
    const float rhs = fmaf(gm.x, fast_lg2(v3), fmaf(-gm.z, v3, fmaf(0.72134752044448170f, x * x, gm.w)));
    AttemptOut o;
    o.ok = fast_lg2(ku) < rhs;
    const float g1 = fmaf(fabsf(pw.x), v3, 1.0f);
    const bool gt = g1 > 3.0f;
    const float kdn = fmaf(v3, pw.y, pw.z);
    const float base = fmaf(gt ? 0.0f : upc, fast_rcp(g1), ff);
    const float k = gt ? kdn : pw.w;
    // k fn + (1 - k) base in the form that lands EXACTLY on fn when k == 1 (base - base == 0): the last step of every
    // trajectory (K1w = 0, K0 = P11 = 1 there), so that a raw arm read at R shows fn itself - which lets the early-stopping
    // kernel skip that run altogether (hb_lazy.cu) and stay bit-identical
    o.fnew = fmaf(k, fn, fmaf(-k, base, base));
    return o;
}

// the same step on a given Gamma draw g (pre-drawn float path, full-precision inputs): K1 = ml (1 - P), K0, P11
__device__ __forceinline__ float step_folded(float f, float g, float fn, float upc, float K1, float K0, float P11) {
    if (g == 2.0f) return f;                                                        // :100-101
    const bool gt = g > 2.0f;
    const float base = gt ? f : fmaf(upc, fast_rcp(1.0f + g), f);
    const float k = gt ? fmaf(g, K1, K0) : P11;
    return fmaf(k, fn, fmaf(-k, base, base));
}

// One arm per lane: runs the lane's Philox4x32-10 stream -> Box-Muller -> Marsaglia-Tsang -> recurrence until every
// lane of the warp has produced position stop_t - 1 (stop_t = R: the whole trajectory), writing the raw arms'
// checkpoint losses to ck[c * stride] as their positions are produced and the smooth arms' Savitzky-Golay read-outs at
// the end.  All 32 lanes must call it; lanes with active == false do nothing.
//
// A lane's epoch is the shared-memory address of its (family, step) row: an accepted attempt advances it by one
// row, a rejected one leaves it, a finished (or padding) lane sits on the sentinel record whose NaN slope fails every
// acceptance test - so the attempt is branch-free and needs no bounds check.  Checkpoints are flagged in the records
// themselves (sign bit of w, set by at2_hb_tables_build for the plan's checkpoint positions): an accepted attempt at a
// flagged step stores the new loss through the lane's row pointer and moves that pointer one row down - two predicated
// instructions, no branch.  Every active lane needs at least R - 1 attempts, so the first (R - 1) / 4 trips (one Philox
// block, four attempts each) run with no vote and no divergence - the integer Philox rounds of the next block overlap the
// MUFU/FP32 chains of the attempts; only the tail, where lanes finish one after the other, tests for completion.
//
// PARTIAL = true (early-stopping evaluation): every lane has its own stop_t and stops advancing there, and only checkpoint
// `only_c` is produced, into ck[0] (instantiate with CK = 1: a smoothed arm then accumulates that one Savitzky-Golay
// read-out - one scalar table fetch and one FMA per step instead of all C of them).  The stream and the arithmetic are
// the same as in the full run, so the value equals the one the full run reads at that checkpoint, bit for bit.
//
// SPLIT = true (full runs): the loop exists twice.  Hyperband's checkpoints are geometric (R eta^-k), so all but the last
// sit in the first 1/eta of the trajectory and the last is the end value: once every lane of the warp has passed the
// second-to-last checkpoint (tb.lean_from) the rest of the run goes through a copy of the loop without the capture
// instructions (4 of ~49 per attempt), and the end value is stored after the loop.  Plans whose last checkpoint is not the
// last position keep the capturing loop throughout (lean_from = INT_MAX).
template <typename T, bool HAS_SMOOTH, int CK, bool PARTIAL = false, bool GTAB = false, bool SPLIT = false>
__device__ __forceinline__ void philox_trajectory(const SimTables& tb, bool active, bool smooth, float f1, float fnf,
                                                  float upc, int pw_off, unsigned long long gid,
                                                  const uint32_t* __restrict__ rk, int stop_t, T* ck, int stride, int C,
                                                  int only_c = 0) {
    const int R = tb.R;
    const unsigned char* const gbase = tb.gbase;
    uint32_t rec_base = tb.rec_addr + (uint32_t)pw_off * (AT2_REC * 4u);
    static_assert(!PARTIAL || CK == 1, "a partial run produces one checkpoint");
    uint32_t sg_base = tb.sg_addr + (PARTIAL ? (uint32_t)only_c * 4u : 0u);   // PARTIAL: column only_c of sg
    // the Savitzky-Golay row of the step a lane stands on: sg rows have the record stride (x2 beyond 12 checkpoints), so it
    // sits at a fixed distance from the lane's record address
    const uint32_t sg_mul = (uint32_t)tb.sg_stride / AT2_REC;
    const uint32_t sg_delta = sg_base - rec_base * sg_mul;
    // a partial run (stop_t < R) of a finished lane must not walk on: park inactive lanes on the sentinel record
    const uint32_t end_addr = rec_base + (uint32_t)stop_t * (AT2_REC * 4u);
    uint32_t addr = rec_base + (uint32_t)(active ? 1 : R) * (AT2_REC * 4u);   // record of the step to produce next
    float ff = f1;
    static_assert(sizeof(T) == 4, "the Philox paths are float");
    // row of the next checkpoint to capture, as a shared-space address (ck always lives in shared memory)
    uint32_t ckp = (uint32_t)__cvta_generic_to_shared(ck);
    const uint32_t row_bytes = (uint32_t)stride * 4u;
    float acc[CK];
#pragma unroll
    for (int c = 0; c < CK; ++c) acc[c] = 0.0f;
    if constexpr (HAS_SMOOTH && PARTIAL) {
        acc[0] = tab32<GTAB>(gbase, sg_base) * ((active && smooth) ? ff : 0.0f);
    } else if constexpr (HAS_SMOOTH) {
        const float f0 = (active && smooth) ? ff : 0.0f;
#pragma unroll
        for (int c4 = 0; c4 < CK; c4 += 4) {                                  // sg[0][c] * fs[0]
            if (CK - c4 == 1) {
                acc[c4] = tab32<GTAB>(gbase, sg_base + c4 * 4u) * f0;
            } else {
                const float4 w = tab128<GTAB>(gbase, sg_base + c4 * 4u);
                acc[c4] = w.x * f0;
                if (c4 + 1 < CK) acc[c4 + 1] = w.y * f0;
                if (c4 + 2 < CK) acc[c4 + 2] = w.z * f0;
                if (c4 + 3 < CK) acc[c4 + 3] = w.w * f0;
            }
        }
    }
    if constexpr (!PARTIAL) {
        if (active && tb.ckpos[0] == 0) {          // position 0 is the start value
            sts32_if(ckp, ff, true);
            ckp += row_bytes;
        }
    }
    const uint32_t g_lo = (uint32_t)gid, g_hi = (uint32_t)(gid >> 32);
    auto attempt = [&](float x, float ku, auto cap_tag) {
        constexpr bool CAPTURE = decltype(cap_tag)::value;
        const float4 gm = tab128<GTAB>(gbase, addr);                               // d, c, d log2e, d log2e + 32
        const float4 pw = tab128<GTAB>(gbase, addr + 16u);                         // +-w, K1w, K0, P11
        const AttemptOut o = mt_attempt(gm, pw, x, ku, ff, fnf, upc);
        bool ok = o.ok;
        if constexpr (PARTIAL) ok = ok && addr < end_addr;
        ff = ok ? o.fnew : ff;
        if constexpr (CAPTURE) {
            // smooth lanes store too (into rows the read-outs below overwrite): no test of their own
            const bool cap = ok && __float_as_int(pw.x) < 0;
            sts32_if(ckp, ff, cap);                  // predicated store: no branch in the attempt
            ckp += cap ? row_bytes : 0u;
        }
        if constexpr (HAS_SMOOTH && PARTIAL) {
            const float fs = (ok && smooth) ? ff : 0.0f;
            acc[0] = fmaf(tab32<GTAB>(gbase, addr * sg_mul + sg_delta), fs, acc[0]);
        } else if constexpr (HAS_SMOOTH) {
            const float fs = (ok && smooth) ? ff : 0.0f;
            const uint32_t a = addr * sg_mul + sg_delta;
            // The smoothed passes are bound by shared-memory wavefronts (a 128-bit load is four of them whatever the lanes'
            // addresses; a 32-bit load of lanes standing on up to eight consecutive steps is one: the 48-byte row stride puts
            // them in distinct banks) - so a lone last weight (CK = 5: R = 81, eta = 3) is fetched as a scalar.
#pragma unroll
            for (int c4 = 0; c4 < CK; c4 += 4) {
                if (CK - c4 == 1) {
                    acc[c4] = fmaf(tab32<GTAB>(gbase, a + c4 * 4u), fs, acc[c4]);
                } else {
                    const float4 w = tab128<GTAB>(gbase, a + c4 * 4u);
                    acc[c4] = fmaf(w.x, fs, acc[c4]);
                    if (c4 + 1 < CK) acc[c4 + 1] = fmaf(w.y, fs, acc[c4 + 1]);
                    if (c4 + 2 < CK) acc[c4 + 2] = fmaf(w.z, fs, acc[c4 + 2]);
                    if (c4 + 3 < CK) acc[c4 + 3] = fmaf(w.w, fs, acc[c4 + 3]);
                }
            }
        }
        if (ok) addr += AT2_REC * 4u;
    };
    auto pair = [&](uint32_t wa, uint32_t wb, auto cap_tag) {
        float x0, x1, ku0, ku1;
        box_muller_pair(wa, wb, x0, x1, ku0, ku1);
        attempt(x0, ku0, cap_tag);
        attempt(x1, ku1, cap_tag);
    };
    auto block = [&](uint32_t ctr) { return at2_philox10_rk(At2U4{ctr, g_lo, g_hi, AT2_STREAM_TRAJ}, rk); };
    auto unfinished = [&]() { return __any_sync(0xffffffffu, addr < end_addr); };
    // a partial run that reads the first point (r = 1) has no step to make: not even the first Philox block
    if (!PARTIAL || unfinished()) {
        // every unfinished lane needs at least this many more attempts
        const unsigned need = __reduce_min_sync(0xffffffffu, active ? (unsigned)(stop_t - 1) : 0xffffffffu);
        // counted trips: eight attempts each in a partial run, four in a full one
        const int n_main = need == 0xffffffffu ? 0 : (int)(need >> (PARTIAL ? 3 : 2));
        uint32_t ctr = 0;
        At2U4 A = block(0), B;
        // ONE loop instance serves the counted trips and the tail (the completion vote only runs once the counted part is
        // over): the kernels that carry a lean and a smoothed trajectory loop were instruction-fetch bound with a separate
        // tail loop each (ncu, six families: no_instruction 1.1 cycles per instruction; merged: 3.13 -> 2.89 ms).
        // A full run's trip is ONE Philox block = four attempts, with the next block's integer rounds issued ahead of them
        // (they overlap the attempts' MUFU chains): half the code of an eight-attempt trip - which is what lets every loop
        // variant carry its capture-free copy - and a tail that overshoots the slowest lane by two attempts on average
        // instead of four (measured against the eight-attempt trip: -1.7 % flat, -2 % with smoothed families).
        // A partial run - often only a few steps long - generates its blocks only once it knows it needs them.
        // run(cap_tag, n_counted, until): trips while fewer than n_counted have run or some lane is below `until`
        int it = 0;
        auto run = [&](auto cap_tag, int n_counted, uint32_t until) {
#pragma unroll 1
            for (;; ++it) {
                if (it >= n_counted && !__any_sync(0xffffffffu, addr < until)) break;
                if constexpr (PARTIAL) {
                    pair(A.x, A.y, cap_tag);
                    pair(A.z, A.w, cap_tag);
                    if (it >= n_counted && !__any_sync(0xffffffffu, addr < until)) break;
                    B = block(ctr + 1);
                    const At2U4 An = block(ctr + 2);
                    ctr += 2;
                    pair(B.x, B.y, cap_tag);
                    pair(B.z, B.w, cap_tag);
                    A = An;
                } else {
                    const At2U4 An = block(ctr + 1);
                    ctr += 1;
                    pair(A.x, A.y, cap_tag);
                    pair(A.z, A.w, cap_tag);
                    A = An;
                }
            }
        };
        if constexpr (PARTIAL) {
            run(std::false_type{}, n_main, end_addr);
        } else if constexpr (!SPLIT) {
            run(std::true_type{}, n_main, end_addr);
        } else {
            // capturing trips until every lane stands at or beyond step lean_from (inactive lanes sit on the sentinel,
            // beyond any step), then the capture-free copy to the end; the end value is the last checkpoint
            const bool split = tb.lean_from < R;
            const uint32_t lean_addr = rec_base + (uint32_t)(split ? tb.lean_from : R) * (AT2_REC * 4u);
            run(std::true_type{}, split ? ((tb.lean_from - 1) >> 2) : n_main, split ? lean_addr : end_addr);
            if (split) {
                run(std::false_type{}, n_main, end_addr);
                // to the last row by its own address, not through ckp: on a short plan a lane may already have captured the
                // end value in the capturing copy (same value - the store is idempotent - but ckp has moved past the row)
                sts32_if((uint32_t)__cvta_generic_to_shared(ck) + (uint32_t)(C - 1) * row_bytes, ff, active);
            }
        }
    }
    if constexpr (PARTIAL) {
        // the run stopped where the read-out is: a raw arm's value is the loss it stands on
        if (active) ck[0] = (T)((HAS_SMOOTH && smooth) ? acc[0] : ff);
    } else if constexpr (HAS_SMOOTH) {
        if (active && smooth) {
#pragma unroll
            for (int c = 0; c < CK; ++c)
                if (c < C) ck[c * stride] = (T)acc[c];
        }
    }
}

// The recurrence is affine in its two inputs: f_t = A_t f_1 + (1 - A_t) f_n + C_t, where (A_t, C_t) depend only on the arm's
// Gamma stream and family (down step: f' = (1-k) f + k f_n; up step: f' = (1-P11)(f + up/(1+g)) + P11 f_n) - and so is the
// Savitzky-Golay read-out (its rows sum to one).  Since f_1 = f + (noise z - start_shift) and f_n = f - end_shift, the loss
// an arm shows at a checkpoint is  f(x, y) + offset  with an offset that is known BEFORE the arm's hyperparameters are.
// This walks the lane's stream (same Philox words, same acceptance arithmetic as philox_trajectory, so the same Gamma
// draws) up to the read-out of checkpoint c and returns that offset.  The TPE kernel computes it for all arms of a bracket
// in parallel (one arm per lane) before its sequential suggest -> observe loop, instead of re-walking each arm's stream
// with 32 redundant lanes once its hyperparameters have been suggested.  All 32 lanes must call it.
//   dz = noise z - start_shift (f_1 - f), s1 = end_shift (f - f_n); stop_t: 1 + last position the read-out needs.
template <bool HAS_SMOOTH>
__device__ __forceinline__ float philox_affine_offset(const SimTables& tb, bool active, bool smooth, float upc,
                                                      int pw_off, unsigned long long gid, const uint32_t* __restrict__ rk,
                                                      int stop_t, int c, float dz, float s1) {
    const uint32_t rec_base = tb.rec_addr + (uint32_t)pw_off * (AT2_REC * 4u);
    const uint32_t end_addr = rec_base + (uint32_t)stop_t * (AT2_REC * 4u);
    uint32_t addr = active ? rec_base + AT2_REC * 4u : end_addr;         // record of the step to produce next
    const uint32_t sg_col = tb.sg_addr + (uint32_t)c * 4u;
    const uint32_t sg_mul = (uint32_t)tb.sg_stride / AT2_REC;
    float A = 1.0f, Cc = 0.0f, accA = 0.0f, accC = 0.0f;
    if constexpr (HAS_SMOOTH) accA = smooth ? lds32(sg_col) : 0.0f;      // sg[0][c] * (A_0 = 1)
    const uint32_t g_lo = (uint32_t)gid, g_hi = (uint32_t)(gid >> 32);
    auto attempt = [&](float x, float ku) {
        const float4 gm = lds128(addr);
        const float4 pw = lds128(addr + 16u);
        const float v = fmaf(gm.y, x, 1.0f);
        const float v3 = v * v * v;
        const float rhs = fmaf(gm.x, fast_lg2(v3), fmaf(-gm.z, v3, fmaf(0.72134752044448170f, x * x, gm.w)));
        const bool ok = fast_lg2(ku) < rhs && addr < end_addr;
        const float g1 = fmaf(fabsf(pw.x), v3, 1.0f);
        const bool dn = g1 > 3.0f;
        const float k = fmaf(v3, pw.y, pw.z);
        const float q = 1.0f - pw.w;
        const float An = dn ? fmaf(-k, A, A) : q * A;
        const float Cn = dn ? fmaf(-k, Cc, Cc) : q * fmaf(upc, fast_rcp(g1), Cc);
        A = ok ? An : A, Cc = ok ? Cn : Cc;
        if constexpr (HAS_SMOOTH) {
            const float w = (ok && smooth) ? lds32(sg_col + (addr - rec_base) * sg_mul) : 0.0f;
            accA = fmaf(w, A, accA), accC = fmaf(w, Cc, accC);
        }
        addr += ok ? (AT2_REC * 4u) : 0u;
    };
    auto pair = [&](uint32_t wa, uint32_t wb) {
        float x0, x1, ku0, ku1;
        box_muller_pair(wa, wb, x0, x1, ku0, ku1);
        attempt(x0, ku0);
        attempt(x1, ku1);
    };
    auto block = [&](uint32_t ctr) { return at2_philox10_rk(At2U4{ctr, g_lo, g_hi, AT2_STREAM_TRAJ}, rk); };
    auto unfinished = [&]() { return __any_sync(0xffffffffu, addr < end_addr); };
    uint32_t ctr = 0;
#pragma unroll 1
    while (unfinished()) {
        const At2U4 A4 = block(ctr);
        ++ctr;
        pair(A4.x, A4.y);
        pair(A4.z, A4.w);
    }
    if (HAS_SMOOTH && smooth) A = accA, Cc = accC;
    return fmaf(A, dz, fmaf(A, s1, -s1)) + Cc;                           // A dz - (1 - A) s1 + C
}

}  // namespace
