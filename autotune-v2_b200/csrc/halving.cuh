// (b) Successive halving of one repetition inside one warp - shared by the simulation and the closest-known-function
// kernels.  Reference: optimisers/sequential/hyperband_optimiser.py:46-57 (stable sorted() top-k), :89-100 (rounds,
// best-in-round), core/optimiser.py:67-68 (OFE = first optimum over the round-bests).
#pragma once
#include "at2_common.cuh"

namespace {

// ---- sort keys: (loss, position) with position breaking ties = python's stable sorted() ---------------------------------
// float : 64-bit key  [orderable(loss) : 32 | position : 32]
// double: 96-bit key  {orderable(loss) : 64, position : 32}
__device__ __forceinline__ uint32_t orderable(float x, bool desc) {
    x += 0.0f;  // -0.0 -> +0.0 (python compares them equal)
    const uint32_t b = __float_as_uint(x);
    const uint32_t o = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    return desc ? ~o : o;
}
__device__ __forceinline__ uint64_t orderable(double x, bool desc) {
    x += 0.0;
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    const uint64_t o = b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    return desc ? ~o : o;
}

struct Key32 {  // float losses
    uint64_t v;
    __device__ __forceinline__ static Key32 make(float loss, int pos, bool desc) {
        return Key32{((uint64_t)orderable(loss, desc) << 32) | (uint32_t)pos};
    }
    __device__ __forceinline__ static Key32 pad() { return Key32{~0ull}; }
    __device__ __forceinline__ int pos() const { return (int)(uint32_t)v; }
    __device__ __forceinline__ bool operator>(const Key32& o) const { return v > o.v; }
    __device__ __forceinline__ Key32 shfl_xor(int m) const { return Key32{__shfl_xor_sync(0xffffffffu, v, m)}; }
    __device__ __forceinline__ Key32 shfl(int l) const { return Key32{__shfl_sync(0xffffffffu, v, l)}; }
};
struct Key64 {  // double losses
    uint64_t k;
    uint32_t p;
    uint32_t _pad;
    __device__ __forceinline__ static Key64 make(double loss, int pos, bool desc) {
        return Key64{orderable(loss, desc), (uint32_t)pos, 0u};
    }
    __device__ __forceinline__ static Key64 pad() { return Key64{~0ull, ~0u, 0u}; }
    __device__ __forceinline__ int pos() const { return (int)p; }
    __device__ __forceinline__ bool operator>(const Key64& o) const { return k > o.k || (k == o.k && p > o.p); }
    __device__ __forceinline__ Key64 shfl_xor(int m) const {
        return Key64{__shfl_xor_sync(0xffffffffu, k, m), __shfl_xor_sync(0xffffffffu, p, m), 0u};
    }
    __device__ __forceinline__ Key64 shfl(int l) const {
        return Key64{__shfl_sync(0xffffffffu, k, l), __shfl_sync(0xffffffffu, p, l), 0u};
    }
};
template <typename T>
struct KeyOf;
template <>
struct KeyOf<float> {
    typedef Key32 type;
};
template <>
struct KeyOf<double> {
    typedef Key64 type;
};

// ascending bitonic network over `npad` (power of two, > 32) keys in shared memory, executed by one warp
template <typename K>
__device__ void warp_bitonic_smem(K* keys, int npad, int lane) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int idx = lane; idx < (npad >> 1); idx += 32) {
                const int a = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));
                const int b = a | j;
                const bool up = (a & k) == 0;
                const K ka = keys[a], kb = keys[b];
                if ((ka > kb) == up) {
                    keys[a] = kb;
                    keys[b] = ka;
                }
            }
            __syncwarp();
        }
    }
}

// ascending bitonic network over the first `n` (<= 32) keys, one per lane, warp shuffles only; lanes >= n hold pads.
// Only the stages of the smallest power-of-two network covering n are executed (n is warp-uniform).
template <typename K>
__device__ __forceinline__ K warp_bitonic_shfl(K mine, int lane, int n) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        if ((k >> 1) >= n) break;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const K other = mine.shfl_xor(j);
            const bool up = (lane & k) == 0;
            const bool lower = (lane & j) == 0;
            // the lower lane of a pair keeps the smaller key in an ascending block, the larger in a descending one
            const bool take_other = (mine > other) == (up == lower);
            if (take_other) mine = other;
        }
    }
    return mine;
}

// ---- float losses: register-resident rounds ---------------------------------------------------------------------------
// n <= 32 keys, one per lane: stable ascending rank of the lane's key by counting.  Equal keys (rare) are ranked by lane =
// by position, python's stable sorted(): MATCH.ANY finds the lanes holding the same key, the ones below add to the rank.
__device__ __forceinline__ int warp_rank_by_counting(uint32_t ord, int lane, int n) {
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) rank += __shfl_sync(0xffffffffu, ord, j) < ord ? 1 : 0;
    return rank + __popc(__match_any_sync(0xffffffffu, ord) & ((1u << lane) - 1u));
}

// ascending bitonic sort of KPL * 32 64-bit keys held KPL per lane (key i lives in lane i / KPL, register i % KPL):
// compare-exchanges at distance < KPL stay inside a lane, the others are one 64-bit shuffle per register - no shared
// memory, no barriers, KPL independent chains per stage.
template <int KPL>
__device__ __forceinline__ void warp_bitonic_regs(uint64_t (&k)[KPL], int lane) {
    // sizes 2 .. KPL: inside a lane.  Block direction alternates with bit `size` of the index; for size == KPL that bit
    // is the lowest lane bit.
#pragma unroll
    for (int size = 2; size <= KPL; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                if ((r & j) == 0) {
                    const bool up = size < KPL ? (r & size) == 0 : ((lane * KPL) & size) == 0;
                    const uint64_t a = k[r], b = k[r | j];
                    const bool sw = (a > b) == up;
                    k[r] = sw ? b : a, k[r | j] = sw ? a : b;
                }
            }
        }
    }
#pragma unroll 1
    for (int size = 2 * KPL; size <= 32 * KPL; size <<= 1) {
        const bool up = ((lane * KPL) & size) == 0;            // ascending block (always, for the last size)
#pragma unroll 1
        for (int lm = size / (2 * KPL); lm > 0; lm >>= 1) {     // partner lane = lane ^ lm, same register
            const bool lower = (lane & lm) == 0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, k[r], lm);
                k[r] = ((k[r] > o) == (up == lower)) ? o : k[r];
            }
        }
#pragma unroll
        for (int j = KPL >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                if ((r & j) == 0) {
                    const uint64_t a = k[r], b = k[r | j];
                    const bool sw = (a > b) == up;
                    k[r] = sw ? b : a, k[r | j] = sw ? a : b;
                }
            }
        }
    }
}

// One successive-halving round of one repetition on float losses (hyperband_optimiser.py:46-57,93-100): ranks the n_cur
// alive arms - arm id ids[j], or first + j in a bracket's first round - by loss_at(j, id) with python's stable order,
// writes the ids of the `keep` best, in sorted order, to ids_next and returns the round's best (loss, arm id).
// keys: npad sort keys of warp-private scratch, used only by rounds wider than 256.  All 32 lanes must call it.
template <typename LossAt>
__device__ __forceinline__ void warp_halving_round_f32(int n_cur, int keep, bool first_round, int first,
                                                       const unsigned short* ids, unsigned short* ids_next, Key32* keys,
                                                       bool desc, int lane, LossAt loss_at, float& rb_loss, int& rb_arm) {
    if (n_cur <= 32) {
        int my_id = 0;
        float my_loss = 0.0f;
        uint32_t ord = 0xffffffffu;
        if (lane < n_cur) {
            my_id = first_round ? first + lane : (int)ids[lane];
            my_loss = loss_at(lane, my_id);
            ord = orderable(my_loss, desc);
        }
        const int rank = warp_rank_by_counting(ord, lane, n_cur);
        if (lane < n_cur && rank < keep) ids_next[rank] = (unsigned short)my_id;
        const int src = __ffs(__ballot_sync(0xffffffffu, lane < n_cur && rank == 0)) - 1;
        rb_loss = __shfl_sync(0xffffffffu, my_loss, src);
        rb_arm = __shfl_sync(0xffffffffu, my_id, src);
    } else if (n_cur <= 256) {
        if (n_cur <= 128) {
            uint64_t k[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = lane * 4 + r;
                const int id = first_round ? first + j : (int)ids[j < n_cur ? j : 0];
                k[r] = j < n_cur ? Key32::make(loss_at(j, id), j, desc).v : ~0ull;
            }
            warp_bitonic_regs<4>(k, lane);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = lane * 4 + r, pos = (int)(uint32_t)k[r];
                if (j < keep) ids_next[j] = (unsigned short)(first_round ? first + pos : (int)ids[pos]);
            }
            const int pos0 = __shfl_sync(0xffffffffu, (int)(uint32_t)k[0], 0);
            rb_arm = first_round ? first + pos0 : (int)ids[pos0];
            rb_loss = loss_at(pos0, rb_arm);
        } else {
            uint64_t k[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int j = lane * 8 + r;
                const int id = first_round ? first + j : (int)ids[j < n_cur ? j : 0];
                k[r] = j < n_cur ? Key32::make(loss_at(j, id), j, desc).v : ~0ull;
            }
            warp_bitonic_regs<8>(k, lane);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int j = lane * 8 + r, pos = (int)(uint32_t)k[r];
                if (j < keep) ids_next[j] = (unsigned short)(first_round ? first + pos : (int)ids[pos]);
            }
            const int pos0 = __shfl_sync(0xffffffffu, (int)(uint32_t)k[0], 0);
            rb_arm = first_round ? first + pos0 : (int)ids[pos0];
            rb_loss = loss_at(pos0, rb_arm);
        }
    } else {
        int npad = 512;
        while (npad < n_cur) npad <<= 1;
        for (int j = lane; j < npad; j += 32) {
            const int id = first_round ? first + j : (int)ids[j < n_cur ? j : 0];
            keys[j] = j < n_cur ? Key32::make(loss_at(j, id), j, desc) : Key32::pad();
        }
        __syncwarp();
        warp_bitonic_smem(keys, npad, lane);
        for (int j = lane; j < keep; j += 32) {
            const int pos = keys[j].pos();
            ids_next[j] = (unsigned short)(first_round ? first + pos : (int)ids[pos]);
        }
        const int pos0 = keys[0].pos();
        rb_arm = first_round ? first + pos0 : (int)ids[pos0];
        rb_loss = loss_at(pos0, rb_arm);
    }
}

template <typename T>
struct HalvingOut {
    T* ofe;
    int* ofe_arm;
    T* round_best;
    int* round_best_arm;
    int* survivors;
};

template <typename T>
struct HalvingState {
    T best;
    int best_arm;
    int surv_w;
};

// Runs the rounds of bracket `b` for one repetition whose checkpoint losses sit in shared memory as
// ck[c * stride + arm]; writes the requested per-round outputs for repetition `rep_g` and folds the round-bests into
// `st`.  keys: npad sort keys, ids_a/ids_b: npad arm ids each (warp-private scratch).  When `latest_ck` is given,
// latest_ck[arm] ends up holding the checkpoint index of the last round the arm was evaluated in (what the hybrids'
// history transfer keys on, hybrid_hyperband_tpe_with_transfer.py:96-98), with bit 7 set when that round's int(r_i) was 0
// and the read wrapped to the last point (the reference records the 0, not R).  All 32 lanes must call it.
template <typename T>
__device__ void warp_halving_bracket(const at2_plan& plan, int b, const T* ck, int stride, typename KeyOf<T>::type* keys,
                                     unsigned short* ids_a, unsigned short* ids_b, bool desc, int lane, long long rep_g,
                                     const HalvingOut<T>& out, HalvingState<T>& st, unsigned char* latest_ck = nullptr) {
    typedef typename KeyOf<T>::type K;
    int n_cur = plan.bracket_n[b];
    const int first = plan.bracket_first_arm[b];
    unsigned short* ids = ids_a;
    unsigned short* ids_next = ids_b;
    for (int rd = plan.bracket_first_round[b]; rd < plan.bracket_first_round[b + 1]; ++rd) {
        const int c = plan.round_ckpt[rd];
        const int keep = min(plan.round_keep[rd], n_cur);
        const bool first_round = rd == plan.bracket_first_round[b];
        const T* col = ck + c * stride;
        T rb_loss;
        int rb_arm;
        if (latest_ck != nullptr)
            for (int j = lane; j < n_cur; j += 32)
                latest_ck[first_round ? first + j : (int)ids[j]] = (unsigned char)(c | (plan.round_res[rd] < 1 ? 0x80 : 0));
        if constexpr (sizeof(T) == 4) {
            warp_halving_round_f32(n_cur, keep, first_round, first, ids, ids_next, keys, desc, lane,
                                   [&](int, int id) { return col[id]; }, rb_loss, rb_arm);
        } else if (n_cur <= 32) {
            K mine = K::pad();
            T my_loss = (T)0;
            int my_id = 0;
            if (lane < n_cur) {
                my_id = first_round ? first + lane : (int)ids[lane];
                my_loss = col[my_id];
                mine = K::make(my_loss, lane, desc);
            }
            mine = warp_bitonic_shfl(mine, lane, n_cur);
            const int src = mine.pos() & 31;            // lane that held the key now ranked `lane`
            const int sid = __shfl_sync(0xffffffffu, my_id, src);
            const T sl = __shfl_sync(0xffffffffu, my_loss, src);
            rb_loss = __shfl_sync(0xffffffffu, sl, 0);
            rb_arm = __shfl_sync(0xffffffffu, sid, 0);
            if (lane < keep) ids_next[lane] = (unsigned short)sid;
        } else {
            int npad = 64;
            while (npad < n_cur) npad <<= 1;
            for (int j = lane; j < npad; j += 32) {
                const int id = first_round ? first + j : (int)ids[j < n_cur ? j : 0];
                keys[j] = j < n_cur ? K::make(col[id], j, desc) : K::pad();
            }
            __syncwarp();
            warp_bitonic_smem(keys, npad, lane);
            for (int j = lane; j < keep; j += 32) {
                const int pos = keys[j].pos();
                ids_next[j] = (unsigned short)(first_round ? first + pos : (int)ids[pos]);
            }
            const int pos0 = keys[0].pos();
            rb_arm = first_round ? first + pos0 : (int)ids[pos0];
            rb_loss = col[rb_arm];
        }
        __syncwarp();
        if (out.survivors != nullptr)
            for (int j = lane; j < keep; j += 32) out.survivors[rep_g * plan.n_surv + st.surv_w + j] = (int)ids_next[j];
        st.surv_w += keep;
        if (lane == 0) {
            if (out.round_best != nullptr) out.round_best[rep_g * plan.n_rounds + rd] = rb_loss;
            if (out.round_best_arm != nullptr) out.round_best_arm[rep_g * plan.n_rounds + rd] = rb_arm;
        }
        // OFE = first optimum over the round-bests in history order (core/optimiser.py:67-68)
        if (st.best_arm < 0 || (desc ? rb_loss > st.best : rb_loss < st.best)) st.best = rb_loss, st.best_arm = rb_arm;
        unsigned short* tmp = ids;
        ids = ids_next;
        ids_next = tmp;
        n_cur = keep;
        __syncwarp();
    }
}

template <typename T>
__device__ __forceinline__ void halving_finish(const HalvingOut<T>& out, const HalvingState<T>& st, int lane, long long rep_g) {
    if (lane == 0) {
        if (out.ofe != nullptr) out.ofe[rep_g] = st.best;
        if (out.ofe_arm != nullptr) out.ofe_arm[rep_g] = st.best_arm;
    }
}

// All brackets of `plan` for one repetition; returns the arm id of the OFE.
template <typename T>
__device__ int warp_halving(const at2_plan& plan, const T* ck, int stride, typename KeyOf<T>::type* keys,
                            unsigned short* ids_a, unsigned short* ids_b, bool desc, int lane, long long rep_g,
                            const HalvingOut<T>& out, T* ofe_value = nullptr) {
    HalvingState<T> st{(T)0, -1, 0};
    for (int b = 0; b < plan.n_brackets; ++b)
        warp_halving_bracket<T>(plan, b, ck, stride, keys, ids_a, ids_b, desc, lane, rep_g, out, st);
    halving_finish<T>(out, st, lane, rep_g);
    if (ofe_value != nullptr) *ofe_value = st.best;
    return st.best_arm;
}

// Successive halving of the G repetitions of a warp's tile TOGETHER (float losses at ck[c * stride + g * A + arm]): the
// rounds of a bracket are walked once for the whole tile; a round of n <= 32 arms ranks 32 / n repetitions side by side
// (segments of n lanes, rank by counting with python's stable tie-break, as warp_rank_by_counting) and the lane that ranks
// first does its repetition's bookkeeping; wider rounds go repetition by repetition through warp_halving_round_f32.
// Results are those of warp_halving<float> per repetition, bit for bit.
//   ids: G * (km_a + km_b) entries - a repetition's survivor ids alternate between buffer A (km_a entries: survivors of a
//        bracket's rounds 0, 2, ..) and buffer B (km_b: rounds 1, 3, ..);  best / best_arm / surv_w: G entries each.
// On return best[g] / best_arm[g] hold repetition g's OFE and its arm (also written to out.ofe / out.ofe_arm).
__device__ __forceinline__ void tile_halving_f32(const at2_plan& plan, const float* ck, int stride, int A, int G, Key32* keys,
                                                 unsigned short* ids, int km_a, int km_b, float* best, int* best_arm,
                                                 int* surv_w, bool desc, int lane, long long rep0,
                                                 const HalvingOut<float>& out) {
    const int kms = km_a + km_b;
    auto idbuf = [&](int g, int fl) { return ids + g * kms + (fl ? 0 : km_a); };   // fl = 1: buffer A
    for (int g = lane; g < G; g += 32) best_arm[g] = -1, surv_w[g] = 0, best[g] = 0.0f;
    __syncwarp();
    for (int b = 0; b < plan.n_brackets; ++b) {
        const int first = plan.bracket_first_arm[b];
        int n_cur = plan.bracket_n[b];
        int flip = 0;
        for (int rd = plan.bracket_first_round[b]; rd < plan.bracket_first_round[b + 1]; ++rd) {
            const float* col = ck + plan.round_ckpt[rd] * stride;
            const bool first_round = rd == plan.bracket_first_round[b];
            const int keep = min(plan.round_keep[rd], n_cur);
            if (n_cur <= 32) {
                const int S = 32 / n_cur, seg = lane / n_cur, j = lane - seg * n_cur;
                const unsigned seg_mask = (n_cur == 32 ? 0xffffffffu : (1u << n_cur) - 1u) << (seg < S ? seg * n_cur : 0);
#pragma unroll 1
                for (int h0 = 0; h0 < G; h0 += S) {
                    const bool act = seg < S && h0 + seg < G;
                    const int g = act ? h0 + seg : 0;
                    const long long rep_g = rep0 + g;
                    const int my_id = first_round ? first + j : (int)idbuf(g, flip)[j];
                    const float my_loss = col[g * A + my_id];
                    const uint32_t ord = act ? orderable(my_loss, desc) : 0xffffffffu;
                    int rank = 0;
#pragma unroll 3
                    for (int k = 0; k < n_cur; ++k) rank += __shfl_sync(0xffffffffu, ord, seg * n_cur + k) < ord ? 1 : 0;
                    rank += __popc(__match_any_sync(0xffffffffu, ord) & seg_mask & ((1u << lane) - 1u));
                    const int sw = surv_w[g];
                    __syncwarp();
                    if (act && rank < keep) {
                        idbuf(g, flip ^ 1)[rank] = (unsigned short)my_id;
                        if (out.survivors != nullptr) out.survivors[rep_g * plan.n_surv + sw + rank] = my_id;
                    }
                    if (act && rank == 0) {
                        surv_w[g] = sw + keep;
                        if (out.round_best != nullptr) out.round_best[rep_g * plan.n_rounds + rd] = my_loss;
                        if (out.round_best_arm != nullptr) out.round_best_arm[rep_g * plan.n_rounds + rd] = my_id;
                        // OFE = first optimum over the round-bests in history order (core/optimiser.py:67-68)
                        const int ba = best_arm[g];
                        const float bl = best[g];
                        if (ba < 0 || (desc ? my_loss > bl : my_loss < bl)) best[g] = my_loss, best_arm[g] = my_id;
                    }
                    __syncwarp();
                }
            } else {
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const long long rep_g = rep0 + g;
                    const float* loss = col + g * A;
                    unsigned short* nxt = idbuf(g, flip ^ 1);
                    float rb_loss;
                    int rb_arm;
                    warp_halving_round_f32(n_cur, keep, first_round, first, idbuf(g, flip), nxt, keys, desc, lane,
                                           [&](int, int id) { return loss[id]; }, rb_loss, rb_arm);
                    __syncwarp();
                    const int sw = surv_w[g];
                    if (out.survivors != nullptr)
                        for (int j = lane; j < keep; j += 32) out.survivors[rep_g * plan.n_surv + sw + j] = (int)nxt[j];
                    if (lane == 0) {
                        surv_w[g] = sw + keep;
                        if (out.round_best != nullptr) out.round_best[rep_g * plan.n_rounds + rd] = rb_loss;
                        if (out.round_best_arm != nullptr) out.round_best_arm[rep_g * plan.n_rounds + rd] = rb_arm;
                        const int ba = best_arm[g];
                        const float bl = best[g];
                        if (ba < 0 || (desc ? rb_loss > bl : rb_loss < bl)) best[g] = rb_loss, best_arm[g] = rb_arm;
                    }
                    __syncwarp();
                }
            }
            flip ^= 1;
            n_cur = keep;
        }
    }
    for (int g = lane; g < G; g += 32) {
        if (out.ofe != nullptr) out.ofe[rep0 + g] = best[g];
        if (out.ofe_arm != nullptr) out.ofe_arm[rep0 + g] = best_arm[g];
    }
    __syncwarp();
}

// entries of the two survivor-id buffers tile_halving_f32 needs per repetition
inline void halving_id_buffer_sizes(const at2_plan* plan, int* km_a, int* km_b) {
    *km_a = 1, *km_b = 1;
    for (int b = 0; b < plan->n_brackets; ++b)
        for (int rd = plan->bracket_first_round[b]; rd < plan->bracket_first_round[b + 1]; ++rd) {
            int* km = ((rd - plan->bracket_first_round[b]) & 1) ? km_b : km_a;
            *km = plan->round_keep[rd] > *km ? plan->round_keep[rd] : *km;
        }
}

// bytes of warp-private sort scratch for brackets up to `npad` wide
template <typename T>
__host__ __device__ inline size_t halving_scratch_bytes(int npad) {
    return (size_t)npad * (sizeof(typename KeyOf<T>::type) + 2 * sizeof(unsigned short));
}

inline int halving_npad(const at2_plan* plan) {
    int widest = 0;
    for (int b = 0; b < plan->n_brackets; ++b) widest = plan->bracket_n[b] > widest ? plan->bracket_n[b] : widest;
    int npad = 64;
    while (npad < widest) npad <<= 1;
    return npad;
}

}  // namespace
