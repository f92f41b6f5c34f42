// (d) Closest-known-loss-function approximation: batched nearest recorded arm + Hyperband on top of it.
//
// Reference: benchmarks/known_loss_fn_problem.py:22-62 (KnownFnEvaluator.evaluate) with core/arm.py:57-70
// (Arm.normalize): for every recorded arm, sum over the proposed arm's hyperparameters (in its attribute order) of
// ((p - min)/(max - min) - (a - min)/(max - min))**2, python sum() left to right, strict `<` so the FIRST minimum
// wins; the loss returned is known_fs[arm*][int(n_resources) - 1].  All of it is float64 in the reference, and the
// device code below executes the same IEEE operations un-fused (__dsub_rn/__ddiv_rn/__dmul_rn/__dadd_rn), so the
// nearest index is bit-exact - including the log-scale quirk (bounds are exponents, values are linear).
//
// At the reference's shapes (A = 425 x D = 4, A = 300 x D = 9) the normalised database is 14-22 KB: it lives in
// shared memory, every lane scans it with broadcast reads, and the kernel is FP64-pipe bound (3 A D flops per query).
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <math_constants.h>

#include "at2_common.cuh"
#include "halving.cuh"

namespace {

struct DomainDev {
    int D;
    double lo[AT2_MAX_PARAMS], span[AT2_MAX_PARAMS];     // min_val, max_val - min_val
    double interval[AT2_MAX_PARAMS], dflt[AT2_MAX_PARAMS];
    int is_log[AT2_MAX_PARAMS], optimised[AT2_MAX_PARAMS], order[AT2_MAX_PARAMS];
};

int make_domain(const at2_domain* dom, DomainDev* d) {
    AT2_REQUIRE(dom != nullptr, "NULL domain");
    AT2_REQUIRE(dom->n_params >= 1 && dom->n_params <= AT2_MAX_PARAMS, "n_params %d outside [1,%d]", dom->n_params,
                AT2_MAX_PARAMS);
    d->D = dom->n_params;
    unsigned seen = 0;
    for (int i = 0; i < d->D; ++i) {
        const at2_param& q = dom->p[i];
        AT2_REQUIRE(q.max_val != q.min_val, "hyperparameter %d has an empty range", i);
        d->lo[i] = q.min_val;
        d->span[i] = q.max_val - q.min_val;                 // arm.py:69 evaluates (max_val - min_val) the same way
        d->interval[i] = q.interval;
        d->dflt[i] = q.default_val;
        d->is_log[i] = q.is_log;
        d->optimised[i] = q.optimised;
        const int o = dom->order[i];
        AT2_REQUIRE(o >= 0 && o < d->D && !(seen & (1u << o)), "domain.order is not a permutation of 0..%d", d->D - 1);
        seen |= 1u << o;
        d->order[i] = o;
    }
    return AT2_OK;
}

// Nearest row of a normalised table to a normalised query.  Both are laid out in SUMMATION order (position k holds
// hyperparameter dom.order[k], the k-th attribute of a drawn arm), so the reference's left-to-right python sum() of squared
// differences (known_loss_fn_problem.py:42-47; 0 + t0 is exact) is a straight walk: s[cnt][D] rows, qo[D] query.  Strict
// `<` keeps the FIRST minimum; rows whose global index base + a falls in [ex_lo, ex_hi) are skipped (cross-validation).
// DT = D at compile time: the query sits in registers and the row loop is unrolled (the scan is then FP64-pipe bound:
// 3 D - 1 DP operations per row); DT = 0 is the generic fallback.
template <int DT>
__device__ __forceinline__ void scan_rows(const double* qo, const double* __restrict__ s, int cnt, int D, int base, int ex_lo,
                                          int ex_hi, double& best, int& best_i) {
    if constexpr (DT > 0) {
        double q[DT];
#pragma unroll
        for (int k = 0; k < DT; ++k) q[k] = qo[k];
#pragma unroll 2
        for (int a = 0; a < cnt; ++a) {
            const double* r = s + a * DT;
            double diff = __dsub_rn(q[0], r[0]);
            double e = __dmul_rn(diff, diff);
#pragma unroll
            for (int k = 1; k < DT; ++k) {
                diff = __dsub_rn(q[k], r[k]);
                e = __dadd_rn(e, __dmul_rn(diff, diff));
            }
            const int ga = base + a;
            if (e < best && !(ga >= ex_lo && ga < ex_hi)) best = e, best_i = ga;
        }
    } else {
        for (int a = 0; a < cnt; ++a) {
            const double* r = s + a * D;
            double e = 0.0;
            for (int k = 0; k < D; ++k) {
                const double diff = __dsub_rn(qo[k], r[k]);
                e = __dadd_rn(e, __dmul_rn(diff, diff));
            }
            const int ga = base + a;
            if (e < best && !(ga >= ex_lo && ga < ex_hi)) best = e, best_i = ga;
        }
    }
}
__device__ __forceinline__ void scan_rows_dispatch(const double* qo, const double* __restrict__ s, int cnt, int D, int base,
                                                   int ex_lo, int ex_hi, double& best, int& best_i) {
    switch (D) {   // the reference's recorded-arm shapes: MNIST D = 4, CIFAR D = 9; test functions D = 2; D = 8: bench
        case 2: return scan_rows<2>(qo, s, cnt, D, base, ex_lo, ex_hi, best, best_i);
        case 4: return scan_rows<4>(qo, s, cnt, D, base, ex_lo, ex_hi, best, best_i);
        case 8: return scan_rows<8>(qo, s, cnt, D, base, ex_lo, ex_hi, best, best_i);
        case 9: return scan_rows<9>(qo, s, cnt, D, base, ex_lo, ex_hi, best, best_i);
        default: return scan_rows<0>(qo, s, cnt, D, base, ex_lo, ex_hi, best, best_i);
    }
}

__device__ __forceinline__ double normalise(double v, int h, const DomainDev& dom) {
    return __ddiv_rn(__dsub_rn(v, dom.lo[h]), dom.span[h]);
}

// ---- standalone batched lookup ----------------------------------------------------------------------------------------
constexpr int NN_BLOCK = 128;
constexpr int NN_TILE_BYTES = 32 * 1024;

__global__ void __launch_bounds__(NN_BLOCK) nn_lookup_kernel(const __grid_constant__ DomainDev dom, const double* __restrict__ db,
                                                            int A, const double* __restrict__ queries, long long nq,
                                                            int n_chunks, int chunk_len, int tile_a,
                                                            double* __restrict__ part_dist, int* __restrict__ part_idx,
                                                            const int* __restrict__ exclude) {
    extern __shared__ __align__(16) double s_db[];   // [tile_a][D] normalised
    const int D = dom.D;
    const long long qi = (long long)blockIdx.x * NN_BLOCK + threadIdx.x;
    const bool active = qi < nq;
    double q[AT2_MAX_PARAMS];   // summation order
    for (int k = 0; k < D; ++k) {
        const int h = dom.order[k];
        q[k] = active ? normalise(queries[qi * D + h], h, dom) : 0.0;
    }
    const int a_lo = blockIdx.y * chunk_len, a_hi = min(A, a_lo + chunk_len);
    const int ex_lo = (exclude != nullptr && active) ? exclude[2 * qi] : 0;
    const int ex_hi = (exclude != nullptr && active) ? exclude[2 * qi + 1] : 0;
    double best = CUDART_INF;
    int best_i = -1;
    for (int base = a_lo; base < a_hi; base += tile_a) {
        const int cnt = min(tile_a, a_hi - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * D; i += NN_BLOCK) {
            const int a = i / D, h = dom.order[i - a * D];
            s_db[i] = normalise(db[(size_t)(base + a) * D + h], h, dom);
        }
        __syncthreads();
        scan_rows_dispatch(q, s_db, cnt, D, base, ex_lo, ex_hi, best, best_i);
    }
    if (active) {
        part_dist[qi * n_chunks + blockIdx.y] = best;
        part_idx[qi * n_chunks + blockIdx.y] = best_i;
    }
}

__global__ void nn_finish_kernel(const double* __restrict__ part_dist, const int* __restrict__ part_idx, long long nq,
                                 int n_chunks, int* __restrict__ idx_out, double* __restrict__ dist_out,
                                 const double* __restrict__ curves, int L, const int* __restrict__ time,
                                 double* __restrict__ loss_out) {
    const long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    double best = CUDART_INF;
    int best_i = -1;
    for (int c = 0; c < n_chunks; ++c) {      // chunks are in database order, so strict < keeps the first minimum
        const double e = part_dist[qi * n_chunks + c];
        if (e < best) best = e, best_i = part_idx[qi * n_chunks + c];
    }
    idx_out[qi] = best_i;
    if (dist_out != nullptr) dist_out[qi] = best;
    if (loss_out != nullptr) {
        int t = time[qi];                     // int(n_resources); python index time-1 wraps for time <= 0
        int pos = t - 1;
        if (pos < 0) pos += L;
        loss_out[qi] = curves[(size_t)best_i * L + pos];
    }
}

// ---- Hyperband on the closest-known-function problem, fused ---------------------------------------------------------------
struct KfParams {
    at2_plan plan;
    DomainDev dom;
    const double* db;        // [A][D] raw recorded arms (domain order)
    const double* curves;    // [A][L]
    int A, L, descending;
    long long n_reps, rep_offset;
    uint32_t key0, key1;
    const double* predrawn;  // [n_reps][n_arms][D] raw proposed arms, or NULL
    int G, npad, stride, warp_bytes;
    long long n_tiles;
    double* ofe;
    int* ofe_arm;
    double* round_best;
    int* round_best_arm;
    int* survivors;
    double* ckpt_loss;
    int* nn_idx;             // [n_reps][n_arms]
    double* best_arm_vals;   // [n_reps][D] hyperparameters of the OFE arm
};

// one proposed arm from the Philox stream of (repetition, arm): core/arm.py:44-55, core/params.py:36-45
__device__ void draw_proposed(const KfParams& p, unsigned long long gid, double* vals) {
    const DomainDev& dom = p.dom;
    At2U4 r = {0, 0, 0, 0};
    int used = 4, blk = 0;
    for (int h = 0; h < dom.D; ++h) {
        if (!dom.optimised[h]) {
            vals[h] = dom.dflt[h];
            continue;
        }
        if (used == 4) {
            r = at2_philox10(At2U4{(uint32_t)blk, (uint32_t)gid, (uint32_t)(gid >> 32), AT2_STREAM_ARM}, p.key0, p.key1);
            ++blk;
            used = 0;
        }
        const uint32_t w = used == 0 ? r.x : used == 1 ? r.y : used == 2 ? r.z : r.w;
        ++used;
        double v = (double)w * 2.3283064365386963e-10 * dom.span[h] + dom.lo[h];   // rand * (max - min) + min
        if (dom.is_log[h]) v = exp(v);                                               // logbase ** v, logbase = e
        if (dom.interval[h] != 0.0) v = floor(v / dom.interval[h]) * dom.interval[h];
        vals[h] = v;
    }
}

__global__ void __launch_bounds__(256) hb_knownfn_kernel(const __grid_constant__ KfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A_arms = p.plan.n_arms, C = p.plan.n_ckpts, D = p.dom.D;
    double* s_db = reinterpret_cast<double*>(smem_raw);                       // [A][D] normalised
    unsigned char* warp_base = smem_raw + (((size_t)p.A * D * sizeof(double) + 15) & ~(size_t)15);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < p.A * D; i += blockDim.x) {           // summation order, see scan_rows
        const int a = i / D, h = p.dom.order[i - a * D];
        s_db[i] = normalise(p.db[a * D + h], h, p.dom);
    }
    __syncthreads();
    double* s_ck = reinterpret_cast<double*>(warp_base + (size_t)warp * p.warp_bytes);   // [C][stride]
    Key64* keys = reinterpret_cast<Key64*>(s_ck + (size_t)C * p.stride);
    unsigned short* ids_a = reinterpret_cast<unsigned short*>(keys + p.npad);
    unsigned short* ids_b = ids_a + p.npad;
    const bool desc = p.descending != 0;

    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < p.n_tiles; tile += (long long)gridDim.x * nwarps) {
        const long long rep0 = tile * p.G;
        const int g_eff = (int)min((long long)p.G, p.n_reps - rep0);
        const int n_items = g_eff * A_arms;
        for (int base = 0; base < n_items; base += 32) {
            const int i = base + lane;
            if (i < n_items) {
                const int rep_l = i / A_arms, arm = i - rep_l * A_arms;
                const long long rep_g = rep0 + rep_l;
                double vals[AT2_MAX_PARAMS], q[AT2_MAX_PARAMS];
                if (p.predrawn != nullptr) {
                    for (int h = 0; h < D; ++h) vals[h] = p.predrawn[(rep_g * A_arms + arm) * D + h];
                } else {
                    draw_proposed(p, (unsigned long long)(p.rep_offset + rep_g) * (unsigned long long)A_arms + arm, vals);
                }
                for (int k = 0; k < D; ++k) {
                    const int h = p.dom.order[k];
                    q[k] = normalise(vals[h], h, p.dom);
                }
                double best = CUDART_INF;
                int best_i = 0;
                scan_rows_dispatch(q, s_db, p.A, D, 0, 0, 0, best, best_i);
                if (p.nn_idx != nullptr) p.nn_idx[rep_g * A_arms + arm] = best_i;
                for (int c = 0; c < C; ++c) {
                    const double loss = p.curves[(size_t)best_i * p.L + (p.plan.ckpt_time[c] - 1)];
                    s_ck[c * p.stride + i] = loss;
                    if (p.ckpt_loss != nullptr) p.ckpt_loss[(rep_g * A_arms + arm) * C + c] = loss;
                }
            }
        }
        __syncwarp();
        for (int rep_l = 0; rep_l < g_eff; ++rep_l) {
            const long long rep_g = rep0 + rep_l;
            HalvingOut<double> ho{p.ofe, p.ofe_arm, p.round_best, p.round_best_arm, p.survivors};
            const int best_arm = warp_halving<double>(p.plan, s_ck + rep_l * A_arms, p.stride, keys, ids_a, ids_b, desc,
                                                      lane, rep_g, ho);
            if (lane == 0 && p.best_arm_vals != nullptr) {
                double vals[AT2_MAX_PARAMS];
                if (p.predrawn != nullptr) {
                    for (int h = 0; h < D; ++h) vals[h] = p.predrawn[(rep_g * A_arms + best_arm) * D + h];
                } else {
                    draw_proposed(p, (unsigned long long)(p.rep_offset + rep_g) * (unsigned long long)A_arms + best_arm, vals);
                }
                for (int h = 0; h < D; ++h) p.best_arm_vals[rep_g * D + h] = vals[h];
            }
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int64_t at2_nn_workspace_bytes(int64_t nq, int32_t A, int32_t D) {
    if (nq <= 0 || A <= 0 || D <= 0) return 0;
    // worst case: the database split into 64 chunks
    return nq * 64 * (int64_t)(sizeof(double) + sizeof(int));
}

extern "C" int at2_nn_lookup_f64(const at2_domain* domain, const double* d_db, int32_t A, const double* d_queries,
                                 int64_t nq, int32_t* d_idx, double* d_dist, const double* d_curves, int32_t L,
                                 const int32_t* d_time, double* d_loss, void* d_workspace, int64_t workspace_bytes,
                                 void* stream) {
    at2_reset_launch_count();
    DomainDev dom;
    const int rc = make_domain(domain, &dom);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(d_db && d_queries && d_idx && d_workspace, "at2_nn_lookup_f64: NULL pointer");
    AT2_REQUIRE(A > 0 && nq > 0, "at2_nn_lookup_f64: need A > 0 recorded arms and nq > 0 queries");
    AT2_REQUIRE((d_loss == nullptr) || (d_curves && d_time && L > 0), "at2_nn_lookup_f64: loss gather needs curves, L and time");
    const int D = dom.D;
    const long long qblocks = (nq + NN_BLOCK - 1) / NN_BLOCK;
    // split the database when there are too few query blocks to fill the GPU
    int n_chunks = 1;
    while (n_chunks < 64 && qblocks * n_chunks < 148 * 4 && A / (n_chunks * 2) >= 64) n_chunks *= 2;
    const int chunk_len = (A + n_chunks - 1) / n_chunks;
    int tile_a = NN_TILE_BYTES / (D * (int)sizeof(double));
    if (tile_a > chunk_len) tile_a = chunk_len;
    const int64_t need = nq * n_chunks * (int64_t)(sizeof(double) + sizeof(int));
    AT2_REQUIRE(workspace_bytes >= need, "at2_nn_lookup_f64: workspace too small (%lld < %lld bytes)",
                (long long)workspace_bytes, (long long)need);
    double* part_dist = (double*)d_workspace;
    int* part_idx = (int*)(part_dist + nq * n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)tile_a * D * sizeof(double);
    nn_lookup_kernel<<<dim3((unsigned)qblocks, n_chunks), NN_BLOCK, smem, st>>>(dom, d_db, A, d_queries, nq, n_chunks,
                                                                                 chunk_len, tile_a, part_dist, part_idx,
                                                                                 nullptr);
    AT2_CUDA_CHECK(cudaGetLastError());
    nn_finish_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(part_dist, part_idx, nq, n_chunks, d_idx, d_dist,
                                                                   d_curves, L, d_time, d_loss);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(2);
    return AT2_OK;
}

extern "C" int at2_nn_lookup_excluding_f64(const at2_domain* domain, const double* d_db, int32_t A, const double* d_queries,
                                           int64_t nq, const int32_t* d_exclude, int32_t* d_idx, double* d_dist,
                                           void* d_workspace, int64_t workspace_bytes, void* stream) {
    at2_reset_launch_count();
    DomainDev dom;
    const int rc = make_domain(domain, &dom);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(d_db && d_queries && d_idx && d_workspace && d_exclude, "at2_nn_lookup_excluding_f64: NULL pointer");
    AT2_REQUIRE(A > 0 && nq > 0, "at2_nn_lookup_excluding_f64: need A > 0 recorded arms and nq > 0 queries");
    const int D = dom.D;
    const long long qblocks = (nq + NN_BLOCK - 1) / NN_BLOCK;
    int tile_a = NN_TILE_BYTES / (D * (int)sizeof(double));
    if (tile_a > A) tile_a = A;
    const int64_t need = nq * (int64_t)(sizeof(double) + sizeof(int));
    AT2_REQUIRE(workspace_bytes >= need, "at2_nn_lookup_excluding_f64: workspace too small");
    double* part_dist = (double*)d_workspace;
    int* part_idx = (int*)(part_dist + nq);
    cudaStream_t st = (cudaStream_t)stream;
    nn_lookup_kernel<<<dim3((unsigned)qblocks, 1), NN_BLOCK, (size_t)tile_a * D * sizeof(double), st>>>(
        dom, d_db, A, d_queries, nq, 1, A, tile_a, part_dist, part_idx, d_exclude);
    AT2_CUDA_CHECK(cudaGetLastError());
    nn_finish_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(part_dist, part_idx, nq, 1, d_idx, d_dist, nullptr, 0,
                                                                   nullptr, nullptr);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(2);
    return AT2_OK;
}

extern "C" int at2_hb_knownfn_f64(const at2_plan* plan, const at2_domain* domain, const double* d_db, int32_t A,
                                  const double* d_curves, int32_t L, int32_t descending, int64_t n_reps,
                                  int64_t rep_offset, uint64_t seed, const double* d_predrawn_arms,
                                  const at2_hb_outputs* out, int32_t* d_nn_idx, double* d_best_arm_vals, void* stream) {
    at2_reset_launch_count();
    KfParams p = {};
    const int rc = make_domain(domain, &p.dom);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(plan && d_db && d_curves && out, "at2_hb_knownfn_f64: NULL plan/database/curves/outputs");
    AT2_REQUIRE(A > 0 && L > 0 && n_reps > 0 && rep_offset >= 0, "at2_hb_knownfn_f64: bad sizes");
    AT2_REQUIRE(plan->n_arms > 0 && plan->n_arms < 65536 && plan->n_ckpts > 0, "at2_hb_knownfn_f64: plan is empty or malformed");
    for (int c = 0; c < plan->n_ckpts; ++c)
        AT2_REQUIRE(plan->ckpt_time[c] >= 1 && plan->ckpt_time[c] <= L,
                    "at2_hb_knownfn_f64: checkpoint %d exceeds the recorded curve length %d (python would raise IndexError)",
                    plan->ckpt_time[c], L);
    p.plan = *plan;
    p.db = d_db, p.curves = d_curves, p.A = A, p.L = L, p.descending = descending;
    p.n_reps = n_reps, p.rep_offset = rep_offset;
    p.key0 = (uint32_t)seed, p.key1 = (uint32_t)(seed >> 32);
    p.predrawn = d_predrawn_arms;
    p.ofe = (double*)out->d_ofe, p.ofe_arm = out->d_ofe_arm, p.round_best = (double*)out->d_round_best;
    p.round_best_arm = out->d_round_best_arm, p.survivors = out->d_survivors, p.ckpt_loss = (double*)out->d_ckpt_loss;
    p.nn_idx = d_nn_idx, p.best_arm_vals = d_best_arm_vals;
    int dev = 0, max_smem = 0, sms = 0;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int nwarps = 4;
    {
        const int w = at2_debug().kf_warps_per_cta;
        if (w > 0 && w <= 8) nwarps = w;
    }
    p.npad = halving_npad(plan);
    p.G = 1;
    {
        if (at2_debug().kf_reps_per_tile > 0) p.G = at2_debug().kf_reps_per_tile;
        if ((long long)p.G > n_reps) p.G = (int)n_reps;
    }
    p.stride = (p.G * plan->n_arms + 31) & ~31;
    p.warp_bytes = (int)(((size_t)plan->n_ckpts * p.stride * sizeof(double) + halving_scratch_bytes<double>(p.npad) + 15) & ~(size_t)15);
    p.n_tiles = (n_reps + p.G - 1) / p.G;
    const size_t db_bytes = ((size_t)A * p.dom.D * sizeof(double) + 15) & ~(size_t)15;
    const size_t smem = db_bytes + (size_t)nwarps * p.warp_bytes;
    if (smem > (size_t)max_smem) {
        at2_set_error("at2_hb_knownfn_f64: recorded-arm database (%d x %d) + halving scratch need %zu B of shared memory, limit %d B",
                      A, p.dom.D, smem, max_smem);
        return AT2_ELIMIT;
    }
    AT2_CUDA_CHECK(cudaFuncSetAttribute(hb_knownfn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 12) per_sm = 12;
    long long grid = (long long)sms * per_sm;
    const long long need = (p.n_tiles + nwarps - 1) / nwarps;
    if (grid > need) grid = need;
    hb_knownfn_kernel<<<(unsigned)grid, nwarps * 32, smem, (cudaStream_t)stream>>>(p);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}
