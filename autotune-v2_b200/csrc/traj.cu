// Full-trajectory output mode of the simulation: every position of every requested arm, for the per-evaluator API
// (`SimulationEvaluator.simulate`, `.non_smooth_fs`, core/simulation_evaluator.py:51-116) and for the loss-curve family
// profiles (experiments/simulation_evaluation/profiles.py).  The hot path never stores trajectories; this kernel exists
// for inspection and consumes exactly the same Philox stream and arithmetic as sim_core.cuh: the value it writes at a
// Hyperband checkpoint is bit-identical to the one the fused kernel reads there for the same stream id.
#include "at2_common.cuh"
#include "sim_core.cuh"

namespace {

__global__ void __launch_bounds__(128) traj_kernel(const float* __restrict__ tables, int R, int F, int n_ckpts_layout,
                                                   int has_smooth, Fam<float> f0, Fam<float> f1_, Fam<float> f2, Fam<float> f3,
                                                   Fam<float> f4, Fam<float> f5, Fam<float> f6, Fam<float> f7, long long n_arms,
                                                   const float* __restrict__ d_f1, const float* __restrict__ d_fn,
                                                   const int* __restrict__ d_family, const unsigned long long* __restrict__ d_gid,
                                                   uint32_t key0, uint32_t key1, float* __restrict__ raw) {
    extern __shared__ __align__(16) float s_tab[];
    const At2TableLayout L{R, F, has_smooth, at2_sg_stride(n_ckpts_layout)};
    for (int i = threadIdx.x; i < L.off_sg(); i += blockDim.x) s_tab[i] = tables[i];
    __syncthreads();
    const Fam<float> fams[8] = {f0, f1_, f2, f3, f4, f5, f6, f7};
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n_arms; a += (long long)gridDim.x * blockDim.x) {
        const Fam<float> fam = fams[d_family[a]];
        const float fn = d_fn[a], upc = fam.up;
        const unsigned long long gid = d_gid[a];
        const float* rec = s_tab + L.off_rec() + (size_t)AT2_REC * fam.pw_off;
        float ff = d_f1[a];
        float* out = raw + a * R;
        out[0] = ff;
        int t = 1;
        uint32_t ctr = 0;
        const uint32_t g_lo = (uint32_t)gid, g_hi = (uint32_t)(gid >> 32);
        // the word stream of sim_core.cuh: one block feeds two attempt pairs
        auto pair = [&](uint32_t wa, uint32_t wb) {
            float xs[2], ku[2];
            box_muller_pair(wa, wb, xs[0], xs[1], ku[0], ku[1]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (t >= R) break;
                const float4 gm = *reinterpret_cast<const float4*>(rec + AT2_REC * t);
                const float4 pw = *reinterpret_cast<const float4*>(rec + AT2_REC * t + 4);
                const AttemptOut o = mt_attempt(gm, pw, xs[k], ku[k], ff, fn, upc);
                if (o.ok) {
                    ff = o.fnew;
                    out[t] = ff;
                    ++t;
                }
            }
        };
        while (t < R) {
            const At2U4 A4 = at2_philox10(At2U4{ctr, g_lo, g_hi, AT2_STREAM_TRAJ}, key0, key1);
            ++ctr;
            pair(A4.x, A4.y);
            pair(A4.z, A4.w);
        }
    }
}

}  // namespace

extern "C" int at2_sim_trajectories_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_arms,
                                        const float* d_f1, const float* d_fn, const int32_t* d_family,
                                        const uint64_t* d_stream_id, uint64_t seed, float* d_raw, void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(plan && cfg && d_tables && d_f1 && d_fn && d_family && d_stream_id && d_raw, "at2_sim_trajectories_f32: NULL pointer");
    AT2_REQUIRE(n_arms > 0 && plan->R >= 2, "at2_sim_trajectories_f32: need n_arms > 0 and R >= 2");
    AT2_REQUIRE(cfg->n_families >= 1 && cfg->n_families <= AT2_MAX_FAMILIES, "at2_sim_trajectories_f32: n_families %d outside [1,%d]",
                cfg->n_families, AT2_MAX_FAMILIES);
    Fam<float> fam[8] = {};
    int slot[AT2_MAX_FAMILIES];
    const int n_slots = at2_family_slot_map(cfg, slot);
    for (int f = 0; f < cfg->n_families; ++f) {
        fam[f].ml = (float)(cfg->fam[f].ml_agg / 100.0);
        fam[f].up = (float)cfg->fam[f].up_spik;
        fam[f].s0 = (float)cfg->fam[f].start_shift;
        fam[f].s1 = (float)cfg->fam[f].end_shift;
        fam[f].smooth = cfg->fam[f].is_smooth;
        fam[f].pw_off = slot[f] * at2_rec_rows(plan->R);
    }
    const int has_smooth = at2_cfg_has_smooth(cfg);
    At2TableLayout L{plan->R, n_slots, has_smooth, at2_sg_stride(plan->n_ckpts)};
    const size_t smem = (size_t)L.off_sg() * sizeof(float);
    AT2_CUDA_CHECK(cudaFuncSetAttribute(traj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (n_arms + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
    traj_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>((const float*)d_tables, plan->R, n_slots, plan->n_ckpts,
                                                                      has_smooth, fam[0], fam[1], fam[2], fam[3], fam[4], fam[5],
                                                                      fam[6], fam[7], n_arms, d_f1, d_fn, d_family,
                                                                      (const unsigned long long*)d_stream_id, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), d_raw);
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}
