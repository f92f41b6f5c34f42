// Post-processing statistics of the Monte-Carlo outputs ("next" rows of SURVEY 8f):
//   * order profiles of a family of loss curves - experiments/simulation_evaluation/profiles.py:29-81
//     (get_dynamic_order_profile is O(F^2 T) pairwise rank agreement between consecutive epochs,
//     get_ends_order_profile the same between the first and the last epoch);
//   * the two-sample Kolmogorov-Smirnov statistic the reference uses to compare OFE vectors
//     (experiments/simulation_evaluation/plot_run_simulations_results.py:69-73 -> scipy.stats.ks_2samp).
#include <cfloat>

#include "at2_common.cuh"

namespace {

// out[t] for t in [0, n_cols): fraction of ordered pairs (i, j), i != j, whose strict order in column col_b(t) equals
// their strict order in column col_a(t).  Integer pair counts -> exact, deterministic.
__global__ void __launch_bounds__(256) order_profile_kernel(const float* __restrict__ curves, int n, int T, int n_out,
                                                            int ends_mode, float* __restrict__ out) {
    extern __shared__ float s_col[];   // [2][n]: the two columns being compared
    const int t = blockIdx.x;
    const int ca = ends_mode ? 0 : t, cb = ends_mode ? T - 1 : t + 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_col[i] = curves[(size_t)i * T + ca];
        s_col[n + i] = curves[(size_t)i * T + cb];
    }
    __syncthreads();
    unsigned long long agree = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float ai = s_col[i], bi = s_col[n + i];
        int cnt = 0;
        for (int j = 0; j < n; ++j) {
            const float aj = s_col[j], bj = s_col[n + j];
            cnt += ((aj < ai) && (bj < bi)) || ((aj > ai) && (bj > bi)) ? 1 : 0;
        }
        agree += cnt;
    }
    __shared__ unsigned long long s_red[256];
    s_red[threadIdx.x] = agree;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && t < n_out)
        out[t] = (float)((double)s_red[0] / ((double)n * (double)(n - 1)));
}

__device__ __forceinline__ long long upper_bound(const float* __restrict__ a, long long n, float x) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// D = max over pooled points x of |F_a(x) - F_b(x)|, F(x) = #{<= x}/n  (scipy.stats.ks_2samp, two-sided statistic)
__global__ void __launch_bounds__(256) ks_statistic_kernel(const float* __restrict__ a, long long n1,
                                                           const float* __restrict__ b, long long n2,
                                                           unsigned int* __restrict__ d_bits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float diff = 0.0f;
    if (i < n1 + n2) {
        const float x = i < n1 ? a[i] : b[i - n1];
        const double fa = (double)upper_bound(a, n1, x) / (double)n1;
        const double fb = (double)upper_bound(b, n2, x) / (double)n2;
        diff = (float)fabs(fa - fb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff = fmaxf(diff, __shfl_xor_sync(0xffffffffu, diff, o));
    if ((threadIdx.x & 31) == 0) atomicMax(d_bits, __float_as_uint(diff));   // non-negative floats order like their bits
}

}  // namespace

extern "C" int at2_order_profile_f32(const float* d_curves, int32_t n_curves, int32_t T, float* d_dynamic,
                                     float* d_ends, void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(d_curves && (d_dynamic || d_ends), "at2_order_profile_f32: NULL pointer");
    AT2_REQUIRE(n_curves >= 2 && T >= 2, "at2_order_profile_f32: need at least 2 curves of at least 2 points");
    const size_t smem = 2 * (size_t)n_curves * sizeof(float);
    AT2_REQUIRE(smem <= 200 * 1024, "at2_order_profile_f32: more than %d curves", (int)(200 * 1024 / 8));
    AT2_CUDA_CHECK(cudaFuncSetAttribute(order_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    if (d_dynamic != nullptr) {
        order_profile_kernel<<<T - 1, 256, smem, st>>>(d_curves, n_curves, T, T - 1, 0, d_dynamic);
        ++launches;
    }
    if (d_ends != nullptr) {
        order_profile_kernel<<<1, 256, smem, st>>>(d_curves, n_curves, T, 1, 1, d_ends);
        ++launches;
    }
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(launches);
    return AT2_OK;
}

extern "C" int at2_ks_statistic_f32(const float* d_a_sorted, int64_t n1, const float* d_b_sorted, int64_t n2,
                                    float* d_statistic, void* stream) {
    at2_reset_launch_count();
    AT2_REQUIRE(d_a_sorted && d_b_sorted && d_statistic, "at2_ks_statistic_f32: NULL pointer");
    AT2_REQUIRE(n1 > 0 && n2 > 0, "at2_ks_statistic_f32: both samples must be non-empty");
    cudaStream_t st = (cudaStream_t)stream;
    AT2_CUDA_CHECK(cudaMemsetAsync(d_statistic, 0, sizeof(float), st));
    const long long n = n1 + n2;
    ks_statistic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a_sorted, n1, d_b_sorted, n2,
                                                                     reinterpret_cast<unsigned int*>(d_statistic));
    AT2_CUDA_CHECK(cudaGetLastError());
    at2_count_launch(1);
    return AT2_OK;
}
