// (b) alone: the successive-halving selection as a stand-alone entry point - "keep the `keep` best of each segment, in
// python's stable sorted() order" - so that the survivor indices of the fused kernels can be checked on their own
// (at2_segmented_topk_*).  Reference: optimisers/sequential/hyperband_optimiser.py:46-57
//   sorted(evaluators, key=loss, reverse=...)[:n]   (stable; reverse=True keeps the original order of equal losses too)
// One warp per segment, running exactly the primitives the fused kernels run (halving.cuh): float segments of <= 32
// losses are ranked by counting, <= 128 / <= 256 sorted by the register bitonic network (4 / 8 keys per lane), wider ones
// (and every double segment wider than a warp) by the shared-memory network; double segments of <= 32 by the shuffle network.
#include "at2_common.cuh"
#include "halving.cuh"

namespace {

template <typename T>
struct TopkParams {
    const T* losses;
    const long long* seg_offsets;   // [n_segments + 1]
    const int* keep;                // [n_segments]
    long long n_segments;
    int descending;
    int npad;                       // power of two >= max_seg_len (shared-memory keys per warp), 0 when no segment needs them
    int* idx_out;                   // same layout as losses
    int* error_flag;                // set to 1 by a segment longer than the declared maximum
};

template <typename T>
__global__ void __launch_bounds__(256) segmented_topk_kernel(const TopkParams<T> p) {
    typedef typename KeyOf<T>::type K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    K* keys = reinterpret_cast<K*>(smem_raw) + (size_t)warp * p.npad;
    const bool desc = p.descending != 0;
    for (long long s = (long long)blockIdx.x * wpc + warp; s < p.n_segments; s += (long long)gridDim.x * wpc) {
        const long long o0 = p.seg_offsets[s], o1 = p.seg_offsets[s + 1];
        const long long n64 = o1 - o0;
        if (n64 <= 0) continue;
        const T* x = p.losses + o0;
        int* out = p.idx_out + o0;
        constexpr int REG_MAX = sizeof(T) == 4 ? 256 : 32;     // widest segment sorted without shared memory
        if (n64 > REG_MAX && n64 > p.npad) {                   // the caller's max_seg_len was wrong: say so, do not guess
            if (lane == 0) atomicExch(p.error_flag, 1);
            continue;
        }
        const int n = (int)n64;
        int keep = p.keep[s];
        keep = keep < 0 ? 0 : (keep > n ? n : keep);           // python slicing [:keep] of n items (keep >= 0)
        for (int j = keep + lane; j < n; j += 32) out[j] = -1;
        if constexpr (sizeof(T) == 4) {
            if (n <= 32) {
                const uint32_t ord = lane < n ? orderable(x[lane], desc) : 0xffffffffu;
                const int rank = warp_rank_by_counting(ord, lane, n);
                if (lane < n && rank < keep) out[rank] = lane;
                continue;
            }
            if (n <= 128) {
                uint64_t k[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int j = lane * 4 + r;
                    k[r] = j < n ? Key32::make(x[j], j, desc).v : ~0ull;
                }
                warp_bitonic_regs<4>(k, lane);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (lane * 4 + r < keep) out[lane * 4 + r] = (int)(uint32_t)k[r];
                continue;
            }
            if (n <= 256) {
                uint64_t k[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int j = lane * 8 + r;
                    k[r] = j < n ? Key32::make(x[j], j, desc).v : ~0ull;
                }
                warp_bitonic_regs<8>(k, lane);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (lane * 8 + r < keep) out[lane * 8 + r] = (int)(uint32_t)k[r];
                continue;
            }
        } else {
            if (n <= 32) {
                K mine = lane < n ? K::make(x[lane], lane, desc) : K::pad();
                mine = warp_bitonic_shfl(mine, lane, n);
                if (lane < keep) out[lane] = mine.pos();
                continue;
            }
        }
        int npad = 64;
        while (npad < n) npad <<= 1;
        for (int j = lane; j < npad; j += 32) keys[j] = j < n ? K::make(x[j], j, desc) : K::pad();
        __syncwarp();
        warp_bitonic_smem(keys, npad, lane);
        for (int j = lane; j < keep; j += 32) out[j] = keys[j].pos();
        __syncwarp();
    }
}

template <typename T>
int segmented_topk(const T* d_losses, const int64_t* d_seg_offsets, const int32_t* d_keep, int64_t n_segments,
                   int32_t max_seg_len, int32_t descending, int32_t* d_idx_out, void* stream, const char* who) {
    typedef typename KeyOf<T>::type K;
    at2_reset_launch_count();
    AT2_REQUIRE(n_segments >= 0, "%s: n_segments must be >= 0 (got %lld)", who, (long long)n_segments);
    if (n_segments == 0) return AT2_OK;
    AT2_REQUIRE(d_losses && d_seg_offsets && d_keep && d_idx_out, "%s: NULL losses/seg_offsets/keep/idx_out", who);
    AT2_REQUIRE(max_seg_len >= 1 && max_seg_len <= 65536, "%s: max_seg_len %d outside [1, 65536]", who, max_seg_len);
    cudaStream_t st = (cudaStream_t)stream;
    TopkParams<T> p = {};
    p.losses = d_losses, p.seg_offsets = (const long long*)d_seg_offsets, p.keep = d_keep, p.n_segments = n_segments;
    p.descending = descending, p.idx_out = d_idx_out;
    const int reg_max = sizeof(T) == 4 ? 256 : 32;
    if (max_seg_len > reg_max) {
        p.npad = 64;
        while (p.npad < max_seg_len) p.npad <<= 1;
    }
    int dev = 0, sms = 0, max_smem = 0;
    AT2_CUDA_CHECK(cudaGetDevice(&dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AT2_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int warps = 8;
    while (warps > 1 && (size_t)warps * p.npad * sizeof(K) > (size_t)max_smem) warps >>= 1;
    const size_t smem = (size_t)warps * p.npad * sizeof(K);
    AT2_REQUIRE(smem <= (size_t)max_smem, "%s: a %d-entry segment needs %zu B of shared memory; device limit %d B", who,
                max_seg_len, smem, max_smem);
    auto kern = segmented_topk_kernel<T>;
    AT2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* d_flag = nullptr;
    AT2_CUDA_CHECK(cudaMallocAsync(&d_flag, sizeof(int), st));
    AT2_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    p.error_flag = d_flag;
    long long grid = (n_segments + warps - 1) / warps;
    const long long cap = (long long)sms * 8;
    if (grid > cap) grid = cap;
    kern<<<(int)grid, warps * 32, smem, st>>>(p);
    cudaError_t rc = cudaGetLastError();
    int flag = 0;
    if (rc == cudaSuccess) rc = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(st);
    cudaFreeAsync(d_flag, st);                 // on every path: the flag is this call's own allocation
    AT2_CUDA_CHECK(rc);
    at2_count_launch(1);
    AT2_REQUIRE(flag == 0, "%s: a segment is longer than max_seg_len = %d", who, max_seg_len);
    return AT2_OK;
}

}  // namespace

extern "C" int at2_segmented_topk_f32(const float* d_losses, const int64_t* d_seg_offsets, const int32_t* d_keep,
                                      int64_t n_segments, int32_t max_seg_len, int32_t descending, int32_t* d_idx_out,
                                      void* stream) {
    return segmented_topk<float>(d_losses, d_seg_offsets, d_keep, n_segments, max_seg_len, descending, d_idx_out, stream,
                                 "at2_segmented_topk_f32");
}

extern "C" int at2_segmented_topk_f64(const double* d_losses, const int64_t* d_seg_offsets, const int32_t* d_keep,
                                      int64_t n_segments, int32_t max_seg_len, int32_t descending, int32_t* d_idx_out,
                                      void* stream) {
    return segmented_topk<double>(d_losses, d_seg_offsets, d_keep, n_segments, max_seg_len, descending, d_idx_out, stream,
                                  "at2_segmented_topk_f64");
}
