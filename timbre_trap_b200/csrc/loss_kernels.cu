// Objectives of timbre_trap/framework/objectives.py as deterministic two-stage reductions.
//   reconstruction / consistency (objectives.py:11-33, 77-104):  sum((a-b)^2) / (B*T)        [sum over C,F; mean over B,T]
//   transcription (objectives.py:36-74):  per frame, bins whose target == 1 are weighted by (F - pos) / (pos + eps),
//                                         pos = sum_f target (weights that come out 0 become 1);  sum / (B*T)
// Stage 1 writes one partial per CTA, stage 2 adds them in a fixed order in fp64: results are bit-reproducible.
#include "../../include/timbre_trap_b200.h"
#include "tt_common.cuh"

namespace tt {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < kLossThreads / 32; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(kLossThreads) sq_diff_partial_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                                        long long n4, const float* __restrict__ a_tail,
                                                                        const float* __restrict__ b_tail, int tail,
                                                                        float* __restrict__ partial) {
    __shared__ float red[kLossThreads / 32];
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = ld_stream(a + i), y = ld_stream(b + i);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
        const float d = a_tail[threadIdx.x] - b_tail[threadIdx.x];
        acc += d * d;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// one thread per frame (b, t); consecutive threads = consecutive t (coalesced over the (B, F, T) layout)
__global__ void __launch_bounds__(kLossThreads) transcription_partial_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                                                                              int B, int F, int T, int weighted,
                                                                              float* __restrict__ partial) {
    __shared__ float red[kLossThreads / 32];
    float acc = 0.f;
    const long long frames = (long long)B * T;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += stride) {
        const long long b = i / T, t = i - b * T;
        const float* e = est + (size_t)b * F * T + t;
        const float* g = tgt + (size_t)b * F * T + t;
        float scale = 1.f;
        if (weighted) {
            float pos = 0.f;
            for (int f = 0; f < F; ++f) pos += g[(size_t)f * T];
            const float neg = (float)F - pos;
            scale = neg / (pos + 1.1920928955078125e-07f);   // torch.finfo().eps
            if (scale == 0.f) scale = 1.f;
        }
        float s = 0.f;
        for (int f = 0; f < F; ++f) {
            const float gv = g[(size_t)f * T];
            const float d = e[(size_t)f * T] - gv;
            s += (gv == 1.f ? scale : 1.f) * d * d;
        }
        acc += s;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void finish_kernel(const float* __restrict__ partial, int n, double scale, float* __restrict__ out) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)partial[i];
    *out = (float)(s * scale);
}

}  // namespace tt

using namespace tt;

extern "C" int tt_loss_scratch_floats(void) { return kLossMaxBlocks; }

extern "C" int tt_sum_sq_diff(const float* a, const float* b, int64_t n, double scale, float* out, float* scratch, void* stream_) {
    TT_REQUIRE(a && b && out && scratch, "null argument");
    TT_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0), "inputs must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n4 = n / 4;
    const int tail = (int)(n - 4 * n4);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n4 + kLossThreads - 1) / kLossThreads, kLossMaxBlocks));
    sq_diff_partial_kernel<<<blocks, kLossThreads, 0, stream>>>((const float4*)a, (const float4*)b, n4, a + 4 * n4, b + 4 * n4, tail, scratch);
    finish_kernel<<<1, 1, 0, stream>>>(scratch, blocks, scale, out);
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(2);
    return TT_OK;
}

extern "C" int tt_transcription_loss(const float* estimate, const float* target, int B, int F, int T, int weight_positive_class,
                                     float* out, float* scratch, void* stream_) {
    TT_REQUIRE(estimate && target && out && scratch, "null argument");
    TT_REQUIRE(B > 0 && F > 0 && T > 0, "empty input");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long frames = (long long)B * T;
    const int blocks = (int)std::min<long long>((frames + kLossThreads - 1) / kLossThreads, kLossMaxBlocks);
    transcription_partial_kernel<<<blocks, kLossThreads, 0, stream>>>(estimate, target, B, F, T, weight_positive_class, scratch);
    finish_kernel<<<1, 1, 0, stream>>>(scratch, blocks, 1.0 / (double)frames, out);
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(2);
    return TT_OK;
}
