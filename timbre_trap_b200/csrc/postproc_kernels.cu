// Evaluation post-processing that follows `transcribe` in the reference, on the device:
//   filter_non_peaks   timbre_trap/utils/processing.py:66-98   keep strict local maxima along the frequency axis (zero-padded edges)
//   threshold          timbre_trap/utils/processing.py:101-124 binarise with `>= t`
//   (both as used by PitchDataset.activations_to_multi_pitch, datasets/PitchDataset.py:309-349, plus the bin mask of
//   experiments/evaluate.py:48)
//   frame-wise multi-pitch matching counts (mir_eval.multipitch via utils/experiments.py:354-396): per frame the size of a maximum
//   matching between estimated and reference bins within a pitch tolerance, and the totals precision / recall are made of.
// Integer / byte work: results are bit-exact against oracle/postproc_ref.py.
#include <algorithm>

#include "../../include/timbre_trap_b200.h"
#include "tt_common.cuh"

namespace tt {

// activations (B, F, T) fp32: a thread owns one (b, t) column pair-wise over f, so that loads are coalesced along T
__global__ void __launch_bounds__(256) filter_non_peaks_kernel(const float* __restrict__ a, float* __restrict__ out, int F, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const size_t i = ((size_t)b * F + f) * T + t;
    const float v = a[i];
    const float lo = f > 0 ? a[i - T] : 0.f, hi = f + 1 < F ? a[i + T] : 0.f;
    out[i] = (v > lo && v > hi) ? v : 0.f;
}

__global__ void __launch_bounds__(256) peak_threshold_kernel(const float* __restrict__ a, unsigned char* __restrict__ out, int F, int T, float thr,
                                                             int peaks_only, int f_lo, int f_hi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const size_t i = ((size_t)b * F + f) * T + t;
    const float v = a[i];
    bool on = v >= thr && f >= f_lo && f < f_hi;
    if (on && peaks_only) {
        const float lo = f > 0 ? a[i - T] : 0.f, hi = f + 1 < F ? a[i + T] : 0.f;
        on = v > lo && v > hi;
    }
    out[i] = on ? 1 : 0;
}

// One thread per frame: two cursors walk the estimated and the reference bins in increasing order; in one dimension this greedy
// pairing is a maximum matching for a distance threshold.  counts[b] += (true positives, estimated, reference) as 64-bit integers.
__global__ void __launch_bounds__(128) multipitch_counts_kernel(const unsigned char* __restrict__ est, const unsigned char* __restrict__ ref, int F, int T,
                                                                int tol, unsigned long long* __restrict__ counts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    unsigned int tp = 0, ne = 0, nr = 0;
    if (t < T) {
        const unsigned char* e = est + (size_t)b * F * T + t;
        const unsigned char* r = ref + (size_t)b * F * T + t;
        int i = 0, j = 0;                       // next candidate bins
        auto next = [&](const unsigned char* p, int k) { while (k < F && !p[(size_t)k * T]) ++k; return k; };
        for (int f = 0; f < F; ++f) { ne += e[(size_t)f * T] != 0; nr += r[(size_t)f * T] != 0; }
        i = next(e, 0); j = next(r, 0);
        while (i < F && j < F) {
            const int d = i - j;
            if (d <= tol && d >= -tol) { ++tp; i = next(e, i + 1); j = next(r, j + 1); }
            else if (i < j) i = next(e, i + 1);
            else j = next(r, j + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o);
        ne += __shfl_xor_sync(0xffffffffu, ne, o);
        nr += __shfl_xor_sync(0xffffffffu, nr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counts + 3 * b, (unsigned long long)tp);
        atomicAdd(counts + 3 * b + 1, (unsigned long long)ne);
        atomicAdd(counts + 3 * b + 2, (unsigned long long)nr);
    }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_filter_non_peaks(const float* activations, float* out, int B, int F, int T, void* stream) {
    TT_REQUIRE(activations && out, "null argument");
    if (B <= 0 || F <= 0 || T <= 0) return TT_OK;
    TT_REQUIRE(F <= 65535 && B <= 65535, "filter_non_peaks: at most 65535 bins / items");
    filter_non_peaks_kernel<<<dim3((T + 255) / 256, F, B), 256, 0, (cudaStream_t)stream>>>(activations, out, F, T);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_peak_threshold(const float* activations, unsigned char* out, int B, int F, int T, float threshold, int peaks_only, int bin_lo,
                                 int bin_hi, void* stream) {
    TT_REQUIRE(activations && out, "null argument");
    if (B <= 0 || F <= 0 || T <= 0) return TT_OK;
    TT_REQUIRE(F <= 65535 && B <= 65535, "peak_threshold: at most 65535 bins / items");
    TT_REQUIRE(bin_lo >= 0 && bin_hi <= F && bin_lo <= bin_hi, "peak_threshold: bad bin mask [%d, %d)", bin_lo, bin_hi);
    peak_threshold_kernel<<<dim3((T + 255) / 256, F, B), 256, 0, (cudaStream_t)stream>>>(activations, out, F, T, threshold, peaks_only, bin_lo, bin_hi);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_multipitch_counts(const unsigned char* est, const unsigned char* ref, int B, int F, int T, int tolerance_bins, int64_t* counts,
                                    void* stream) {
    TT_REQUIRE(est && ref && counts, "null argument");
    TT_REQUIRE(tolerance_bins >= 0, "negative tolerance");
    if (B <= 0) return TT_OK;
    TT_REQUIRE(B <= 65535, "multipitch_counts: at most 65535 items");
    TT_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)B * 3 * sizeof(int64_t), (cudaStream_t)stream));
    if (F <= 0 || T <= 0) return TT_OK;
    multipitch_counts_kernel<<<dim3((T + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(est, ref, F, T, tolerance_bins, (unsigned long long*)counts);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}
