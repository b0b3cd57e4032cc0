// Audio front-end of the data path, on the device:
//   * AudioDataset.get_audio (reference: timbre_trap/datasets/AudioDataset.py:67-77): mono mix, band-limited polyphase resampling to
//     the model's sample rate (torchaudio.functional.resample: windowed-sinc FIR bank, one filter per output phase, applied with a
//     stride of orig/gcd input samples), infinity-norm normalise - one pass over the input, one over the output;
//   * PitchDataset.multi_pitch_to_activations (datasets/PitchDataset.py:233-307): frame-wise pitch lists -> (F, T) activation targets
//     (nearest bin, Gaussian blur along frequency, renormalise so that annotated cells are >= 1, clip to [0, 1]).
// HBM-bound byte / float streaming; nothing here is GEMM-shaped enough to leave the CUDA cores (K <= ~460 taps in fp32, which the
// 1e-6 parity against torchaudio needs).
#include <math.h>

#include <algorithm>

#include "../../include/timbre_trap_b200.h"
#include "tt_common.cuh"

namespace tt {

constexpr int kRsThreadsFe = 256;
constexpr int kFramesPerThread = 4;

// out[i * new_f + j] = sum_k kern[j][k] * mono[i * orig + k - width],  mono[n] = mean over channels (zero outside [0, N)).
// A CTA owns `frames` consecutive output frames i (frames * new_f outputs) and stages the mono-mixed input span once in shared memory.
__global__ void __launch_bounds__(kRsThreadsFe)
resample_mono_kernel(const float* __restrict__ audio, int C, long long N, const float* __restrict__ kern, int orig, int new_f, int width,
                     int K, int frames, float* __restrict__ out, long long n_out, unsigned int* __restrict__ peak_bits) {
    extern __shared__ float xs[];
    const long long i0 = (long long)blockIdx.x * frames;
    const int span = frames * orig + K;
    const float inv_c = 1.0f / (float)C;
    for (int s = threadIdx.x; s < span; s += kRsThreadsFe) {
        const long long n = i0 * orig + s - width;
        float v = 0.f;
        if (n >= 0 && n < N) {
            for (int c = 0; c < C; ++c) v += __ldg(audio + (size_t)c * N + n);
            v = C > 1 ? v * inv_c : v;
        }
        xs[s] = v;
    }
    __syncthreads();
    // work items: (frame group of kFramesPerThread frames, phase j); a thread reuses every filter tap for its frames
    const int groups = (frames + kFramesPerThread - 1) / kFramesPerThread;
    float peak = 0.f;
    for (int w = threadIdx.x; w < groups * new_f; w += kRsThreadsFe) {
        const int g = w / new_f, j = w - g * new_f;
        const float* kr = kern + (size_t)j * K;
        float acc[kFramesPerThread];
#pragma unroll
        for (int f = 0; f < kFramesPerThread; ++f) acc[f] = 0.f;
        const float* x0 = xs + (size_t)g * kFramesPerThread * orig;
        for (int k = 0; k < K; ++k) {
            const float c = __ldg(kr + k);
#pragma unroll
            for (int f = 0; f < kFramesPerThread; ++f)
                if (g * kFramesPerThread + f < frames) acc[f] = fmaf(c, x0[f * orig + k], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < kFramesPerThread; ++f) {
            const int fr = g * kFramesPerThread + f;
            const long long o = (i0 + fr) * new_f + j;
            if (fr < frames && o < n_out) {
                out[o] = acc[f];
                peak = fmaxf(peak, fabsf(acc[f]));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((threadIdx.x & 31) == 0 && peak > 0.f) atomicMax(peak_bits, __float_as_uint(peak));
}

// One thread per frame: nearest-bin scatter of the frame's pitches (a bit set), Gaussian blur along frequency, store of the frame's
// column (coalesced across the threads of a warp), running minimum of the blurred value at annotated cells.
constexpr int kMaxBinWords = 32;      // up to 1024 frequency bins

__global__ void rasterise_kernel(const double* __restrict__ pitches_hz, int T, int P, const double* __restrict__ midi_freqs, int F,
                                 const float* __restrict__ blur, int R, float* __restrict__ act, unsigned int* __restrict__ min_bits) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < T;                        // no early exit: every lane takes part in the warp reduction below
    unsigned int bits[kMaxBinWords];
#pragma unroll
    for (int i = 0; i < kMaxBinWords; ++i) bits[i] = 0u;
    const double lb = midi_freqs[0], ub = midi_freqs[F - 1];
    for (int p = 0; live && p < P; ++p) {
        const double hz = pitches_hz[(size_t)t * P + p];
        if (!(hz != 0.0)) continue;                                     // zeros are "no pitch" (PitchDataset.py:263)
        const double midi = 12.0 * (log2(hz) - log2(440.0)) + 69.0;     // librosa.hz_to_midi
        if (!(midi >= lb && midi <= ub)) continue;                      // out of range: dropped (PitchDataset.py:271-274)
        // scipy interp1d(kind='nearest'): index = number of midpoints strictly below the query (ties go to the lower bin)
        int lo = 0, hi = F - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (0.5 * (midi_freqs[mid] + midi_freqs[mid + 1]) < midi) lo = mid + 1;
            else hi = mid;
        }
        bits[lo >> 5] |= 1u << (lo & 31);
    }
    float mn = __int_as_float(0x7f800000);
    for (int f = 0; live && f < F; ++f) {
        float v = 0.f;
        for (int r = -R; r <= R; ++r) {
            const int q = f + r;
            if (q >= 0 && q < F && ((bits[q >> 5] >> (q & 31)) & 1u)) v += blur[r + R];
        }
        act[(size_t)f * T + t] = v;
        if ((bits[f >> 5] >> (f & 31)) & 1u) mn = fminf(mn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) atomicMin(min_bits, __float_as_uint(mn));   // positive floats order like their bit patterns
}

__global__ void rasterise_finish_kernel(float* __restrict__ act, long long n, const unsigned int* __restrict__ min_bits) {
    const float mn = __uint_as_float(*min_bits);
    if (!(mn < __int_as_float(0x7f800000))) return;                     // no valid annotation at all: the map stays zero (PitchDataset.py:284)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) act[i] = fminf(fmaxf(act[i] / mn, 0.f), 1.f);
}

// per-channel energy terms of the signal-to-distortion ratio (see tt_sdr_correlations)
__global__ void __launch_bounds__(256)
sdr_corr_kernel(const float* __restrict__ target, const float* __restrict__ preds, long long N, int lags, double* __restrict__ r0,
                double* __restrict__ b, double* __restrict__ norms) {
    // grid.x: chunks of the signal; every thread owns a set of lags l and accumulates sum_n t[n] t[n+l] and sum_n t[n] p[n+l] in fp64 over
    // the CTA's span, staged in shared memory
    extern __shared__ float sm[];
    const int span = 4096;
    float* st = sm;                      // target[n0 .. n0 + span + lags)
    float* sp = sm + span + lags;        // preds [n0 .. n0 + span + lags)
    const long long n0 = (long long)blockIdx.x * span;
    const long long item = blockIdx.y;
    const float* tg = target + item * N;
    const float* pr = preds + item * N;
    for (int i = threadIdx.x; i < span + lags; i += 256) {
        const long long n = n0 + i;
        st[i] = n < N ? tg[n] : 0.f;
        sp[i] = n < N ? pr[n] : 0.f;
    }
    __syncthreads();
    const int len = (int)std::min<long long>(span, N - n0);
    for (int l = threadIdx.x; l < lags; l += 256) {
        double a = 0.0, c = 0.0;
        for (int n = 0; n < len; ++n) {
            const double tv = (double)st[n];
            a += tv * (double)st[n + l];
            c += tv * (double)sp[n + l];
        }
        atomicAdd(r0 + item * lags + l, a);
        atomicAdd(b + item * lags + l, c);
    }
    if (threadIdx.x < 32) {
        double tt_ = 0.0, pp = 0.0;
        for (int n = threadIdx.x; n < len; n += 32) {
            tt_ += (double)st[n] * (double)st[n];
            pp += (double)sp[n] * (double)sp[n];
        }
        for (int o = 16; o > 0; o >>= 1) {
            tt_ += __shfl_xor_sync(0xffffffffu, tt_, o);
            pp += __shfl_xor_sync(0xffffffffu, pp, o);
        }
        if (threadIdx.x == 0) {
            atomicAdd(norms + item * 2, tt_);
            atomicAdd(norms + item * 2 + 1, pp);
        }
    }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_resample_mono(const float* audio, int channels, int64_t n_in, const float* kernel, int orig, int new_f, int width,
                                float* out, int64_t n_out, float* peak, void* stream_) {
    TT_REQUIRE(audio && kernel && out && peak, "null argument");
    TT_REQUIRE(channels >= 1 && orig >= 1 && new_f >= 1 && width >= 0, "bad resampling geometry");
    cudaStream_t stream = (cudaStream_t)stream_;
    TT_CUDA_CHECK(cudaMemsetAsync(peak, 0, sizeof(float), stream));
    if (n_in <= 0 || n_out <= 0) return TT_OK;
    const int K = 2 * width + orig;
    // frames per CTA: about 2048 outputs, input span (frames * orig + K floats) within 44 KB of shared memory
    long long frames = std::max(1, 2048 / new_f);
    while (frames > 1 && (frames * orig + K) * 4 > 44 * 1024) frames /= 2;
    TT_REQUIRE((frames * orig + K) * 4 <= 200 * 1024, "resampling ratio %d -> %d needs too much shared memory", orig, new_f);
    const size_t smem = (size_t)(frames * orig + K) * 4;
    if (smem > 48 * 1024) TT_CUDA_CHECK(cudaFuncSetAttribute(resample_mono_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_frames = (n_out + new_f - 1) / new_f;
    const long long ctas = (n_frames + frames - 1) / frames;
    TT_REQUIRE(ctas < (1ll << 31), "clip too long for one launch");
    resample_mono_kernel<<<(unsigned)ctas, kRsThreadsFe, smem, stream>>>(audio, channels, n_in, kernel, orig, new_f, width, K, (int)frames, out, n_out,
                                                                         (unsigned int*)peak);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_rasterise_pitches(const double* pitches_hz, int T, int P, const double* midi_freqs, int F, const float* blur, int R,
                                    float* activations, float* min_scratch, void* stream_) {
    TT_REQUIRE(pitches_hz && midi_freqs && blur && activations && min_scratch, "null argument");
    TT_REQUIRE(F >= 2 && F <= 32 * kMaxBinWords && R >= 0 && P >= 0, "rasterise: 2..%d bins", 32 * kMaxBinWords);
    if (T <= 0) return TT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const unsigned int inf_bits = 0x7f800000u;
    TT_CUDA_CHECK(cudaMemcpyAsync(min_scratch, &inf_bits, sizeof(float), cudaMemcpyHostToDevice, stream));
    rasterise_kernel<<<(T + 127) / 128, 128, 0, stream>>>(pitches_hz, T, P, midi_freqs, F, blur, R, activations, (unsigned int*)min_scratch);
    const long long n = (long long)F * T;
    rasterise_finish_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(activations, n, (const unsigned int*)min_scratch);
    tt_count_launches(2);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_sdr_correlations(const float* target, const float* preds, int batch, int64_t n, int lags, double* r0, double* b,
                                   double* norms, void* stream_) {
    TT_REQUIRE(target && preds && r0 && b && norms, "null argument");
    TT_REQUIRE(lags >= 1 && lags <= 2048, "sdr: 1..2048 lags");
    if (batch <= 0 || n <= 0) return TT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    TT_CUDA_CHECK(cudaMemsetAsync(r0, 0, (size_t)batch * lags * sizeof(double), stream));
    TT_CUDA_CHECK(cudaMemsetAsync(b, 0, (size_t)batch * lags * sizeof(double), stream));
    TT_CUDA_CHECK(cudaMemsetAsync(norms, 0, (size_t)batch * 2 * sizeof(double), stream));
    const size_t smem = (size_t)2 * (4096 + lags) * sizeof(float);
    if (smem > 48 * 1024) TT_CUDA_CHECK(cudaFuncSetAttribute(sdr_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n + 4095) / 4096), batch);
    sdr_corr_kernel<<<grid, 256, smem, stream>>>(target, preds, n, lags, r0, b, norms);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}
