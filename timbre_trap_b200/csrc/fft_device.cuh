// FFT building blocks shared by the CQT analysis / synthesis kernels.
//
//  * dft_small<R,SIGN>      in-register R-point DFT for R in {2,3,4,5,7}
//  * stockham_pass<R>       one CTA-cooperative Stockham autosort pass over a batch of equal-length
//                           FFTs held in shared memory (mixed radix; natural order in and out)
//  * run_passes             the whole radix schedule, ping-ponging two shared buffers
//  * fft_reg_dif<N,SIGN>    fully unrolled radix-2 DIF on a register array (N = 2,4,...,32);
//                           results are left in bit-reversed positions (use brev<N>(i))
#pragma once

#include "tt_common.cuh"

namespace tt {

constexpr int kMaxPasses = 12;

struct FftSpec {
    int n;
    int n_passes;
    int radix[kMaxPasses];
    // exact division by multiplication for the per-butterfly index split (dividends < 2^20, divisors <= 4096):
    // x / d == __umulhi(x, magic) with magic = floor(2^32 / d) + 1
    unsigned magic_per[kMaxPasses];    // d = n / radix   (butterflies per transform)
    unsigned magic_s[kMaxPasses];      // d = stride of the pass
};

// ---------------------------------------------------------------------------------------------
// small DFT constants: cos / sin of 2*pi*k/R, folded at compile time after unrolling
// ---------------------------------------------------------------------------------------------
template <int R> __device__ __forceinline__ float dft_cos(int k);
template <int R> __device__ __forceinline__ float dft_sin(int k);

template <> __device__ __forceinline__ float dft_cos<3>(int k) { return k == 0 ? 1.f : -0.5f; }
template <> __device__ __forceinline__ float dft_sin<3>(int k) {
    return k == 0 ? 0.f : (k == 1 ? 0.86602540378443865f : -0.86602540378443865f);
}
template <> __device__ __forceinline__ float dft_cos<5>(int k) {
    switch (k) {
        case 0: return 1.f;
        case 1: case 4: return 0.30901699437494742f;
        default: return -0.80901699437494742f;
    }
}
template <> __device__ __forceinline__ float dft_sin<5>(int k) {
    switch (k) {
        case 0: return 0.f;
        case 1: return 0.95105651629515357f;
        case 2: return 0.58778525229247313f;
        case 3: return -0.58778525229247313f;
        default: return -0.95105651629515357f;
    }
}
template <> __device__ __forceinline__ float dft_cos<7>(int k) {
    switch (k) {
        case 0: return 1.f;
        case 1: case 6: return 0.62348980185873353f;
        case 2: case 5: return -0.22252093395631440f;
        default: return -0.90096886790241913f;
    }
}
template <> __device__ __forceinline__ float dft_sin<7>(int k) {
    switch (k) {
        case 0: return 0.f;
        case 1: return 0.78183148246802981f;
        case 2: return 0.97492791218182361f;
        case 3: return 0.43388373911755812f;
        case 4: return -0.43388373911755812f;
        case 5: return -0.97492791218182361f;
        default: return -0.78183148246802981f;
    }
}

// y[u] = sum_v x[v] * exp(SIGN * 2 pi i u v / R), in place
template <int R, int SIGN>
__device__ __forceinline__ void dft_small(float2 (&v)[R]) {
    if constexpr (R == 2) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
        float2 a2 = cadd(v[1], v[3]), a3 = csub(v[1], v[3]);
        float2 r = SIGN > 0 ? cmul_i(a3) : cmul_mi(a3);
        v[0] = cadd(a0, a2);
        v[2] = csub(a0, a2);
        v[1] = cadd(a1, r);
        v[3] = csub(a1, r);
    } else {
        constexpr int H = (R - 1) / 2;
        float2 s[H], d[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            s[j] = cadd(v[j + 1], v[R - 1 - j]);
            d[j] = csub(v[j + 1], v[R - 1 - j]);
        }
        float2 x0 = v[0];
        float2 sum = x0;
#pragma unroll
        for (int j = 0; j < H; ++j) sum = cadd(sum, s[j]);
        v[0] = sum;
#pragma unroll
        for (int u = 1; u <= H; ++u) {
            float2 a = x0, b = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const int k = (u * (j + 1)) % R;
                const float c = dft_cos<R>(k), sn = dft_sin<R>(k);
                a.x = fmaf(c, s[j].x, a.x);
                a.y = fmaf(c, s[j].y, a.y);
                b.x = fmaf(sn, d[j].x, b.x);
                b.y = fmaf(sn, d[j].y, b.y);
            }
            float2 ib = SIGN > 0 ? cmul_i(b) : cmul_mi(b);
            v[u] = cadd(a, ib);
            v[R - u] = csub(a, ib);
        }
    }
}

// composite radices (6 = 2x3, 9 = 3x3, 10 = 2x5) as one in-register Cooley-Tukey step: fewer shared-memory passes
//   X[k1 + R1 k2] = sum_n2 W_R2^{n2 k2} ( W_R^{n2 k1} sum_n1 x[R2 n1 + n2] W_R1^{n1 k1} )
template <int R> __device__ __forceinline__ float2 comp_tw(int k);     // exp(+2 pi i k / R)
template <> __device__ __forceinline__ float2 comp_tw<6>(int k) {
    switch (k) { case 0: return make_float2(1.f, 0.f); case 1: return make_float2(0.5f, 0.86602540378443865f);
                 default: return make_float2(-0.5f, 0.86602540378443865f); }
}
template <> __device__ __forceinline__ float2 comp_tw<9>(int k) {
    switch (k) { case 0: return make_float2(1.f, 0.f); case 1: return make_float2(0.76604444311897804f, 0.64278760968653933f);
                 case 2: return make_float2(0.17364817766693035f, 0.98480775301220806f);
                 default: return make_float2(-0.93969262078590838f, 0.34202014332566873f); }   // k = 4
}
template <> __device__ __forceinline__ float2 comp_tw<10>(int k) {
    switch (k) { case 0: return make_float2(1.f, 0.f); case 1: return make_float2(0.80901699437494742f, 0.58778525229247313f);
                 case 2: return make_float2(0.30901699437494742f, 0.95105651629515357f);
                 case 3: return make_float2(-0.30901699437494742f, 0.95105651629515357f);
                 default: return make_float2(-0.80901699437494742f, 0.58778525229247313f); }  // k = 4
}

template <int R1, int R2, int SIGN>
__device__ __forceinline__ void dft_composite(float2 (&v)[R1 * R2]) {
    constexpr int R = R1 * R2;
    float2 a[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        float2 t[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
        dft_small<R1, SIGN>(t);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            if (n2 * k1 == 0) a[n2][k1] = t[k1];
            else {
                const float2 w = comp_tw<R>(n2 * k1);
                a[n2][k1] = SIGN > 0 ? cmul(t[k1], w) : cmulc(t[k1], w);
            }
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        float2 t[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) t[n2] = a[n2][k1];
        dft_small<R2, SIGN>(t);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
}

template <int R, int SIGN>
__device__ __forceinline__ void dft_any(float2 (&v)[R]) {
    if constexpr (R == 6) dft_composite<2, 3, SIGN>(v);
    else if constexpr (R == 9) dft_composite<3, 3, SIGN>(v);
    else if constexpr (R == 10) dft_composite<2, 5, SIGN>(v);
    else dft_small<R, SIGN>(v);
}

// ---------------------------------------------------------------------------------------------
// shared-memory Stockham autosort (decimation in frequency).  For the pass with remaining length
// n, stride s and m = n / R, butterfly (p, q), p < m, q < s:
//     y[q + s (R p + u)] = ( sum_v x[q + s (p + v m)] w_R^{u v} ) * w_n^{p u}
// `tw` holds exp(-2 pi i k / N) for k < N (forward sign); the inverse uses the conjugate.
// ---------------------------------------------------------------------------------------------
template <int R, int SIGN>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst, int N, int n,
                                              int s, const float2* __restrict__ tw, int n_fft, int tid, int n_threads,
                                              unsigned magic_per, unsigned magic_s) {
    const int m = n / R;
    const int per = N / R;
    const int total = n_fft * per;
    for (int w = tid; w < total; w += n_threads) {
        const int f = per == 1 ? w : (int)__umulhi((unsigned)w, magic_per);
        const int b = w - f * per;
        const int p = s == 1 ? b : (int)__umulhi((unsigned)b, magic_s);
        const int q = b - p * s;
        const float2* x = src + f * N;
        float2* y = dst + f * N;
        float2 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = x[q + s * (p + i * m)];
        dft_any<R, SIGN>(v);
        const int step = p * s;   // w_n^{p u} = w_N^{p u s}
        int ti = 0;
#pragma unroll
        for (int u = 1; u < R; ++u) {
            ti += step;
            if (ti >= N) ti -= N;
            float2 t = tw[ti];
            v[u] = SIGN > 0 ? cmulc(v[u], t) : cmul(v[u], t);
        }
#pragma unroll
        for (int u = 0; u < R; ++u) y[q + s * (R * p + u)] = v[u];
    }
}

// Runs every pass of `spec` over n_fft transforms of length spec.n stored back to back in `a`.
// Returns the buffer (a or b) that holds the result.  Ends with a __syncthreads().
template <int SIGN>
__device__ __forceinline__ float2* run_passes(const FftSpec& spec, float2* a, float2* b, const float2* __restrict__ tw,
                                              int n_fft, int tid, int n_threads) {
    int n = spec.n, s = 1;
    for (int i = 0; i < spec.n_passes; ++i) {
        const int r = spec.radix[i];
        const unsigned mp = spec.magic_per[i], ms = spec.magic_s[i];
        switch (r) {
            case 2: stockham_pass<2, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 3: stockham_pass<3, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 4: stockham_pass<4, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 5: stockham_pass<5, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 6: stockham_pass<6, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 7: stockham_pass<7, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            case 9: stockham_pass<9, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
            default: stockham_pass<10, SIGN>(a, b, spec.n, n, s, tw, n_fft, tid, n_threads, mp, ms); break;
        }
        n /= r;
        s *= r;
        __syncthreads();
        float2* t = a;
        a = b;
        b = t;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------
// register FFT (radix-2 DIF, bit-reversed output)
// ---------------------------------------------------------------------------------------------
__constant__ float2 c_tw32[32];   // exp(+2 pi i k / 32), filled at plan creation

template <int N>
__host__ __device__ constexpr int brev(int i) {
    int r = 0;
    for (int b = 1; b < N; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// v * exp(SIGN * 2 pi i k / 32) with the trivial cases folded (k is a compile-time constant after unrolling)
template <int SIGN>
__device__ __forceinline__ float2 mul_tw32(float2 v, int k) {
    k &= 31;
    if (k == 0) return v;
    if (k == 8) return SIGN > 0 ? cmul_i(v) : cmul_mi(v);
    if (k == 16) return make_float2(-v.x, -v.y);
    if (k == 24) return SIGN > 0 ? cmul_mi(v) : cmul_i(v);
    float2 t = c_tw32[k];
    return SIGN > 0 ? cmul(v, t) : cmulc(v, t);
}

template <int N, int SIGN>
__device__ __forceinline__ void fft_reg_dif(float2 (&v)[N]) {
#pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < N; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                float2 a = v[base + j], b = v[base + j + half];
                v[base + j] = cadd2(a, b);
                v[base + j + half] = mul_tw32<SIGN>(csub2(a, b), j * (16 / half));
            }
        }
    }
}

}  // namespace tt
