// EncoderBlock.sconv and DecoderBlock.tconv (reference: timbre_trap/framework/modules.py:626-629, 685-688) as row-pipelined
// tcgen05 kernels, same machinery as the first design of the residual-block kernel (TMA row ring -> MMA issuer warps -> TMEM -> epilogue warp groups), one
// GEMM stage, bias folded into the GEMM, ELU in the epilogue:
//
//   DOWN  Conv2d(Cin, Cout, (4,1), stride (2,1)):       out[q] = ELU(b + sum_kh W[kh] x[2q + kh])          K = (kh, ci)
//   UP    ConvTranspose2d(Cin, Cout, (4,1), stride (2,1), output_padding) as a polyphase GEMM:
//         out[2q + r] = ELU(b + W[r+2] x[q-1] + W[r] x[q]),   N = (r, co),  K = (row q-1 | row q, ci);  rows outside the
//         input are zero (TMA out-of-bounds fill), which also produces the bias-only rows of the output padding.
//
// Every input row is fetched once per strip; with 8 channels per row (one core-matrix column) an MMA pairs the row with the
// constant ones operand (bias / zero weights), otherwise it pairs two channel groups of the row.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/timbre_trap_b200.h"
#include "strip_common.cuh"

namespace tt {

constexpr int kUdEpiWarps = 16;
constexpr int kUdIssuers = 4;
constexpr int kUdThreads = (kUdEpiWarps + 1 + kUdIssuers) * 32;
constexpr int kUdSlots = 8;
constexpr int kUdGroups = 4;

struct UpDownParams {
    __nv_bfloat16* y;
    const __nv_bfloat16* w;
    int B, Hin, Hout, T;
    int CGout;             // channel groups of one output row
    int groups;            // row groups in total (DOWN: Hout, UP: ceil(Hout / 2))
    int groups_per_strip;
    int act;               // 1: ELU on the output (the model's layers); 0: linear (the same kernels as data-gradient convolutions)
};

template <int CGIN, int N>
struct UdSmem {
    static constexpr int kRing = CGIN == 1 ? 32 : (CGIN == 2 ? 16 : (CGIN == 4 ? 12 : 8));
    static constexpr int kSlotBytes = CGIN * kStripTileT * 16;
    static constexpr int kW = 1024;
    __host__ __device__ static constexpr int kg(int rows) { return CGIN == 1 ? 2 * rows : rows * CGIN + 2; }
    __host__ __device__ static constexpr int ring_base(int rows) { return (kW + kg(rows) * N * 16 + 127) / 128 * 128; }
    __host__ __device__ static constexpr int ones_off(int rows) { return ring_base(rows) + kRing * kSlotBytes; }
    __host__ __device__ static constexpr int total(int rows) { return ones_off(rows) + 4096; }
};

// P4 (packed 4-channel layout on one side): DOWN reads it - a GEMM row is a frame PAIR, the 16 output columns are (frame parity, 8
// channels) and go to two consecutive frames of the C8 planar output; UP writes it - only the first 4 channels of each output row.
template <int CGIN, int N, int NCOL, bool UP, bool P4>
__global__ void __launch_bounds__(kUdThreads, (N <= 32 && !UP) ? 2 : 1) updown_strip_kernel(const __grid_constant__ CUtensorMap tmap_x, const UpDownParams p) {
    using S = UdSmem<CGIN, N>;
    constexpr int kRing = S::kRing;
    constexpr int R = UP ? 2 : 4;                 // input rows per output row group
    constexpr int STEP = UP ? 1 : 2;              // ring rows consumed per group
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* ring_full = bars;                   // [kRing]
    uint64_t* ring_free = bars + kRing;           // [kRing]  2 commits (each input row feeds two row groups)
    uint64_t* acc_full = bars + 2 * kRing;        // [kUdSlots]
    uint64_t* acc_free = acc_full + kUdSlots;     // [kUdSlots] 128 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 960);
    uint8_t* sW = smem + S::kW;
    uint8_t* sRing = smem + S::ring_base(R);
    uint8_t* sOnes = smem + S::ones_off(R);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * kStripTileT;
    const int g_start = blockIdx.y * p.groups_per_strip;
    const int n_out = min(p.groups, g_start + p.groups_per_strip) - g_start;
    const int b = blockIdx.z;
    const int first_row = UP ? g_start - 1 : 2 * g_start;       // input row held by ring index 0
    const int n_rows = UP ? n_out + 1 : 2 * n_out + 2;
    constexpr uint32_t ncols = kUdSlots * N < 32 ? 32 : kUdSlots * N;

    if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
    if (tid == 32) {
        for (int i = 0; i < kRing; ++i) {
            umma::mbar_init(&ring_full[i], 1);
            umma::mbar_init(&ring_free[i], 2);
        }
        for (int i = 0; i < kUdSlots; ++i) {
            umma::mbar_init(&acc_full[i], 1);
            umma::mbar_init(&acc_free[i], 128);
        }
        umma::mbar_fence_init();
    }
    for (int i = tid; i < S::kg(R) * N; i += kUdThreads) reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(p.w) + i);
    for (int i = tid; i < 256; i += kUdThreads)
        reinterpret_cast<uint4*>(sOnes)[i] = i < 128 ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == kUdEpiWarps) {
        // ---- producer: one TMA box per input row ----
        if (lane == 0) {
            for (int idx = 0; idx < n_rows; ++idx) {
                const int slot = idx % kRing;
                if (idx >= kRing) umma::mbar_wait(&ring_free[slot], (uint32_t)((idx / kRing - 1) & 1));
                mbar_expect_tx(&ring_full[slot], (uint32_t)S::kSlotBytes);
                tma_load_5d(sRing + (size_t)slot * S::kSlotBytes, &tmap_x, &ring_full[slot], 0, t0, first_row + idx, 0, b);
            }
        }
    } else if (warp > kUdEpiWarps) {
        // ---- MMA issuers (row group it -> issuer it % kUdIssuers) ----
        const int par = warp - (kUdEpiWarps + 1);
        const bool issuer = lane == 0;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint32_t ring0 = umma::smem_u32(sRing), ones0 = umma::smem_u32(sOnes), w0 = umma::smem_u32(sW);
        const uint32_t b_lo0 = desc_lo(w0, N * 16u), b_step = (2u * N * 16u) >> 4;
        if (par == 0 && issuer) {
            // ring rows whose first user would be the (non-existent) row group -1
            for (int k = 0; k < R; ++k) {
                const int r = -STEP + k;
                if (r >= 0) mbar_arrive(&ring_free[r % kRing]);
            }
        }
        for (int it = par; it < n_out; it += kUdIssuers) {
            const int u = it / kUdSlots, a = it % kUdSlots;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int idx = STEP * it + k;
                umma::mbar_wait(&ring_full[idx % kRing], (uint32_t)((idx / kRing) & 1));
            }
            if (u > 0) umma::mbar_wait(&acc_free[a], (uint32_t)((u - 1) & 1));
            umma::fence_after_sync();
            const uint32_t acc = tmem + (uint32_t)(a * N);
            uint32_t b_lo = b_lo0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t row = ring0 + (uint32_t)((STEP * it + k) % kRing) * S::kSlotBytes;
                if constexpr (CGIN == 1) {
                    if (issuer) umma::mma_bf16(acc, desc64(desc_lo(row, ones0 - row)), desc64(b_lo), idesc, k > 0);
                    b_lo += b_step;
                } else {
                    const uint32_t a_lo = desc_lo(row, 2048u);
#pragma unroll
                    for (int q = 0; q < CGIN / 2; ++q) {
                        if (issuer) umma::mma_bf16(acc, desc64(a_lo + (uint32_t)(2 * q) * (2048u >> 4)), desc64(b_lo), idesc, (k | q) != 0);
                        b_lo += b_step;
                    }
                }
            }
            if constexpr (CGIN > 1) {
                if (issuer) umma::mma_bf16(acc, desc64(desc_lo(ones0, 2048u)), desc64(b_lo), idesc, true);
            }
            if (issuer) {
                umma::commit(&acc_full[a]);
#pragma unroll
                for (int k = 0; k < R; ++k) umma::commit(&ring_free[(STEP * it + k) % kRing]);
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue groups: TMEM -> ELU -> bf16 -> global ----
        const int quad = warp & 3, g = warp >> 2;
        const int j = quad * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const bool t_ok = t0 + j < p.T;
        // The strided convs with N <= 32 run two CTAs per SM (TMEM: 8 slots x N columns each; 40 registers per thread, hence 8-column
        // TMEM loads): measured 0.69 -> 0.53 ms (4 -> 8 channels) and 0.68 -> 0.49 ms (8 -> 16); the transposed convs did not gain.
        constexpr int NV = (NCOL >= 16 && (N > 32 || UP)) ? 16 : 8;
        for (int it = g; it < n_out; it += kUdGroups) {
            const int u = it / kUdSlots, a = it % kUdSlots;
            const int grp = g_start + it;
            umma::mbar_wait(&acc_full[a], (uint32_t)(u & 1));
            umma::fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < NCOL; c0 += NV) {
                float v[NV];
                tmem_load<NV>(lane_addr + (uint32_t)(a * N + c0), v);
#pragma unroll
                for (int k = 0; k < NV; k += 8) {
                    const int c = c0 + k;
                    int ho, cg;
                    if constexpr (UP) { ho = 2 * grp + (c >= N / 2 ? 1 : 0); cg = (c % (N / 2)) >> 3; }
                    else { ho = grp; cg = c >> 3; }
                    uint4 o;
                    if (p.act) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[k + e] = elu_f(v[k + e]);
                    }
                    o.x = pack2(v[k], v[k + 1]);
                    o.y = pack2(v[k + 2], v[k + 3]);
                    o.z = pack2(v[k + 4], v[k + 5]);
                    o.w = pack2(v[k + 6], v[k + 7]);
                    if constexpr (P4 && !UP) {
                        // pair j -> frames 2 (t0 + j) + (c >> 3) of the 8-channel output (p.T counts pairs)
                        if (t_ok && ho < p.Hout)
                            reinterpret_cast<uint4*>(p.y)[((size_t)b * p.Hout + ho) * (2 * (size_t)p.T) + 2 * (size_t)(t0 + j) + (c >> 3)] = o;
                    } else if constexpr (P4 && UP) {
                        if (t_ok && ho < p.Hout)
                            reinterpret_cast<uint2*>(p.y)[((size_t)b * p.Hout + ho) * p.T + t0 + j] = make_uint2(o.x, o.y);
                    } else {
                        if (t_ok && cg < p.CGout && ho < p.Hout)
                            reinterpret_cast<uint4*>(p.y)[(((size_t)b * p.CGout + cg) * p.Hout + ho) * p.T + t0 + j] = o;
                    }
                }
            }
            umma::fence_before_sync();
            mbar_arrive(&acc_free[a]);
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

template <int CGIN, int N, int NCOL, bool UP, bool P4 = false>
static int launch_updown(const void* x, const UpDownParams& p, cudaStream_t stream) {
    using S = UdSmem<CGIN, N>;
    constexpr int R = UP ? 2 : 4;
    CUtensorMap map;
    const int rc = make_row_map(&map, x, p.B, CGIN, p.Hin, p.T, kStripTileT);
    if (rc) return rc;
    const int smem = S::total(R);
    static bool configured = false;
    if (!configured) {
        TT_CUDA_CHECK(cudaFuncSetAttribute(updown_strip_kernel<CGIN, N, NCOL, UP, P4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((p.T + kStripTileT - 1) / kStripTileT, (p.groups + p.groups_per_strip - 1) / p.groups_per_strip, p.B);
    updown_strip_kernel<CGIN, N, NCOL, UP, P4><<<grid, kUdThreads, smem, stream>>>(map, p);
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(1);
    return TT_OK;
}

static int strip_groups(int B, int T, int groups) {
    const long long tiles = (long long)B * ((T + kStripTileT - 1) / kStripTileT);
    int per = groups;
    const long long target = 6 * 2 * 148;
    if (tiles < target) {
        const int splits = (int)std::min<long long>((target + tiles - 1) / tiles, std::max(1, groups / 16));
        per = (groups + splits - 1) / splits;
    }
    if (strip_rows_override() > 0) per = strip_rows_override();
    return std::min(per, groups);
}

}  // namespace tt

using namespace tt;

extern "C" int tt_conv_down_strip(const void* x, void* y, const void* w, int B, int Cin, int Cout, int Hin, int T, int packed4_in,
                                  int act_elu, void* stream) {
    TT_REQUIRE(x && y && w, "null argument");
    TT_REQUIRE(!packed4_in || (Cin == 8 && Cout == 8 && T % 2 == 0), "packed input: 4 -> 8 channels only, even frame count");
    if (packed4_in) T /= 2;                 // the kernel works on frame pairs
    if (B <= 0 || T <= 0) return TT_OK;
    const int Hout = (Hin - 4) / 2 + 1;
    TT_REQUIRE(Hout >= 1, "conv_down: input too short");
    UpDownParams p;
    p.y = (__nv_bfloat16*)y; p.w = (const __nv_bfloat16*)w;
    p.B = B; p.Hin = Hin; p.Hout = Hout; p.T = T; p.CGout = Cout / 8; p.groups = Hout; p.act = act_elu;
    p.groups_per_strip = strip_groups(B, T, p.groups);
    cudaStream_t s = (cudaStream_t)stream;
    if (packed4_in) return launch_updown<1, 16, 16, false, true>(x, p, s);
    if (Cin == 8 && Cout == 8) return launch_updown<1, 16, 8, false>(x, p, s);
    if (Cin == 8 && Cout == 16) return launch_updown<1, 16, 16, false>(x, p, s);
    if (Cin == 16 && Cout == 32) return launch_updown<2, 32, 32, false>(x, p, s);
    if (Cin == 32 && Cout == 64) return launch_updown<4, 64, 64, false>(x, p, s);
    tt_set_error("conv_down_strip: unsupported channel pair %d -> %d", Cin, Cout);
    return TT_ERR_UNSUPPORTED;
}

extern "C" int tt_conv_up_strip(const void* x, void* y, const void* w, int B, int Cin, int Cout, int Hin, int out_pad, int T,
                                int packed4_out, int act_elu, void* stream) {
    TT_REQUIRE(x && y && w, "null argument");
    TT_REQUIRE(!packed4_out || (Cin == 8 && Cout == 8), "packed output: 8 -> 4 channels only");
    if (B <= 0 || T <= 0 || Hin <= 0) return TT_OK;
    UpDownParams p;
    p.y = (__nv_bfloat16*)y; p.w = (const __nv_bfloat16*)w;
    p.B = B; p.Hin = Hin; p.Hout = 2 * Hin + 2 + out_pad; p.T = T; p.CGout = Cout / 8; p.groups = (p.Hout + 1) / 2; p.act = act_elu;
    p.groups_per_strip = strip_groups(B, T, p.groups);
    cudaStream_t s = (cudaStream_t)stream;
    if (Cin == 64 && Cout == 32) return launch_updown<8, 64, 64, true>(x, p, s);
    if (Cin == 32 && Cout == 16) return launch_updown<4, 32, 32, true>(x, p, s);
    if (Cin == 16 && Cout == 8) return launch_updown<2, 16, 16, true>(x, p, s);
    if (packed4_out) return launch_updown<1, 16, 16, true, true>(x, p, s);
    if (Cin == 8 && Cout == 8) return launch_updown<1, 16, 16, true>(x, p, s);
    tt_set_error("conv_up_strip: unsupported channel pair %d -> %d", Cin, Cout);
    return TT_ERR_UNSUPPORTED;
}
