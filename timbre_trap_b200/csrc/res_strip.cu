// Fused ResidualConv2dBlock (reference: timbre_trap/framework/modules.py:721-777) as a warp-specialised, row-pipelined
// tcgen05 kernel:   y = x + ELU(W2 * ELU(W1 (*)_d x + b1) + b2)      (3x3 dilated 'same' conv, 1x1 conv, residual)
//
// One CTA owns a strip: 128 consecutive frames (T) x a run of rows (H) of one item, and walks down the rows:
//
//   warp 16 (producer)  one TMA box per input row (all channel groups, 128 + 2d frames, zero-filled outside the image)
//                       into a 16-slot shared-memory ring; every input row is fetched once per strip (no row halo re-reads)
//   warps 17-20 (3x3)   (row h -> issuer h % 4) for row h: the 3x3 taps are start-address offsets into the ring slots of rows h-d, h, h+d
//                       (implicit GEMM, M = 128 frames, N = C, K = 9 C) -> TMEM accumulator acc1[h % 4]
//   warps 21-22 (1x1)   the 1x1 conv of row h from the bf16 intermediate in shared memory -> acc2[h % 4].
//                       The biases ride along as one extra K group against a constant "ones" operand (bias split into
//                       bf16 hi + lo, so it is fp32-accurate), which removes the bias adds from the epilogues.
//   warps 0-15          four epilogue groups (row h -> group h % 4; warp quadrant = TMEM lane quadrant):
//                       acc1 -> ELU -> bf16 -> shared memory (A operand of the 1x1 conv);  acc2 -> ELU -> + x (from the
//                       ring slot of row h) -> bf16 -> coalesced 16 B stores.
//
// All hand-offs are mbarriers (ring_full / ring_free / acc1_full / acc1_free / mid_full / acc2_full); nothing in the row
// loop is a CTA-wide barrier.  HBM traffic is the algorithmic minimum (x read once, y written once) plus the 2d halo
// frames per row; the epilogues (one exp per conv output) are the co-limiter - see DESIGN.md.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "../../include/timbre_trap_b200.h"
#include "strip_common.cuh"

namespace tt {

constexpr int kGroups = 4;         // epilogue warp groups (row it -> group it % 4)
constexpr int kEpiWarps = 16;
// Warps issuing the 3x3 MMAs (row it -> issuer it % kIssuers1).  The waits, descriptor arithmetic and commits of one row cost
// ~1500 cycles of a single warp's time - far more than the tensor pipe needs for the MMAs themselves - so the issue work is
// spread over several warps; small-C layers (cheap rows, many of them) get more.
template <int CG> constexpr int issuers1() { return CG == 1 ? 8 : 4; }
constexpr int kIssuers2 = 2;       // warps issuing the 1x1 MMAs
template <int CG> constexpr int strip_threads() { return (kEpiWarps + 1 + issuers1<CG>() + kIssuers2) * 32; }

struct ResStripParams {
    __nv_bfloat16* y;
    const __nv_bfloat16* w1;   // packed, see packing.pack_res_strip
    const __nv_bfloat16* w2;
    int B, H, T;
    int d;                     // dilation (rows)
    int halo;                  // column halo in 16-byte units (= d; for the packed 4-channel layout, where a unit is a pixel PAIR, ceil((d+1)/2))
    int col_step;              // distance between the K groups of one tap row, in 16-byte units (= d; 1 for the packed layout)
    int groups_per_row;        // K groups per tap row (3; 3 or 5 for the packed layout), CG = 1 only
    int kg1;                   // K groups in the packed W1
    int rows_per_strip;
};

// shared-memory plan (bytes)
template <int CG>
struct StripSmem {
    // input-row ring slots: 2d+1 rows are pinned by the 3x3 window and ~5 by the epilogue lag; the rest is prefetch depth,
    // i.e. HBM bytes in flight per CTA (small-C rows are only 2-4 KB, so they get many more slots)
    static constexpr int kRing = CG == 1 ? 32 : (CG == 2 ? 14 : 16);
    // accumulator / intermediate slots (row it -> slot it % kSlots).  The kernel is latency-bound per row (MMA -> commit -> epilogue
    // -> MMA -> commit -> epilogue is ~3000 cycles), so throughput = rows in flight / latency: 8 slots where shared memory allows
    static constexpr int kSlots = CG == 4 ? 4 : 8;
    static constexpr int N = CG >= 4 ? 8 * CG : 16;                  // MMA N (padded)
    static constexpr int KG1 = CG == 1 ? 18 : 9 * CG + 2;            // K groups of W1 incl. the bias / padding groups (CG = 1: room for 3 x 6)
    static constexpr int KG2 = CG == 1 ? 2 : CG + 2;
    static constexpr int kBars = 0;                                  // 2 * kRing + 16 mbarriers (< 960 bytes)
    static constexpr int kTmemSlot = 960;
    static constexpr int kW1 = 1024;
    static constexpr int kW2 = kW1 + KG1 * N * 16;
    static constexpr int kMid = (kW2 + KG2 * N * 16 + 127) / 128 * 128;
    static constexpr int kMidSlot = CG * 2048;
    static constexpr int kRingBase = kMid + kSlots * kMidSlot;
    __host__ __device__ static constexpr int slot_bytes(int halo) { return (CG * (kStripTileT + 2 * halo) * 16 + 127) / 128 * 128; }
    __host__ __device__ static constexpr int ones_off(int d) { return kRingBase + kRing * slot_bytes(d); }
    __host__ __device__ static constexpr int total(int d) { return ones_off(d) + 4096; }
};

template <int CG, int NREAL>
__global__ void __launch_bounds__(strip_threads<CG>(), CG <= 2 ? 2 : 1) res_strip_kernel(const __grid_constant__ CUtensorMap tmap_x, const ResStripParams p) {
    using S = StripSmem<CG>;
    constexpr int N = S::N;
    constexpr int kRing = S::kRing;
    constexpr int kAcc = S::kSlots;
    constexpr int kIssuers1 = issuers1<CG>();
    constexpr int kStripThreads = strip_threads<CG>();
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
    uint64_t* ring_full = bars;                 // [kRing]  TMA landed
    uint64_t* ring_free = bars + kRing;         // [kRing]  3 commits (the 3x3 MMAs of the three rows using it) + 128 residual readers
    uint64_t* acc1_full = bars + 2 * kRing;     // [kAcc]
    uint64_t* acc1_free = acc1_full + kAcc;     // [kAcc]   128 arrivals
    uint64_t* mid_full = acc1_free + kAcc;      // [kAcc]   128 arrivals
    uint64_t* acc2_full = mid_full + kAcc;      // [kAcc]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = p.d, halo = p.halo;
    const int TW = kStripTileT + 2 * halo;
    const int slot_bytes = S::slot_bytes(halo);
    uint8_t* sW1 = smem + S::kW1;
    uint8_t* sW2 = smem + S::kW2;
    uint8_t* sMid = smem + S::kMid;
    uint8_t* sRing = smem + S::kRingBase;
    uint8_t* sOnes = smem + S::ones_off(halo);

    const int t0 = blockIdx.x * kStripTileT;
    const int h_start = blockIdx.y * p.rows_per_strip;
    const int h_end = min(p.H, h_start + p.rows_per_strip);
    const int b = blockIdx.z;
    const int first_row = h_start - d;                       // ring index 0
    constexpr uint32_t ncols = 2 * kAcc * N < 32 ? 32 : 2 * kAcc * N;

    // ---- one-time setup ---------------------------------------------------------------------------------------
    if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
    if (tid == 32) {
        for (int i = 0; i < kRing; ++i) {
            umma::mbar_init(&ring_full[i], 1);
            umma::mbar_init(&ring_free[i], 131);
        }
        for (int i = 0; i < kAcc; ++i) {
            umma::mbar_init(&acc1_full[i], 1);
            umma::mbar_init(&acc1_free[i], 128);
            umma::mbar_init(&mid_full[i], 128);
            umma::mbar_init(&acc2_full[i], 1);
        }
        umma::mbar_fence_init();
    }
    for (int i = tid; i < p.kg1 * N; i += kStripThreads) reinterpret_cast<uint4*>(sW1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1) + i);
    for (int i = tid; i < S::KG2 * N; i += kStripThreads) reinterpret_cast<uint4*>(sW2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2) + i);
    // "ones" operand: plane 0 rows = (1, 1, 0, ..., 0), plane 1 = zeros
    for (int i = tid; i < 256; i += kStripThreads)
        reinterpret_cast<uint4*>(sOnes)[i] = i < 128 ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == kEpiWarps) {
        // =================================== producer ===================================
        if (lane == 0) {
            const int n_rows = (h_end - h_start) + 2 * d;
            const uint32_t bytes = (uint32_t)CG * TW * 16u;
            for (int idx = 0; idx < n_rows; ++idx) {
                const int slot = idx % kRing;
                if (idx >= kRing) umma::mbar_wait(&ring_free[slot], (uint32_t)((idx / kRing - 1) & 1));
                mbar_expect_tx(&ring_full[slot], bytes);
                tma_load_5d(sRing + (size_t)slot * slot_bytes, &tmap_x, &ring_full[slot], 0, t0 - halo, first_row + idx, 0, b);
            }
        }
    } else if (warp > kEpiWarps && warp <= kEpiWarps + kIssuers1) {
        // =================================== 3x3 MMA issuers (rows it = par, par + kIssuers1, ...) ===================================
        // The whole warp runs the loop (warp-uniform control flow and descriptor arithmetic); one lane issues.
        const int par = warp - (kEpiWarps + 1);
        const bool issuer = lane == 0;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint32_t ring0 = umma::smem_u32(sRing), ones0 = umma::smem_u32(sOnes), w1_0 = umma::smem_u32(sW1);
        const uint32_t plane = (uint32_t)TW * 16u;
        const int n_out = h_end - h_start;
        const uint32_t b_lo0 = desc_lo(w1_0, N * 16u), b_step = (2u * N * 16u) >> 4;
        if (par == 0 && issuer) {
            // ring row r is released by 131 arrivals: the commits of the 3x3 MMAs of output rows r, r-d, r-2d (ring indices) and the
            // 128 residual readers of its own epilogue.  Rows near the top of the strip lack some of those users: stand in for them.
            for (int idx = 0; idx < d; ++idx) mbar_arrive_n(&ring_free[idx % kRing], 128);
            for (int v = -2 * d; v < 0; ++v)
                for (int k = 0; k < 3; ++k)
                    if (v + k * d >= 0) mbar_arrive(&ring_free[(v + k * d) % kRing]);
        }
        for (int it = par; it < n_out; it += kIssuers1) {
            const int u = it / kAcc, a = it % kAcc;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = it + k * d;
                umma::mbar_wait(&ring_full[idx % kRing], (uint32_t)((idx / kRing) & 1));
            }
            if (u > 0) umma::mbar_wait(&acc1_free[a], (uint32_t)((u - 1) & 1));
            umma::fence_after_sync();
            const uint32_t acc = tmem + (uint32_t)(a * N);
            uint32_t b_lo = b_lo0;
            if constexpr (CG == 1) {
                // per tap row: G K groups at column offsets g * col_step, consumed as pairs (g, g+1); an unpaired last group meets the
                // ones operand, which carries the bias for ky = 0 and zero weights otherwise (never an arbitrary neighbour: stale
                // shared memory times zero could be NaN)
                const uint32_t cs = (uint32_t)p.col_step * 16u;
                auto issue_rows = [&](auto g_tag) {
                    constexpr int G = decltype(g_tag)::value;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t row = ring0 + (uint32_t)((it + ky * d) % kRing) * slot_bytes;
#pragma unroll
                        for (int g = 0; g < G; g += 2) {
                            const uint32_t a = row + (uint32_t)g * cs;
                            const uint32_t lbo = g + 1 < G ? cs : ones0 - a;
                            if (issuer) umma::mma_bf16(acc, desc64(desc_lo(a, lbo)), desc64(b_lo), idesc, (ky | g) != 0);
                            b_lo += b_step;
                        }
                    }
                };
                if (p.groups_per_row == 3) issue_rows(std::integral_constant<int, 3>{});
                else issue_rows(std::integral_constant<int, 5>{});
            } else {
                const uint32_t a_lbo = ((plane >> 4) & 0x3FFFu) << 16;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t row_lo = (((ring0 + (uint32_t)((it + ky * d) % kRing) * slot_bytes) >> 4) & 0x3FFFu) | a_lbo;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                        for (int q = 0; q < CG / 2; ++q) {
                            if (issuer)
                                umma::mma_bf16(acc, desc64(row_lo + (uint32_t)(kx * d) + (uint32_t)(2 * q) * (plane >> 4)), desc64(b_lo), idesc,
                                               (ky | kx | q) != 0);
                            b_lo += b_step;
                        }
                    }
                }
                if (issuer) umma::mma_bf16(acc, desc64(desc_lo(ones0, 2048u)), desc64(b_lo), idesc, true);
            }
            if (issuer) {
                umma::commit(&acc1_full[a]);
#pragma unroll
                for (int k = 0; k < 3; ++k) umma::commit(&ring_free[(it + k * d) % kRing]);
            }
            __syncwarp();
        }
    } else if (warp > kEpiWarps + kIssuers1) {
        // =================================== 1x1 MMA issuers (rows it = par2, par2 + kIssuers2, ...) ===================================
        const int par2 = warp - (kEpiWarps + kIssuers1 + 1);
        const bool issuer = lane == 0;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint32_t ones0 = umma::smem_u32(sOnes), mid0 = umma::smem_u32(sMid), w2_0 = umma::smem_u32(sW2);
        const uint32_t b_lo0 = desc_lo(w2_0, N * 16u), b_step = (2u * N * 16u) >> 4;
        const int n_out = h_end - h_start;
        for (int it = par2; it < n_out; it += kIssuers2) {
            const int u = it / kAcc, a = it % kAcc;
            umma::mbar_wait(&mid_full[a], (uint32_t)(u & 1));
            umma::fence_after_sync();
            const uint32_t acc = tmem + (uint32_t)((kAcc + a) * N);
            const uint32_t mid = mid0 + (uint32_t)a * S::kMidSlot;
            if (issuer) {
                if constexpr (CG == 1) {
                    umma::mma_bf16(acc, desc64(desc_lo(mid, ones0 - mid)), desc64(b_lo0), idesc, false);
                } else {
                    const uint32_t a_lo = desc_lo(mid, 2048u);
#pragma unroll
                    for (int q = 0; q < CG / 2; ++q)
                        umma::mma_bf16(acc, desc64(a_lo + (uint32_t)(2 * q) * (2048u >> 4)), desc64(b_lo0 + (uint32_t)q * b_step), idesc, q > 0);
                    umma::mma_bf16(acc, desc64(desc_lo(ones0, 2048u)), desc64(b_lo0 + (uint32_t)(CG / 2) * b_step), idesc, true);
                }
                umma::commit(&acc2_full[a]);
            }
            __syncwarp();
        }
    } else {
        // =================================== epilogue groups ===================================
        const int quad = warp & 3, g = warp >> 2;
        const int j = quad * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const bool t_ok = t0 + j < p.T;
        const int n_out = h_end - h_start;
        constexpr int NV = NREAL >= 16 ? 16 : NREAL;                  // columns per TMEM load
        // ---- 3x3 accumulator -> ELU -> bf16 intermediate (A operand of the 1x1 conv) ----
        auto epi1 = [&](int it) {
            const int u = it / kAcc, a = it % kAcc;
            umma::mbar_wait(&acc1_full[a], (uint32_t)(u & 1));
            umma::fence_after_sync();
            uint8_t* mid = sMid + (size_t)a * S::kMidSlot + (size_t)j * 16u;
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += NV) {
                float v[NV];
                tmem_load<NV>(lane_addr + (uint32_t)(a * N + c0), v);
#pragma unroll
                for (int k = 0; k < NV; ++k) v[k] = elu_f(v[k]);
#pragma unroll
                for (int k = 0; k < NV; k += 8) {
                    uint4 o;
                    o.x = pack2(v[k], v[k + 1]);
                    o.y = pack2(v[k + 2], v[k + 3]);
                    if constexpr (NV >= 8) { o.z = pack2(v[k + 4], v[k + 5]); o.w = pack2(v[k + 6], v[k + 7]); }
                    else { o.z = 0u; o.w = 0u; }
                    *reinterpret_cast<uint4*>(mid + (size_t)((c0 + k) >> 3) * 2048u) = o;
                    if constexpr (NV < 8) break;
                }
            }
            umma::fence_before_sync();
            mbar_arrive(&acc1_free[a]);
            umma::fence_proxy_async();
            mbar_arrive(&mid_full[a]);
        };
        // ---- 1x1 accumulator -> ELU -> + x -> bf16 -> global ----
        auto epi2 = [&](int it) {
            const int u = it / kAcc, a = it % kAcc;
            const int h = h_start + it;
            umma::mbar_wait(&acc2_full[a], (uint32_t)(u & 1));
            umma::fence_after_sync();
            const int ridx = it + d;                                   // ring index of row h
            const uint8_t* res = sRing + (size_t)(ridx % kRing) * slot_bytes + (size_t)(j + halo) * 16u;
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += NV) {
                float v[NV];
                tmem_load<NV>(lane_addr + (uint32_t)((kAcc + a) * N + c0), v);
#pragma unroll
                for (int k = 0; k < NV; k += 8) {
                    const int cg = (c0 + k) >> 3;
                    const uint4 rx = *reinterpret_cast<const uint4*>(res + (size_t)cg * TW * 16u);
                    const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rx);
                    float r[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(rh[e]);
                        r[2 * e] = f.x;
                        r[2 * e + 1] = f.y;
                    }
                    constexpr int NE = NV >= 8 ? 8 : NV;
#pragma unroll
                    for (int e = 0; e < NE; ++e) r[e] += elu_f(v[k + e]);
                    uint4 o;
                    o.x = pack2(r[0], r[1]);
                    o.y = pack2(r[2], r[3]);
                    if constexpr (NE >= 8) { o.z = pack2(r[4], r[5]); o.w = pack2(r[6], r[7]); }
                    else { o.z = 0u; o.w = 0u; }
                    if (t_ok) reinterpret_cast<uint4*>(p.y)[(((size_t)b * CG + cg) * p.H + h) * p.T + t0 + j] = o;
                    if constexpr (NV < 8) break;
                }
            }
            umma::fence_before_sync();
            mbar_arrive(&ring_free[ridx % kRing]);
        };
        if constexpr (kAcc >= 2 * kGroups) {
            // two slot sets per group: the 1x1 round trip of row it overlaps the first epilogue of the group's next row
            int prev = -1;
            for (int it = g; it < n_out; it += kGroups) {
                epi1(it);
                if (prev >= 0) epi2(prev);
                prev = it;
            }
            if (prev >= 0) epi2(prev);
        } else {
            for (int it = g; it < n_out; it += kGroups) {
                epi1(it);
                epi2(it);
            }
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

template <int CG, int NREAL>
static int launch_strip(const CUtensorMap& map, const ResStripParams& p, cudaStream_t stream) {
    const int smem = StripSmem<CG>::total(p.halo);
    static int configured = 0;
    if (smem > configured) {
        TT_CUDA_CHECK(cudaFuncSetAttribute(res_strip_kernel<CG, NREAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    dim3 grid((p.T + kStripTileT - 1) / kStripTileT, (p.H + p.rows_per_strip - 1) / p.rows_per_strip, p.B);
    // optional thread-block clusters along T (TT_STRIP_CLUSTER = 2 / 4 / 8): co-schedules the CTAs that own adjacent 128-frame tiles
    // of the same rows.  Measured on B200: no gain (the kernel is bound by shared-memory operand bandwidth, not DRAM locality),
    // so the default is 1.
    static int cluster = -1;
    if (cluster < 0) { const char* e = getenv("TT_STRIP_CLUSTER"); cluster = e ? atoi(e) : 1; }
    int cx = 1;
    for (int c = cluster; c > 1; c >>= 1) if (grid.x % c == 0) { cx = c; break; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(strip_threads<CG>()); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cx; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, res_strip_kernel<CG, NREAL>, map, p));
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(1);
    return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_res_block_strip(const void* x, void* y, const void* w1, const void* w2, int B, int C, int c_real, int H, int T,
                                  int dilation, int packed4, void* stream) {
    TT_REQUIRE(x && y && w1 && w2, "null argument");
    TT_REQUIRE(C == 8 || C == 16 || C == 32, "res block: padded channel count must be 8, 16 or 32 (got %d)", C);
    TT_REQUIRE(!packed4 || (C == 8 && c_real <= 4 && T % 2 == 0), "packed layout: at most 4 channels and an even frame count");
    TT_REQUIRE(dilation >= 1 && dilation <= 3, "dilation must be in [1,3]");
    TT_REQUIRE(c_real >= 1 && c_real <= C, "bad real channel count");
    if (B <= 0 || H <= 0 || T <= 0) return TT_OK;
    ResStripParams p;
    p.y = (__nv_bfloat16*)y; p.w1 = (const __nv_bfloat16*)w1; p.w2 = (const __nv_bfloat16*)w2;
    // packed4: memory is (B, H, T, 4) bf16; a 16-byte unit is a PAIR of frames (e, 4 channels), so the kernel sees an 8-channel
    // tensor with T/2 "frames" whose taps are the pair offsets -halo..halo (Toeplitz-expanded weights, packing.pack_res_strip_pairs)
    if (packed4) T /= 2;
    p.B = B; p.H = H; p.T = T; p.d = dilation;
    p.halo = packed4 ? (dilation + 1) / 2 : dilation;
    p.col_step = packed4 ? 1 : dilation;
    p.groups_per_row = packed4 ? 2 * p.halo + 1 : 3;
    p.kg1 = C == 8 ? 3 * (p.groups_per_row + 1) : 9 * (C / 8) + 2;
    // whole-height strips when the batch alone fills the GPU, shorter ones otherwise (each strip re-reads 2d halo rows)
    // strips: whole-height when the batch alone gives >= 6 waves of CTAs (2 CTAs/SM), shorter otherwise (each extra split re-reads
    // 2d halo rows but evens out the last wave)
    const long long tiles = (long long)B * ((T + kStripTileT - 1) / kStripTileT);
    int rows = H;
    const long long target = 6 * 2 * 148;
    if (tiles < target) {
        const int splits = (int)std::min<long long>((target + tiles - 1) / tiles, std::max(1, H / 32));
        rows = (H + splits - 1) / splits;
    }
    const char* env = getenv("TT_STRIP_ROWS");
    if (env) rows = std::max(1, atoi(env));
    p.rows_per_strip = std::min(rows, H);
    CUtensorMap map;
    const int CG = C / 8;
    const int rc = make_row_map(&map, x, B, CG, H, T, kStripTileT + 2 * p.halo);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (C == 8) return (c_real <= 4 && !packed4) ? launch_strip<1, 4>(map, p, s) : launch_strip<1, 8>(map, p, s);
    if (C == 16) return launch_strip<2, 16>(map, p, s);
    return launch_strip<4, 32>(map, p, s);
}
