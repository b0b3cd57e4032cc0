// Fused ResidualConv2dBlock (reference: timbre_trap/framework/modules.py:721-777), row-stationary form:
//     y = x + ELU(W2 * ELU(W1 (*)_d x + b1) + b2)      (3x3 dilated 'same' conv, 1x1 conv, residual)
//
// An implicit GEMM with one MMA per tap (N = C; the first design of this kernel) re-reads every activation row nine times from
// shared memory, and below N ~ 100 an MMA costs the tensor pipe its A-operand fetch, not its math (profiles/r01_tcgen05_microbench.md).
// Here every INPUT row is multiplied once per horizontal tap against the weights of all three vertical taps at once:
//
//     D[128 rows][ (out row r-d | out row r | out row r+d) x C ]  +=  A[row r, shifted by kx*d][K = C] * B[kx][3C][K]
//
// i.e. N = 3C and a third of the operand reads.  The accumulators of the output rows live in TMEM as d rings of S slots
// (rows of equal residue mod d are neighbours in their ring, so the three targets of one input row are adjacent columns; at the
// ring's wrap-around the MMA is split in two).  Every MMA accumulates: a slot starts out holding the bias, written with
// tcgen05.st by the epilogue that drained it (so the biases are fp32 and cost no operand traffic).
//
//   warp 16 (producer)  one TMA box per input row (all channel groups, 128 + 2 halo GEMM rows, zero-filled outside the image) into
//                       a shared-memory ring; every input row is fetched once per strip
//   warp 17 (scout)     does the 3x3 issuer's barrier waiting (row landed, slot of the newest output row drained) and publishes one
//                       counter of cleared rows
//   warp 18 (3x3)       one thread: per input row the N = 3C MMAs, in row order (fixed accumulation order = bit-reproducible
//                       results; one commit per row covers all three contributions of output row r - d)
//   warp 19 (1x1)       one thread: the 1x1 conv of each output row from the bf16 intermediate in shared memory
//   warps 0-15          four epilogue groups (row -> group row % 4; warp quadrant = TMEM lane quadrant): accumulator -> ELU -> bf16
//                       intermediate;  accumulator -> ELU -> + x (from the ring) -> bf16 -> coalesced stores; both re-initialise
//                       their slot with the bias and release ring slots / accumulators with one arrival per warp
// All hand-offs are mbarriers; nothing in the row loop is a CTA-wide barrier.  For C <= 8 the rows are FOLDED (see the layout modes
// below) so that a GEMM row is always 16 values wide.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <type_traits>

#include "../../include/timbre_trap_b200.h"
#include "strip_common.cuh"

#ifdef TT_RS_PROFILE
#include <stdio.h>
#define TT_PROF(...) __VA_ARGS__
#else
#define TT_PROF(...)
#endif

namespace tt {

constexpr int kRsGroups = 4;
constexpr int kRsEpiWarps = 16;
// warps: 16 epilogue + TMA producer + scout + 3x3 issuer + 1x1 issuer (the last four use one thread each)
constexpr int kRsThreads = (kRsEpiWarps + 4) * 32;
constexpr int kRsSlots = 16;       // TMEM accumulator slots (3x3 rings + 1x1 slots)
#ifndef TT_RS_EPI_SLEEP
#define TT_RS_EPI_SLEEP 40
#endif
constexpr int kRsCmdSlots = 32;    // >= the deepest input-row ring: the scout can never be further ahead of the issuer than that

struct ResRsParams {
    __nv_bfloat16* y;
    __nv_bfloat16* mid;        // optional (training): the inner activation ELU(W1 * x + b1), in the layout of y
    const __nv_bfloat16* w1;   // packed (KG1, 3 NC, 8), see packing.pack_res_rs
    const __nv_bfloat16* w2;   // packed (KG2, NC, 8)
    const float* bias;         // (2, NC): accumulator-column biases of the 3x3 and the 1x1 conv
    int B, H, T;
    int rows_per_strip;
};

// Layout modes.  A GEMM row is normally one frame of a C8 planar tensor.  Small channel counts are FOLDED so that a GEMM row is
// 16 values wide whatever C is (every barrier hand-off, MMA and epilogue pass then covers 2-4x more frames, and those per-row
// costs are what bounds the kernel): the 3x3 kernel is Toeplitz-expanded over the frames of a row on the host.
constexpr int kRsPlanar = 0;   // C8 planar (B, CG, H, T, 8); taps are frame shifts of d
constexpr int kRsPairs8 = 1;   // packed (B, H, T, 4) seen as T/2 rows of 8 values (CG = 1); shifts of one pair
constexpr int kRsFold2 = 2;    // C <= 8: C8 planar with one channel group seen as T/2 rows of 16 values (2 frames x 8 channels)
constexpr int kRsFold4 = 4;    // C <= 4: packed (B, H, T, 4) seen as T/4 rows of 16 values (4 frames x 4 channels)

// compile-time plan of one instantiation: CG channel groups (of the GEMM row), dilation D, layout mode
template <int CG, int D, int MODE>
struct RsPlan {
    static constexpr bool P4 = MODE == kRsPairs8;
    static constexpr bool kFolded = MODE >= kRsFold2;
    static_assert(!kFolded || CG == 2, "folded rows are 16 values wide");
    static constexpr int kHalo = (P4 || MODE == kRsFold2) ? (D + 1) / 2 : (MODE == kRsFold4 ? 1 : D);   // column halo in GEMM rows (16-byte units)
    static constexpr int kColStep = (P4 || kFolded) ? 1 : D;         // distance between the horizontal taps / K groups of a row, 16-byte units
    static constexpr int kG = P4 ? 2 * kHalo + 1 : 3;                // K groups per row (CG = 1)
    // Folded rows: the frames a row's 3 horizontal taps touch are picked K GROUP by K GROUP (a 16-byte half of a GEMM row), not row
    // by row: MMA kx reads the first half of row i + s and the second half of row i + s - adj (adj = 1: LBO is one row short of the
    // plane distance).  fold 4, d <= 2 and fold 2, d = 1: frames -2..5 / -1..2 of the row = 2 MMAs, not 3 whole-row shifts;
    // fold 2, d = 3: frames {-3,-2, 0,1, 3,4} = 3 MMAs, not 5.
    static constexpr bool kTwoMma = (MODE == kRsFold4 && D <= 2) || (MODE == kRsFold2 && D == 1);
    static constexpr bool kSparse3 = MODE == kRsFold2 && D == 3;
    static constexpr int kShifts = kFolded ? (kTwoMma ? 2 : 3) : 3;  // MMAs per input row and channel-group pair (CG >= 2)
    static constexpr int fold_s(int kx) { return kTwoMma ? kx : (kSparse3 ? (kx == 2 ? 2 : kx - 1) : kx - 1); }
    static constexpr int fold_adj(int kx) { return kTwoMma ? 1 : (kSparse3 ? (kx == 1 ? 0 : 1) : 0); }
    // A-descriptor offset of MMA kx (low word: start address in 16-byte units, LBO in bits 16..29)
    static constexpr uint32_t a_off(int kx) {
        return kFolded ? (uint32_t)(fold_s(kx) + kHalo) - ((uint32_t)fold_adj(kx) << 16) : (uint32_t)(kx * kColStep);
    }
#ifndef TT_RS_RING2
#define TT_RS_RING2 17
#endif
#ifndef TT_RS_RING4
#define TT_RS_RING4 20
#endif
    // input-row ring: a row stays until the epilogue of ITS output row has read it as the residual (~4-7 row periods after its MMAs),
    // and the producer throttles on ring_free (measured: 280-320 cycles per row at 14 / 16 slots) - as deep as shared memory allows
    static constexpr int kRing = CG == 1 ? 32 : (CG == 2 ? TT_RS_RING2 : TT_RS_RING4);
    static constexpr int NC = CG >= 4 ? 32 : 16;                     // accumulator columns per output row (padded)
    static constexpr int N3 = 3 * NC;
    static constexpr int KG1 = CG == 1 ? kG + 1 : kShifts * CG;
    static constexpr int KG2 = CG == 1 ? 2 : CG;
    // accumulator slots: the 1x1 stage gets 8 (two per epilogue group, hiding its round trip) where the 3x3 rings still fit
    static constexpr int A2 = (CG == 4 || D == 3) ? 4 : 8;
    static constexpr int SR = (kRsSlots - A2 > 12 ? 12 : kRsSlots - A2) / D;   // slots per residue ring
    static constexpr int TW = kStripTileT + 2 * kHalo;
    static constexpr int kBars = 0;                                  // 2 kRing + 40 mbarriers (<= 832 bytes)
    static constexpr int kTmemSlot = 1008;
    static constexpr int kBias = 1024;
    static constexpr int kW1 = 1280;
    static constexpr int kW2 = kW1 + KG1 * N3 * 16;
    static constexpr int kMid = (kW2 + KG2 * NC * 16 + 127) / 128 * 128;
    static constexpr int kMidSlot = CG * 2048;
    static constexpr int kRingBase = kMid + A2 * kMidSlot;
    static constexpr int kSlotBytes = (CG * TW * 16 + 127) / 128 * 128;
    static constexpr int kZero = kRingBase + kRing * kSlotBytes;
    static constexpr int kCmd = kZero + 2048;                        // per-row issue commands, written by the scout (kRsCmdSlots x 32 bytes)
    static constexpr int kTotal = kCmd + kRsCmdSlots * 32;
    static_assert(SR >= 3 && D * SR + A2 <= kRsSlots, "TMEM slot plan");
    static_assert(kRing <= kRsCmdSlots, "one command slot per ring slot at least");
    static_assert(kTotal <= (CG <= 2 ? 115712 : 232448), "shared memory: two CTAs per SM for CG <= 2, one for CG = 4");
};

template <int NV>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const float* v) {
    if constexpr (NV == 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else if constexpr (NV == 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                     "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    } else {
        static_assert(NV == 16, "tmem_store: 4, 8 or 16 columns");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                     "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                     "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
    }
}
__device__ __forceinline__ void tmem_store_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int CG, int NREAL, int D, int MODE, bool MID>
__global__ void __launch_bounds__(kRsThreads, CG <= 2 ? 2 : 1) res_rs_kernel(const __grid_constant__ CUtensorMap tmap_x, const ResRsParams p) {
    using S_ = RsPlan<CG, D, MODE>;
    constexpr int NC = S_::NC, N3 = S_::N3, SR = S_::SR, A2 = S_::A2;
    constexpr int kRing = S_::kRing, TW = S_::TW, halo = S_::kHalo;
    constexpr int slot_bytes = S_::kSlotBytes;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_::kBars);
    uint64_t* ring_full = bars;                 // [kRing]  TMA landed
    uint64_t* ring_free = bars + kRing;         // [kRing]  8 arrivals: the 4 warps that saw the row's 3x3 MMAs complete + the 4 warps reading it as the residual
    uint64_t* acc1_full = bars + 2 * kRing;     // [12]     commit after the last contributing input row
    uint64_t* acc1_free = acc1_full + 12;       // [12]     4 arrivals (slot drained and re-initialised with the bias)
    uint64_t* mid_full = acc1_free + 12;        // [8]      4 arrivals
    uint64_t* acc2_full = mid_full + 8;         // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S_::kTmemSlot);
    uint32_t* rows_ready = tmem_slot + 1;       // number of input rows the scout has cleared for the 3x3 issuer
    float* sBias = reinterpret_cast<float*>(smem + S_::kBias);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sW1 = smem + S_::kW1;
    uint8_t* sW2 = smem + S_::kW2;
    uint8_t* sMid = smem + S_::kMid;
    uint8_t* sRing = smem + S_::kRingBase;
    uint8_t* sZero = smem + S_::kZero;

    const int t0 = blockIdx.x * kStripTileT;
    const int h_start = blockIdx.y * p.rows_per_strip;
    const int h_end = min(p.H, h_start + p.rows_per_strip);
    const int n_out = h_end - h_start;
    const int b = blockIdx.z;
    // input rows of the strip, relative to h_start: ri_first .. ri_last (rows outside the image contribute nothing and are skipped)
    const int ri_first = max(-D, -h_start);
    const int ri_last = min(n_out + D - 1, p.H - 1 - h_start);
    constexpr uint32_t ncols = kRsSlots * NC;
    constexpr int acc2_col0 = (kRsSlots - A2) * NC;
    // accumulator slot of output row it: ring (it % D), position (it / D) % SR; its use number (barrier phase) is (it / D) / SR
    auto slot_of = [](int it) { return (it % D) * SR + (it / D) % SR; };

    // ---- one-time setup ---------------------------------------------------------------------------------------
    if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
    if (tid == 32) {
        for (int i = 0; i < kRing; ++i) {
            umma::mbar_init(&ring_full[i], 1);
            umma::mbar_init(&ring_free[i], 8);
        }
        for (int i = 0; i < 12; ++i) {
            umma::mbar_init(&acc1_full[i], 1);
            umma::mbar_init(&acc1_free[i], 4);
        }
        for (int i = 0; i < 8; ++i) {
            umma::mbar_init(&mid_full[i], 4);
            umma::mbar_init(&acc2_full[i], 1);
        }
        *rows_ready = 0;
        umma::mbar_fence_init();
    }
    for (int i = tid; i < S_::KG1 * N3; i += kRsThreads) reinterpret_cast<uint4*>(sW1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1) + i);
    for (int i = tid; i < S_::KG2 * NC; i += kRsThreads) reinterpret_cast<uint4*>(sW2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2) + i);
    for (int i = tid; i < 2 * NC; i += kRsThreads) sBias[i] = __ldg(p.bias + i);
    for (int i = tid; i < 128; i += kRsThreads) reinterpret_cast<uint4*>(sZero)[i] = make_uint4(0u, 0u, 0u, 0u);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    constexpr int NV = CG == 4 ? 16 : (NREAL >= 8 ? 8 : NREAL);       // columns per TMEM load / store (40 registers per thread when two CTAs share an SM)
    if (warp < kRsEpiWarps) {
        // every accumulator slot starts out holding its bias
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int s = warp >> 2; s < kRsSlots; s += kRsGroups) {
            const float* bsrc = sBias + (s < kRsSlots - A2 ? 0 : NC);
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += NV) tmem_store<NV>(lane_addr + (uint32_t)(s * NC + c0), bsrc + c0);
        }
        tmem_store_wait();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    if (warp == kRsEpiWarps) {
        // =================================== producer ===================================
        if (lane == 0) {
            constexpr uint32_t bytes = (uint32_t)CG * TW * 16u;
            // The producer also does ALL of the 3x3 issuer's per-row index arithmetic (ring positions, wrap-around split, descriptor words)
            // and leaves it as a 32-byte command in shared memory before it starts the row's load: the issuer is the pacing thread of the
            // kernel (~830 cycles per row, of which ~300 were this arithmetic) and now only reads the command and issues.  The producer is
            // never more than kRing rows ahead of the issuer (ring_free), so kRsCmdSlots >= kRing entries suffice; the command is ordered
            // before the issuer's read through arrive.expect_tx (release) -> the scout's barrier wait -> its release store of rows_ready.
            constexpr uint32_t idesc1 = umma::make_idesc_bf16(128, NC), idesc2 = umma::make_idesc_bf16(128, 2 * NC), idesc3 = umma::make_idesc_bf16(128, N3);
            const uint32_t ring0 = umma::smem_u32(sRing);
            const uint32_t b_base = desc_lo(umma::smem_u32(sW1), N3 * 16u);
            uint4* cmds = reinterpret_cast<uint4*>(smem + S_::kCmd);
            TT_PROF(long long t_plan = 0, t_free = 0, t_tma = 0, tp = clock64();)
            for (int ri = ri_first; ri <= ri_last; ++ri) {
                const int idx = ri - ri_first, slot = idx % kRing;
                // target blocks j = 0, 1, 2 <-> output rows ri - D, ri, ri + D (vertical taps ky = 2, 1, 0)
                const int j0 = ri >= D ? 0 : (ri >= 0 ? 1 : 2);
                const int j1 = ri + D < n_out ? 2 : (ri < n_out ? 1 : 0);
                const int res = (ri + D) % D;                            // residue class of the three targets
                const int q1 = (ri + D) / D - 1;                         // ring sequence number of output row ri
                const int pa = (q1 - 1 + j0) % SR;
                const int na = min(j1 - j0 + 1, SR - pa), nb = (j1 - j0 + 1) - na;
                uint4 c0, c1;
                c0.x = ring0 + (uint32_t)slot * slot_bytes;                                           // the input row in the ring
                c0.y = tmem + (uint32_t)((res * SR + pa) * NC);                                       // first run of target slots
                c0.z = na <= 0 ? 0u : (na == 3 ? idesc3 : (na == 2 ? idesc2 : idesc1));               // (0 = nothing to issue)
                c0.w = b_base + (uint32_t)(j0 * NC);
                c1.x = tmem + (uint32_t)((res * SR) * NC);                                            // the run after the ring's wrap-around
                c1.y = nb <= 0 ? 0u : (nb == 3 ? idesc3 : (nb == 2 ? idesc2 : idesc1));
                c1.z = b_base + (uint32_t)((j0 + na) * NC);
                c1.w = ri - D >= 0 ? umma::smem_u32(&acc1_full[slot_of(ri - D)]) : 0u;                // barrier of the row this one completes
                TT_PROF(t_plan += clock64() - tp; tp = clock64();)
                if (idx >= kRing) umma::mbar_wait(&ring_free[slot], (uint32_t)((idx / kRing - 1) & 1));
                TT_PROF(t_free += clock64() - tp; tp = clock64();)
                cmds[2 * (idx % kRsCmdSlots)] = c0;
                cmds[2 * (idx % kRsCmdSlots) + 1] = c1;
                mbar_expect_tx(&ring_full[slot], bytes);
                tma_load_5d(sRing + (size_t)slot * slot_bytes, &tmap_x, &ring_full[slot], 0, t0 - halo, h_start + ri, 0, b);
                TT_PROF(t_tma += clock64() - tp; tp = clock64();)
            }
            TT_PROF(if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
                        const int nr = ri_last - ri_first + 1;
                        printf("rs producer: cycles/row: plan %lld, ring_free wait %lld, command + tma %lld\n", t_plan / nr, t_free / nr, t_tma / nr);
                    })
        }
    } else if (warp == kRsEpiWarps + 1) {
        // =================================== scout ===================================
        // Does the 3x3 issuer's waiting for it: row landed (TMA), accumulator slot of the row's newest target drained and re-initialised
        // (the other two targets were acquired with earlier rows).  A barrier probe costs ~100 cycles even when it succeeds; the issuer
        // only reads one counter.
        if (lane == 0) {
            TT_PROF(long long t_full = 0, t_acc = 0, tp = clock64();)
            for (int ri = ri_first; ri <= ri_last; ++ri) {
                const int idx = ri - ri_first;
                umma::mbar_wait(&ring_full[idx % kRing], (uint32_t)((idx / kRing) & 1));
                TT_PROF(t_full += clock64() - tp; tp = clock64();)
                const int it_new = ri + D;
                if (it_new < n_out) {
                    const int u = (it_new / D) / SR;
                    if (u > 0) umma::mbar_wait(&acc1_free[slot_of(it_new)], (uint32_t)((u - 1) & 1));
                }
                asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(umma::smem_u32(rows_ready)), "r"((uint32_t)idx + 1u) : "memory");
                TT_PROF(t_acc += clock64() - tp; tp = clock64();)
            }
            TT_PROF(if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
                        const int nr = ri_last - ri_first + 1;
                        printf("rs scout: cycles/row: row landed %lld, accumulator free + publish %lld\n", t_full / nr, t_acc / nr);
                    })
        }
    } else if (warp == kRsEpiWarps + 2) {
        // =================================== 3x3 MMA issuer: one thread, input rows in order ===================================
        // (a single in-order issuer keeps the accumulation order of every output row fixed - results are bit-reproducible - and needs
        // one commit per row: it covers everything issued before, i.e. all three contributions of output row ri - D)
        // (measured: running the loop warp-uniformly with one elected lane issuing is not faster - the cost is the tcgen05 hand-off)
        if (lane == 0) {
            constexpr bool issuer = true;
            const uint32_t zero0 = umma::smem_u32(sZero);
            constexpr uint32_t plane = (uint32_t)TW * 16u;
            constexpr uint32_t b_step = (2u * N3 * 16u) >> 4;             // two K groups per MMA
            const uint4* cmds = reinterpret_cast<const uint4*>(smem + S_::kCmd);
            uint32_t ready = 0;
            TT_PROF(long long t_wait = 0, t_mma = 0, t_commit = 0, tp = clock64();)
            for (int ri = ri_first; ri <= ri_last; ++ri) {
                const int idx = ri - ri_first;
                while (ready <= (uint32_t)idx)
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(ready) : "r"(umma::smem_u32(rows_ready)) : "memory");
                umma::fence_after_sync();
                const uint4 c0 = cmds[2 * (idx % kRsCmdSlots)], c1 = cmds[2 * (idx % kRsCmdSlots) + 1];
                TT_PROF(t_wait += clock64() - tp; tp = clock64();)
                const uint32_t row = c0.x;
                auto issue = [&](uint32_t acc, uint32_t idesc, uint32_t b_lo) {
                    if constexpr (CG == 1) {
                        // kG K groups at column offsets g * kColStep, consumed as pairs (g, g+1); the unpaired last group meets a zero
                        // operand (never an arbitrary neighbour: stale shared memory times zero could be NaN)
                        constexpr uint32_t cs = (uint32_t)S_::kColStep * 16u;
#pragma unroll
                        for (int g = 0; g < S_::kG; g += 2) {
                            const uint32_t a = row + (uint32_t)g * cs;
                            const uint32_t lbo = g + 1 < S_::kG ? cs : zero0 - a;
                            if (issuer) umma::mma_bf16(acc, desc64(desc_lo(a, lbo)), desc64(b_lo), idesc, true);
                            b_lo += b_step;
                        }
                    } else {
                        const uint32_t row_lo = ((row >> 4) & 0x3FFFu) | (((plane >> 4) & 0x3FFFu) << 16);
#pragma unroll
                        for (int kx = 0; kx < S_::kShifts; ++kx) {
#pragma unroll
                            for (int q = 0; q < CG / 2; ++q) {
                                if (issuer) umma::mma_bf16(acc, desc64(row_lo + S_::a_off(kx) + (uint32_t)(2 * q) * (plane >> 4)), desc64(b_lo), idesc, true);
                                b_lo += b_step;
                            }
                        }
                    }
                };
                if (c0.z) issue(c0.y, c0.z, c0.w);
                if (c1.y) issue(c1.x, c1.y, c1.z);
                TT_PROF(t_mma += clock64() - tp; tp = clock64();)
                if (c1.w) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c1.w) : "memory");
                if (ri == ri_last)
                    for (int it = max(0, ri_last - D + 1); it < n_out; ++it) umma::commit(&acc1_full[slot_of(it)]);
                TT_PROF(t_commit += clock64() - tp; tp = clock64();)
            }
            TT_PROF(if (issuer && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
                        const int nr = ri_last - ri_first + 1;
                        printf("rs issuer1: rows %d  cycles/row: wait %lld, mma %lld, commit %lld\n", nr, t_wait / nr, t_mma / nr, t_commit / nr);
                    })
        }
    } else if (warp == kRsEpiWarps + 3) {
        // =================================== 1x1 MMA issuer ===================================
        if (lane == 0) {
            constexpr bool issuer = true;
            constexpr uint32_t idesc = umma::make_idesc_bf16(128, NC);
            const uint32_t zero0 = umma::smem_u32(sZero), mid0 = umma::smem_u32(sMid), w2_0 = umma::smem_u32(sW2);
            const uint32_t b_lo0 = desc_lo(w2_0, NC * 16u);
            constexpr uint32_t b_step = (2u * NC * 16u) >> 4;
            TT_PROF(long long t_mid = 0, t_iss = 0, tp = clock64();)
            for (int it = 0; it < n_out; ++it) {
                const int u = it / A2, a = it % A2;
                umma::mbar_wait(&mid_full[a], (uint32_t)(u & 1));
                umma::fence_after_sync();
                TT_PROF(t_mid += clock64() - tp; tp = clock64();)
                const uint32_t acc = tmem + (uint32_t)(acc2_col0 + a * NC);
                const uint32_t mid = mid0 + (uint32_t)a * S_::kMidSlot;
                if constexpr (CG == 1) {
                    if (issuer) umma::mma_bf16(acc, desc64(desc_lo(mid, zero0 - mid)), desc64(b_lo0), idesc, true);
                } else {
                    const uint32_t a_lo = desc_lo(mid, 2048u);
#pragma unroll
                    for (int q = 0; q < CG / 2; ++q)
                        if (issuer) umma::mma_bf16(acc, desc64(a_lo + (uint32_t)(2 * q) * (2048u >> 4)), desc64(b_lo0 + (uint32_t)q * b_step), idesc, true);
                }
                umma::commit(&acc2_full[a]);
                TT_PROF(t_iss += clock64() - tp; tp = clock64();)
            }
            TT_PROF(if (issuer && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
                        printf("rs issuer2: cycles/row: mid wait %lld, issue %lld\n", t_mid / n_out, t_iss / n_out);)
        }
    } else if (warp < kRsEpiWarps) {
        // =================================== epilogue groups ===================================
        const int quad = warp & 3, g = warp >> 2;
        const int j = quad * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const bool t_ok = t0 + j < p.T;
        // planar: y[b][cg][h][t]; folded: the two 16-byte halves of a GEMM row are adjacent in memory, y[b][h][t][cg]
        uint4* const y_thread = S_::kFolded ? reinterpret_cast<uint4*>(p.y) + (((size_t)b * p.H + h_start) * p.T + t0 + j) * 2
                                            : reinterpret_cast<uint4*>(p.y) + ((size_t)b * CG * p.H + h_start) * p.T + t0 + j;

        TT_PROF(long long t_w1 = 0, t_e1 = 0, t_w2 = 0, t_e2 = 0, tp = clock64();)
        // ---- 3x3 accumulator -> ELU -> bf16 intermediate (A operand of the 1x1 conv); slot <- bias ----
        auto epi1 = [&](int it) {
            const int sl = slot_of(it), a = it % A2;
            TT_PROF(tp = clock64();)
            umma::mbar_wait_ns(&acc1_full[sl], (uint32_t)(((it / D) / SR) & 1), TT_RS_EPI_SLEEP);
            umma::fence_after_sync();
            TT_PROF(t_w1 += clock64() - tp; tp = clock64();)
            uint8_t* mid = sMid + (size_t)a * S_::kMidSlot + (size_t)j * 16u;
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += NV) {
                float v[NV];
                tmem_load<NV>(lane_addr + (uint32_t)(sl * NC + c0), v);
                tmem_store<NV>(lane_addr + (uint32_t)(sl * NC + c0), sBias + c0);
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
                    if constexpr (MID) if (t_ok) {
                        // the loss step keeps the inner activation (its backward needs it) instead of recomputing the 3x3 conv
                        uint4* gm = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.mid) +
                                                             (reinterpret_cast<const uint8_t*>(y_thread) - reinterpret_cast<const uint8_t*>(p.y)));
                        const int cg = (c0 + k) >> 3;
                        if constexpr (S_::kFolded) gm[(size_t)it * (2 * p.T) + cg] = o;
                        else gm[(size_t)cg * p.H * p.T + (size_t)it * p.T] = o;
                    }
                    if constexpr (NV < 8) break;
                }
            }
            tmem_store_wait();
            umma::fence_before_sync();
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&acc1_free[sl]);
                mbar_arrive(&mid_full[a]);
                // acc1_full[it] was committed after the MMAs of row it + D, in row order: the 3x3 stage is done with every row up to
                // it + D.  Release the ring slot of it + D - and of the rows that are no output row's "it + D" (the first D rows of the
                // strip); halo rows have no residual readers, so their release stands in for both halves of the count.
                if (it + D <= ri_last) mbar_arrive_n(&ring_free[(it + D - ri_first) % kRing], it + D >= n_out ? 2u : 1u);
                if (it < D) {
                    mbar_arrive(&ring_free[(it - ri_first) % kRing]);
                    if (it - D >= ri_first) mbar_arrive_n(&ring_free[(it - D - ri_first) % kRing], 2u);
                }
            }
            TT_PROF(t_e1 += clock64() - tp;)
        };
        // ---- 1x1 accumulator -> ELU -> + x -> bf16 -> global; slot <- bias ----
        auto epi2 = [&](int it) {
            const int u = it / A2, a = it % A2;
            TT_PROF(tp = clock64();)
            umma::mbar_wait_ns(&acc2_full[a], (uint32_t)(u & 1), TT_RS_EPI_SLEEP);
            umma::fence_after_sync();
            TT_PROF(t_w2 += clock64() - tp; tp = clock64();)
            const int ridx = it - ri_first;                            // ring index of row h
            const uint8_t* res = sRing + (size_t)(ridx % kRing) * slot_bytes + (size_t)(j + halo) * 16u;
            uint4 folded_out[2];                                       // folded rows: both halves of the thread's 32 contiguous bytes
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += NV) {
                float v[NV];
                tmem_load<NV>(lane_addr + (uint32_t)(acc2_col0 + a * NC + c0), v);
                tmem_store<NV>(lane_addr + (uint32_t)(acc2_col0 + a * NC + c0), sBias + NC + c0);
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
                    if constexpr (S_::kFolded) folded_out[cg] = o;
                    else if (t_ok) y_thread[(size_t)cg * p.H * p.T + (size_t)it * p.T] = o;
                    if constexpr (NV < 8) break;
                }
            }
            if constexpr (S_::kFolded) {
                // one 32-byte store per thread: a warp writes 1 KB contiguous
                if (t_ok)
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(y_thread + (size_t)it * (2 * p.T)), "r"(folded_out[0].x),
                                 "r"(folded_out[0].y), "r"(folded_out[0].z), "r"(folded_out[0].w), "r"(folded_out[1].x), "r"(folded_out[1].y),
                                 "r"(folded_out[1].z), "r"(folded_out[1].w) : "memory");
            }
            tmem_store_wait();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring_free[ridx % kRing]);
            TT_PROF(t_e2 += clock64() - tp;)
        };
        if constexpr (A2 >= 2 * kRsGroups) {
            // two slot sets per group: the 1x1 round trip of a row overlaps the first epilogue of the group's next row - unless that
            // row's accumulator is not there yet (measured: ~1000 cycles of waiting per row): then the pending second epilogue goes
            // first, which hands its input-row slot back to the producer and its accumulator to the 1x1 issuer that much earlier
            int prev = -1;
            for (int it = g; it < n_out; it += kRsGroups) {
#ifndef TT_RS_STATIC_ORDER
                if (prev >= 0) {
                    const bool there = umma::mbar_test(&acc1_full[slot_of(it)], (uint32_t)(((it / D) / SR) & 1));
                    if (!__all_sync(0xffffffffu, there)) {
                        epi2(prev);
                        prev = -1;
                    }
                }
#endif
                epi1(it);
                if (prev >= 0) epi2(prev);
                prev = it;
            }
            if (prev >= 0) epi2(prev);
        } else {
            for (int it = g; it < n_out; it += kRsGroups) {
                epi1(it);
                epi2(it);
            }
        }
        TT_PROF(if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
                    const int nr = (n_out + kRsGroups - 1) / kRsGroups;
                    printf("rs epilogue group 0: rows %d  cycles/row: acc1 wait %lld, epi1 %lld, acc2 wait %lld, epi2 %lld\n", nr, t_w1 / nr,
                           t_e1 / nr, t_w2 / nr, t_e2 / nr);
                })
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// folded view of a tensor whose GEMM rows are 32 contiguous bytes: 5-D tensor (8, Tf, H, 2, B) whose "channel group" dimension is
// the 16-byte half of the row, so that one TMA box (8, TW, 1, 2, 1) lands as two planes [half][row][8] - the planar operand layout
static inline int make_folded_row_map(CUtensorMap* map, const void* x, int B, int H, int Tf, int TW) {
    EncodeTiledFn fn = encode_fn();
    TT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[5] = {8, (cuuint64_t)Tf, (cuuint64_t)H, 2, (cuuint64_t)B};
    const cuuint64_t strides[4] = {32, (cuuint64_t)Tf * 32, 16, (cuuint64_t)H * Tf * 32};
    const cuuint32_t box[5] = {8, (cuuint32_t)TW, 1, 2, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (folded rows) failed with code %d", (int)r);
    return TT_OK;
}

template <int CG, int NREAL, int D, int MODE>
static int launch_rs(const void* x, ResRsParams p, cudaStream_t stream) {
    using S_ = RsPlan<CG, D, MODE>;
    static bool configured = false;
    if (!configured) {
        TT_CUDA_CHECK(cudaFuncSetAttribute(res_rs_kernel<CG, NREAL, D, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_::kTotal));
        TT_CUDA_CHECK(cudaFuncSetAttribute(res_rs_kernel<CG, NREAL, D, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_::kTotal));
        configured = true;
    }
    CUtensorMap map;
    const int rc = S_::kFolded ? make_folded_row_map(&map, x, p.B, p.H, p.T, S_::TW) : make_row_map(&map, x, p.B, CG, p.H, p.T, S_::TW);
    if (rc) return rc;
    dim3 grid((p.T + kStripTileT - 1) / kStripTileT, (p.H + p.rows_per_strip - 1) / p.rows_per_strip, p.B);
    // the variant that also writes the inner activation is a separate instantiation: the inference kernel keeps its register budget
    if (p.mid != nullptr) res_rs_kernel<CG, NREAL, D, MODE, true><<<grid, kRsThreads, S_::kTotal, stream>>>(map, p);
    else res_rs_kernel<CG, NREAL, D, MODE, false><<<grid, kRsThreads, S_::kTotal, stream>>>(map, p);
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(1);
    return TT_OK;
}

template <int CG, int NREAL, int MODE>
static int launch_rs_d(const void* x, const ResRsParams& p, int dilation, cudaStream_t stream) {
    if (dilation == 1) return launch_rs<CG, NREAL, 1, MODE>(x, p, stream);
    if (dilation == 2) return launch_rs<CG, NREAL, 2, MODE>(x, p, stream);
    return launch_rs<CG, NREAL, 3, MODE>(x, p, stream);
}

static std::atomic<int> g_strip_rows{0};
int strip_rows_override() { return g_strip_rows.load(std::memory_order_relaxed); }

}  // namespace tt

using namespace tt;

extern "C" int tt_set_strip_rows(int rows) {
    TT_REQUIRE(rows >= 0, "rows must be >= 0 (0 = automatic)");
    g_strip_rows.store(rows, std::memory_order_relaxed);
    return TT_OK;
}

extern "C" int tt_res_block_rs(const void* x, void* y, const void* w1, const void* w2, const float* bias, int B, int C, int c_real,
                               int H, int T, int dilation, int layout, void* stream) {
    return tt_res_block_rs_mid(x, y, nullptr, w1, w2, bias, B, C, c_real, H, T, dilation, layout, stream);
}

extern "C" int tt_res_block_rs_mid(const void* x, void* y, void* mid, const void* w1, const void* w2, const float* bias, int B, int C, int c_real,
                                   int H, int T, int dilation, int layout, void* stream) {
    TT_REQUIRE(x && y && w1 && w2 && bias, "null argument");
    TT_REQUIRE(C == 8 || C == 16 || C == 32, "res block: padded channel count must be 8, 16 or 32 (got %d)", C);
    TT_REQUIRE(layout == kRsPlanar || layout == kRsPairs8 || layout == kRsFold2 || layout == kRsFold4, "unknown layout mode %d", layout);
    TT_REQUIRE(layout == kRsPlanar || C == 8, "packed / folded layouts are for C <= 8");
    TT_REQUIRE((layout != kRsPairs8 && layout != kRsFold4) || c_real <= 4, "packed 4-channel layouts: at most 4 channels");
    TT_REQUIRE((layout != kRsPairs8 && layout != kRsFold2) || T % 2 == 0, "frame pairs need an even frame count");
    TT_REQUIRE(layout != kRsFold4 || T % 4 == 0, "frame quads need a frame count divisible by 4");
    TT_REQUIRE(dilation >= 1 && dilation <= 3, "dilation must be in [1,3]");
    TT_REQUIRE(c_real >= 1 && c_real <= C, "bad real channel count");
    if (B <= 0 || H <= 0 || T <= 0) return TT_OK;
    ResRsParams p;
    p.y = (__nv_bfloat16*)y; p.mid = (__nv_bfloat16*)mid; p.w1 = (const __nv_bfloat16*)w1; p.w2 = (const __nv_bfloat16*)w2; p.bias = bias;
    // the kernel's T counts GEMM rows: frames, frame pairs or frame quads
    if (layout == kRsPairs8 || layout == kRsFold2) T /= 2;
    if (layout == kRsFold4) T /= 4;
    p.B = B; p.H = H; p.T = T;
    // whole-height strips when the batch alone gives >= 6 waves of CTAs (2 CTAs/SM), shorter otherwise (each extra split re-reads
    // 2d halo rows but evens out the last wave)
    const long long tiles = (long long)B * ((T + kStripTileT - 1) / kStripTileT);
    int rows = H;
    const long long target = 6 * 2 * 148;
    if (tiles < target) {
        const int splits = (int)std::min<long long>((target + tiles - 1) / tiles, std::max(1, H / 32));
        rows = (H + splits - 1) / splits;
    }
    if (strip_rows_override() > 0) rows = strip_rows_override();
    p.rows_per_strip = std::min(rows, H);
    cudaStream_t s = (cudaStream_t)stream;
    if (layout == kRsFold4) return launch_rs_d<2, 16, kRsFold4>(x, p, dilation, s);
    if (layout == kRsFold2) return launch_rs_d<2, 16, kRsFold2>(x, p, dilation, s);
    if (layout == kRsPairs8) return launch_rs_d<1, 8, kRsPairs8>(x, p, dilation, s);
    if (C == 8) return c_real <= 4 ? launch_rs_d<1, 4, kRsPlanar>(x, p, dilation, s) : launch_rs_d<1, 8, kRsPlanar>(x, p, dilation, s);
    if (C == 16) return launch_rs_d<2, 16, kRsPlanar>(x, p, dilation, s);
    return launch_rs_d<4, 32, kRsPlanar>(x, p, dilation, s);
}
