// Pieces shared by the row-pipelined ("strip") conv kernels: TMA row loads, mbarrier arrive variants, TMEM loads, descriptor
// arithmetic, the ELU epilogue math, and the host-side tensor-map construction for C8 planar activations.
#pragma once

#include <cuda.h>

#include "tt_common.cuh"
#include "umma.cuh"

namespace tt {

constexpr int kStripTileT = 128;   // frames per strip = M of one tcgen05.mma

// rows (row groups) per strip forced by tt_set_strip_rows(); 0 = the automatic split (defined in res_rs.cu)
int strip_rows_override();

// ---- extra PTX: TMA tile load, mbarrier arrive variants ---------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(umma::smem_u32(smem_dst)), "l"(map), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ELU with the hardware exp2 directly (ex2.approx.ftz: one MUFU op, no denormal fix-up code)
__device__ __forceinline__ float elu_f(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
    return v > 0.f ? v : e - 1.f;
}

// descriptor pieces: the low word holds (address >> 4) in bits 0-13 and (LBO >> 4) in bits 16-29, the high word (SBO >> 4)
// in bits 0-13 and the version bit 14; shared-memory addresses stay below 2^18, so adding (delta >> 4) to a low word moves
// the start address without touching the LBO field
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);     // SBO = 128 B, version 1
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int NV>
__device__ __forceinline__ void tmem_load(uint32_t taddr, float (&v)[NV]) {
    uint32_t r[NV];
    if constexpr (NV == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    } else if constexpr (NV == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    } else {
        static_assert(NV == 16, "tmem_load: 4, 8 or 16 columns");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// C8 planar activations (B, CG, H, T, 8) bf16 as a 5-D tensor (8, T, H, CG, B); one box = one row: (8, TW, 1, CG, 1)
static inline int make_row_map(CUtensorMap* map, const void* x, int B, int CG, int H, int T, int TW) {
    EncodeTiledFn fn = encode_fn();
    TT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[5] = {8, (cuuint64_t)T, (cuuint64_t)H, (cuuint64_t)CG, (cuuint64_t)B};
    const cuuint64_t strides[4] = {16, (cuuint64_t)T * 16, (cuuint64_t)H * T * 16, (cuuint64_t)CG * H * T * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)TW, 1, (cuuint32_t)CG, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return TT_OK;
}


}  // namespace tt
