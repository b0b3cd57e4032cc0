// Self-test of the tcgen05 plumbing in umma.cuh: D (128 x N, fp32) = A (128 x K, bf16) * B (N x K, bf16)^T on one
// CTA, operands staged by plain loads into the SWIZZLE_NONE K-major canonical layout.  The GPU test suite runs it
// before any conv kernel so that a descriptor-encoding mistake shows up as a GEMM mismatch, not as a wrong model.
// TEST-ONLY: built into tests/csrc/libtt_selftest.so by timbre_trap_b200.build.build_selftest(); not part of the product ABI.
#include <stdarg.h>
#include <stdio.h>

#include "../../timbre_trap_b200/csrc/tt_common.cuh"
#include "../../timbre_trap_b200/csrc/umma.cuh"

// the product library's error / launch bookkeeping hooks, private to this test library
static thread_local char g_selftest_err[512] = "";
void tt_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_selftest_err, sizeof(g_selftest_err), fmt, ap);
    va_end(ap);
}
void tt_count_launches(int) {}
extern "C" const char* tt_selftest_last_error(void) { return g_selftest_err; }

namespace tt {

// smem layout: core matrix (row-group g, k-group kg) of an R-row operand at  kg * (R * 16) + g * 128  bytes
//   -> SBO (8-row groups) = 128 B, LBO (k-groups) = R * 16 B
__global__ void __launch_bounds__(128) umma_probe_kernel(const __nv_bfloat16* __restrict__ A,
                                                         const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                                                         int N, int K, int swap_lbo_sbo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)(K / 8) * 128 * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 128 * (K / 8); i += 128) {
        const int row = i % 128, kg = i / 128;
        *reinterpret_cast<uint4*>(sA + (size_t)kg * 128 * 16 + row * 16) =
            *reinterpret_cast<const uint4*>(A + (size_t)row * K + kg * 8);
    }
    for (int i = tid; i < N * (K / 8); i += 128) {
        const int row = i % N, kg = i / N;
        *reinterpret_cast<uint4*>(sB + (size_t)kg * N * 16 + row * 16) =
            *reinterpret_cast<const uint4*>(Bm + (size_t)row * K + kg * 8);
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
        for (int k = 0; k < K / 16; ++k) {
            const uint32_t lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
            uint64_t da = swap_lbo_sbo ? umma::make_desc(a0 + k * 2 * lbo_a, sbo, lbo_a) : umma::make_desc(a0 + k * 2 * lbo_a, lbo_a, sbo);
            uint64_t db = swap_lbo_sbo ? umma::make_desc(b0 + k * 2 * lbo_b, sbo, lbo_b) : umma::make_desc(b0 + k * 2 * lbo_b, lbo_b, sbo);
            umma::mma_bf16(tmem, da, db, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// MN-major operands (the layout of C8 planar activations when the GEMM's K axis is the PIXEL axis - weight gradients):
// A[m][k] lives at  (m / 8) * (K * 16) + k * 16 + (m % 8) * 2  bytes, i.e. [m-group][k][8 m] - 16 contiguous bytes hold 8 consecutive
// M (or N) elements of one k; B alike.  The instruction descriptor carries a_major = b_major = 1 (bits 15 / 16); the operand
// descriptor is (group stride K*16, k-group stride 128) in the order the `order` flag selects, so the test finds out which field
// is which on the hardware.  k_shift16 starts the A operand k_shift16 * 16 bytes into its buffer (a convolution tap as an address
// offset along K): D = A[:, shift:shift+K'] * B[:, :K']^T.
__global__ void __launch_bounds__(128) umma_probe_mn_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm,
                                                            float* __restrict__ D, int N, int K, int order, int k_shift, int k_use) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)16 * K * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i % 128, k = i / 128;
        *reinterpret_cast<__nv_bfloat16*>(sA + (size_t)(m / 8) * K * 16 + (size_t)k * 16 + (m % 8) * 2) = A[(size_t)m * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i % N, k = i / N;
        *reinterpret_cast<__nv_bfloat16*>(sB + (size_t)(n / 8) * K * 16 + (size_t)k * 16 + (n % 8) * 2) = Bm[(size_t)n * K + k];
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
        const uint32_t a0 = umma::smem_u32(sA) + (uint32_t)k_shift * 16u, b0 = umma::smem_u32(sB);
        const uint32_t group = (uint32_t)K * 16u, kgroup = 128u;
        for (int k = 0; k < k_use / 16; ++k) {
            const uint32_t off = (uint32_t)k * 256u;
            const uint64_t da = order ? umma::make_desc(a0 + off, group, kgroup) : umma::make_desc(a0 + off, kgroup, group);
            const uint64_t db = order ? umma::make_desc(b0 + off, group, kgroup) : umma::make_desc(b0 + off, kgroup, group);
            umma::mma_bf16(tmem, da, db, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace tt

extern "C" int tt_umma_probe(const void* a_bf16, const void* b_bf16, float* d, int n, int k, int swap_lbo_sbo, void* stream) {
    TT_REQUIRE(a_bf16 && b_bf16 && d, "null argument");
    TT_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && k % 16 == 0 && k >= 16 && k <= 256, "probe supports N,K in [16,256], multiples of 16");
    const size_t smem = (size_t)(k / 8) * (128 + n) * 16;
    TT_CUDA_CHECK(cudaFuncSetAttribute(tt::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tt::umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, d, n, k,
                                                                   swap_lbo_sbo);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

// order 0: descriptor (LBO = k-group stride 128 B, SBO = M/N-group stride); order 1: the two exchanged
extern "C" int tt_umma_probe_mn(const void* a_bf16, const void* b_bf16, float* d, int n, int k, int order, int k_shift, int k_use, void* stream) {
    TT_REQUIRE(a_bf16 && b_bf16 && d, "null argument");
    TT_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && k % 16 == 0 && k >= 16 && k <= 512, "probe supports N in [16,256], K in [16,512]");
    TT_REQUIRE(k_use % 16 == 0 && k_shift >= 0 && k_shift + k_use <= k, "bad K window");
    const size_t smem = (size_t)(16 + n / 8) * k * 16;
    TT_REQUIRE(smem <= 200 * 1024, "probe operands too large");
    TT_CUDA_CHECK(cudaFuncSetAttribute(tt::umma_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tt::umma_probe_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, d, n, k, order,
                                                                      k_shift, k_use);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}
