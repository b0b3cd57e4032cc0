// Microbenchmark: cost of tcgen05.mma (M = 128, K = 16, bf16, SWIZZLE_NONE operands) as a function of N and of how many
// independent accumulators the issue stream rotates over (1 = one dependent chain).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I timbre_trap_b200/csrc scripts/microbench/mma_chain.cu -o variants/mma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace tt;

__global__ void __launch_bounds__(512) chain_kernel(int N, int nacc, int n_mma, int a_distinct, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // zero operands: 8 A tiles of 4 KB (128 rows x K16), B tile up to 256 rows x K16 = 8 KB
    for (int i = threadIdx.x; i < (8 * 4096 + 8192) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) umma::tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 32) { umma::mbar_init(&bar, nacc); umma::mbar_fence_init(); }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    // nacc = number of issuing warps here; warp w accumulates into its own column range (wrapping at 512 columns)
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nacc) {
        const int w = threadIdx.x >> 5;
        const uint32_t a0 = umma::smem_u32(smem), b0 = a0 + 8 * 4096;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint64_t bdesc = umma::make_desc(b0, (uint32_t)N * 16u, 128u);
        const uint32_t acc = tmem + (uint32_t)((w * N) % (512 - N + 1) / 16 * 16);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint64_t adesc = umma::make_desc(a0 + (a_distinct ? ((i + w) & 7) * 4096 : 0), 2048u, 128u);
            umma::mma_bf16(acc, adesc, bdesc, idesc, true);
        }
        umma::commit(&bar);
        long long t1 = clock64();
        umma::mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 + 8192);
    const int n_mma = 256;
    printf("%5s %5s %6s %14s %14s\n", "N", "warps", "ctas", "issue cyc/mma", "total cyc/mma");
    for (int ctas : {148}) {
        for (int N : {16, 48, 96, 256}) {
            for (int nacc : {1, 2, 4, 8, 16}) {
                for (int rep = 0; rep < 2; ++rep) {
                    chain_kernel<<<ctas, 512, 8 * 4096 + 8192>>>(N, nacc, n_mma, 1, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                printf("%5d %5d %6d %14.1f %14.1f\n", N, nacc, ctas, (double)out[0] / n_mma / nacc, (double)out[1] / n_mma / nacc);
            }
        }
    }
    return 0;
}
