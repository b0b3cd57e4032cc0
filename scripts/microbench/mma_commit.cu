// Microbenchmark: cost of tcgen05.commit next to tcgen05.mma (M = 128, N = 48, K = 16).  Each of `nw` warps loops over
// { n_mma MMAs into its own accumulator; n_commit commits to its own mbarriers } and we report SM cycles per loop iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I timbre_trap_b200/csrc scripts/microbench/mma_commit.cu -o variants/mma_commit
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace tt;

__global__ void __launch_bounds__(512) commit_kernel(int N, int nw, int n_mma, int n_commit, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16 * 4];
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (4096 + 8192) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) umma::tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 32) {
        for (int i = 0; i < 64; ++i) umma::mbar_init(&bars[i], 1);
        umma::mbar_init(&done, nw);
        umma::mbar_fence_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < nw) {
        const uint32_t a0 = umma::smem_u32(smem), b0 = a0 + 4096;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint64_t adesc = umma::make_desc(a0, 2048u, 128u), bdesc = umma::make_desc(b0, (uint32_t)N * 16u, 128u);
        const uint32_t acc = tmem + (uint32_t)((w * 64) % 448);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            for (int k = 0; k < n_mma; ++k) umma::mma_bf16(acc, adesc, bdesc, idesc, true);
            for (int k = 0; k < n_commit; ++k) umma::commit(&bars[w * 4 + k]);
        }
        umma::commit(&done);
        umma::mbar_wait(&done, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && w == 0) out[0] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + 8192);
    const int iters = 512;
    printf("%5s %6s %8s %16s\n", "warps", "mma", "commits", "cycles/iter (SM)");
    for (int nw : {1, 4}) {
        for (int n_mma : {1, 3}) {
            for (int n_commit : {0, 1, 2, 4}) {
                commit_kernel<<<148, 512, 4096 + 8192>>>(48, nw, n_mma, n_commit, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("%5d %6d %8d %16.1f\n", nw, n_mma, n_commit, (double)out[0] / iters / nw);
            }
        }
    }
    return 0;
}
