// Microbenchmark / hazard check: several warps issue tcgen05.mma (M = 128, N, K = 16, all-ones bf16 operands) into the SAME TMEM
// accumulator concurrently.  If the tensor pipe orders (or correctly interlocks) accumulations from different issuing threads,
// every element ends up as warps * n_mma * 16 exactly; a lost update shows up as a smaller value.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I timbre_trap_b200/csrc scripts/microbench/mma_shared_acc.cu -o variants/mma_shared_acc
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace tt;

__global__ void __launch_bounds__(512) shared_acc_kernel(int N, int nw, int n_mma, float* result, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (4096 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;   // bf16 ones
    if (threadIdx.x < 32) umma::tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 32) { umma::mbar_init(&bar, nw); umma::mbar_fence_init(); }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    if (lane == 0 && w >= 4 && w < 4 + nw) {
        const uint32_t a0 = umma::smem_u32(smem), b0 = a0 + 4096;
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint64_t adesc = umma::make_desc(a0, 2048u, 128u), bdesc = umma::make_desc(b0, (uint32_t)N * 16u, 128u);
        // the very first MMA (warp 4) overwrites, everything else accumulates; warp 4 gets a head start through the named flag
        if (w == 4) {
            umma::mma_bf16(tmem, adesc, bdesc, idesc, false);
            for (int i = 1; i < n_mma; ++i) umma::mma_bf16(tmem, adesc, bdesc, idesc, true);
        } else {
            __nanosleep(2000);
            for (int i = 0; i < n_mma; ++i) umma::mma_bf16(tmem, adesc, bdesc, idesc, true);
        }
        umma::commit(&bar);
    }
    if (w < 4) {
        umma::mbar_wait(&bar, 0);
        umma::fence_after_sync();
        long long t1 = clock64();
        float v[8];
        umma::tmem_ld8(tmem + ((uint32_t)(w * 32) << 16), v);
        umma::tmem_ld_wait();
        float mn = v[0], mx = v[0];
        for (int i = 1; i < 8; ++i) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
        if (lane == 0) { result[(blockIdx.x * 4 + w) * 2] = mn; result[(blockIdx.x * 4 + w) * 2 + 1] = mx; }
        if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
    float* result; long long* cycles;
    const int ctas = 148;
    cudaMallocManaged(&result, ctas * 8 * sizeof(float));
    cudaMallocManaged(&cycles, 8);
    cudaFuncSetAttribute(shared_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + 8192);
    const int n_mma = 2048;
    for (int N : {16, 48, 96}) {
        for (int nw : {1, 2, 4, 8}) {
            shared_acc_kernel<<<ctas, 512, 4096 + 8192>>>(N, nw, n_mma, result, cycles);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            float mn = 1e30f, mx = -1e30f;
            for (int i = 0; i < ctas * 4; ++i) { mn = fminf(mn, result[2 * i]); mx = fmaxf(mx, result[2 * i + 1]); }
            printf("N %3d warps %d: expected %.0f, min %.0f max %.0f   %.1f cycles/mma\n", N, nw, (double)nw * n_mma * 16, mn, mx,
                   (double)(cycles[0] - (nw > 1 ? 0 : 0)) / (n_mma * nw));
        }
    }
    return 0;
}
