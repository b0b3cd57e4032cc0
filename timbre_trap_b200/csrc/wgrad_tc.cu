// Weight gradients of the 'same' convolutions of the residual blocks (reference: the autograd of timbre_trap/framework/modules.py:743-777
// inside the loss step, experiments/train.py:470-472) on the tensor cores.
//
//     dW[ky][kx][co][ci] = sum over (b, h, t) of  dZ[b, co, h, t] * X[b, ci, h + (ky-1) d, t + (kx-1) d]        (3x3 dilated, zero padding)
//     db[co]             = sum over (b, h, t) of  dZ[b, co, h, t]
//
// is a GEMM whose K axis is the PIXEL axis.  In the C8 planar activation layout (B, CG, H, T, 8) sixteen contiguous bytes hold 8
// channels of one pixel and consecutive pixels are 16 bytes apart: exactly the tcgen05 canonical layout of an MN-major operand
// (a core matrix = 8 k-rows of 16 bytes, SWIZZLE_NONE; LBO = 128 B between 8-pixel groups, SBO = plane stride between channel groups).
// So both operands are the rows the TMA brings in, untouched:
//     A = dZ row   (M = co, 128 lanes of which the first 8 CGo are real)             K = 16 pixels per tcgen05.mma
//     B = X row of the tap, started (kx-1) d pixels further along K (an address offset), N = ci
// and every tap owns NPAD fp32 accumulator columns in TMEM for the whole life of the CTA; a tenth "tap" multiplies dZ with a constant
// ones operand and yields db.  A CTA walks a strip of image rows of one 128-pixel column tile (X rows live in a ring, every row is
// fetched once per strip, out-of-image rows arrive as zeros from the TMA), then writes its accumulators to a partial buffer; a
// second kernel sums the partials in a fixed order (deterministic) into the PyTorch weight layout (co, ci, kh, kw).
#include <algorithm>

#include "../../include/timbre_trap_b200.h"
#include "strip_common.cuh"

namespace tt {

constexpr int kWgThreads = 128;        // warp 0: TMA producer + epilogue, warp 1: MMA issuer, warps 2-3: epilogue helpers (idle in the loop)
constexpr int kWgZSlots = 3;           // dZ row ring
constexpr int kWgMaxTaps = 10;         // 9 conv taps + the bias "tap"

struct WgradParams {
    float* partial;                    // (n_ctas, kWgMaxTaps, 32, NPAD) fp32
    int B, H, T;
    int CGi, CGo;
    int d;                             // dilation (= halo in pixels)
    int rows_per_strip;
};

// shared memory plan (bytes): [barriers 1 KB][ones 4 KB][dZ ring][X ring]; the A operand reads 16 channel groups = 32 KB from its slot, the
// B operand NPAD/8 groups: both stay inside the allocation because the X ring follows and the allocation is padded (see wgrad_smem)
template <int NPAD, int KS>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_same_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z,
                                                                   const WgradParams p) {
    constexpr int TAPS = KS * KS;
    constexpr uint32_t ncols = (TAPS + 1) * NPAD <= 32 ? 32 : ((TAPS + 1) * NPAD <= 64 ? 64 : ((TAPS + 1) * NPAD <= 128 ? 128 : ((TAPS + 1) * NPAD <= 256 ? 256 : 512)));
    extern __shared__ __align__(1024) uint8_t smem[];
    const int d = KS == 3 ? p.d : 0;
    const int TW = kStripTileT + 2 * d;
    const int xring = KS == 3 ? 2 * d + 3 : 2;
    const uint32_t z_slot = (uint32_t)p.CGo * kStripTileT * 16u;
    const uint32_t x_plane = (uint32_t)TW * 16u;
    const uint32_t x_slot = ((uint32_t)p.CGi * x_plane + 127u) & ~127u;
    uint64_t* z_full = reinterpret_cast<uint64_t*>(smem);            // [kWgZSlots]
    uint64_t* z_empty = z_full + kWgZSlots;                           // [kWgZSlots]
    uint64_t* x_full = z_empty + kWgZSlots;                           // [xring <= 9]
    uint64_t* x_empty = x_full + 16;                                  // [xring]
    uint64_t* done = x_empty + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    uint8_t* sOnes = smem + 1024;                                      // [2 groups][128 px][8] bf16 ones (group 1 only pads N to 16)
    uint8_t* sZ = smem + 1024 + 4096;
    uint8_t* sX = sZ + kWgZSlots * z_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * kStripTileT;
    const int h0 = blockIdx.y * p.rows_per_strip;
    const int h1 = min(p.H, h0 + p.rows_per_strip);
    const int b = blockIdx.z;
    const int n_rows = h1 - h0;                    // dZ rows of this strip
    const int n_xrows = n_rows + 2 * d;            // X rows h0 - d .. h1 - 1 + d

    if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
    if (tid == 32) {
        for (int i = 0; i < kWgZSlots; ++i) { umma::mbar_init(&z_full[i], 1); umma::mbar_init(&z_empty[i], 1); }
        for (int i = 0; i < xring; ++i) { umma::mbar_init(&x_full[i], 1); umma::mbar_init(&x_empty[i], 1); }
        umma::mbar_init(done, 1);
        umma::mbar_fence_init();
    }
    for (int i = tid; i < 2 * kStripTileT * 4; i += kWgThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;   // bf16 1.0 pairs
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ================= producer: X rows run d ahead of the dZ rows =================
        int xr = 0;                                                    // next X row (relative index 0 .. n_xrows-1 <-> image row h0 - d + xr)
        auto load_x = [&]() {
            const int slot = xr % xring;
            if (xr >= xring) umma::mbar_wait(&x_empty[slot], (uint32_t)((xr / xring - 1) & 1));
            mbar_expect_tx(&x_full[slot], (uint32_t)p.CGi * x_plane);
            tma_load_5d(sX + (size_t)slot * x_slot, &tmap_x, &x_full[slot], 0, t0 - d, h0 - d + xr, 0, b);
            ++xr;
        };
        for (int r = 0; r < n_rows; ++r) {
            while (xr < n_xrows && xr <= r + 2 * d + 1) load_x();      // rows r .. r + 2d (+1 prefetch) relative = image rows h0+r-d .. h0+r+d
            const int slot = r % kWgZSlots;
            if (r >= kWgZSlots) umma::mbar_wait(&z_empty[slot], (uint32_t)((r / kWgZSlots - 1) & 1));
            mbar_expect_tx(&z_full[slot], z_slot);
            tma_load_5d(sZ + (size_t)slot * z_slot, &tmap_z, &z_full[slot], 0, t0, h0 + r, 0, b);
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = umma::make_idesc_bf16(128, NPAD) | (1u << 15) | (1u << 16);     // both operands MN-major
        const uint32_t z0 = umma::smem_u32(sZ), x0 = umma::smem_u32(sX), ones0 = umma::smem_u32(sOnes);
        bool first = true;                                             // the very first MMA of every accumulator overwrites
        for (int r = 0; r < n_rows; ++r) {
            const int zs = r % kWgZSlots;
            umma::mbar_wait(&z_full[zs], (uint32_t)((r / kWgZSlots) & 1));
            // X rows r .. r + 2d (relative); the newest one (and, at the start, all of them) may still be in flight
            for (int xr = (r == 0 ? 0 : r + 2 * d); xr <= r + 2 * d; ++xr) umma::mbar_wait(&x_full[xr % xring], (uint32_t)((xr / xring) & 1));
            umma::fence_after_sync();
            const uint32_t za = z0 + (uint32_t)zs * z_slot;
#pragma unroll 1
            for (int s = 0; s < kStripTileT / 16; ++s) {
                const uint64_t da = umma::make_desc(za + (uint32_t)s * 256u, 128u, (uint32_t)kStripTileT * 16u);
#pragma unroll
                for (int ky = 0; ky < KS; ++ky) {
                    const int xr = r + ky * d;                         // relative X row of this vertical tap
                    const uint32_t xa = x0 + (uint32_t)(xr % xring) * x_slot;
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
                        const uint64_t db = umma::make_desc(xa + (uint32_t)(kx * d) * 16u + (uint32_t)s * 256u, 128u, x_plane);
                        umma::mma_bf16(tmem + (uint32_t)((ky * KS + kx) * NPAD), da, db, idesc, !(first && s == 0));
                    }
                }
                const uint64_t dones = umma::make_desc(ones0 + (uint32_t)s * 256u, 128u, (uint32_t)kStripTileT * 16u);
                umma::mma_bf16(tmem + (uint32_t)(TAPS * NPAD), da, dones, idesc, !(first && s == 0));
            }
            first = false;
            umma::commit(&z_empty[zs]);
            // X row r (relative) is not needed by later rows
            umma::commit(&x_empty[r % xring]);
        }
        umma::commit(done);
    }
    __syncwarp();
    // ================= epilogue: accumulators -> partial buffer (lanes 0..31 = output channels; NPAD columns per tap) =================
    if (warp == 0) {
        umma::mbar_wait_warp(done, 0);
        umma::fence_after_sync();
        const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        float* dst = p.partial + (cta * kWgMaxTaps * 32 + lane) * NPAD;
#pragma unroll 1
        for (int tap = 0; tap <= TAPS; ++tap) {
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                umma::tmem_ld16(tmem + (uint32_t)(tap * NPAD + c0), v);
                umma::tmem_ld_wait();
                float4* o = reinterpret_cast<float4*>(dst + (size_t)tap * 32 * NPAD + c0);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
                o[2] = make_float4(v[8], v[9], v[10], v[11]);
                o[3] = make_float4(v[12], v[13], v[14], v[15]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// dW (co, ci, KS, KS) and db (co) += fixed-order sums of the per-CTA partials
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int n_ctas, int npad, int ks, int co, int ci, float* __restrict__ dw,
                                    float* __restrict__ db) {
    const int taps = ks * ks;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;              // over (tap' in [0, taps], co, ci)
    const int total = (taps + 1) * co * ci;
    if (i >= total) return;
    const int tap = i / (co * ci), rem = i - tap * co * ci;
    const int o = rem / ci, c = rem - o * ci;
    if (tap == taps && c != 0) return;
    const float* src = partial + ((size_t)tap * 32 + o) * npad + (tap == taps ? 0 : c);
    const size_t stride = (size_t)kWgMaxTaps * 32 * npad;
    float acc = 0.f;
    for (int k = 0; k < n_ctas; ++k) acc += src[(size_t)k * stride];
    if (tap == taps) {
        if (db) db[o] += acc;
    } else {
        dw[((size_t)o * ci + c) * taps + tap] += acc;                 // (co, ci, ky, kx): tap = ky * ks + kx
    }
}

// ---- element-wise pieces of the backward pass on bf16 tensors of ANY common layout -------------------------------------------------
__device__ __forceinline__ float elu_grad_from_output(float a) { return a > 0.f ? 1.f : a + 1.f; }

// dz = gy * ELU'(z) through the activated output a
__global__ void elu_bwd_bf16_kernel(const uint4* __restrict__ gy, const uint4* __restrict__ a, uint4* __restrict__ dz, long long n8) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 g = __ldcs(gy + i), av = __ldcs(a + i);
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&av);
        uint4 r;
        uint32_t* rw = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fg = __bfloat1622float2(gh[k]), fa = __bfloat1622float2(ah[k]);
            rw[k] = pack2(fg.x * elu_grad_from_output(fa.x), fg.y * elu_grad_from_output(fa.y));
        }
        dz[i] = r;
    }
}

// residual block: dz2 = gy * ELU'(z2) with the activated 1x1 output recovered as y - x
__global__ void res_out_bwd_bf16_kernel(const uint4* __restrict__ gy, const uint4* __restrict__ y, const uint4* __restrict__ x, uint4* __restrict__ dz,
                                        long long n8) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 g = __ldcs(gy + i), yv = __ldcs(y + i), xv = __ldcs(x + i);
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
        const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&yv);
        const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xv);
        uint4 r;
        uint32_t* rw = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fg = __bfloat1622float2(gh[k]), fy = __bfloat1622float2(yh[k]), fx = __bfloat1622float2(xh[k]);
            rw[k] = pack2(fg.x * elu_grad_from_output(fy.x - fx.x), fg.y * elu_grad_from_output(fy.y - fx.y));
        }
        dz[i] = r;
    }
}

// packed 4-channel (B, H, T, 4) <-> C8 planar with one channel group (B, 1, H, T, 8): the first / last stage's tensors for the kernels
// that only know C8 planar
__global__ void p4_to_c8_kernel(const uint2* __restrict__ p4, uint4* __restrict__ c8, long long n_px) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
        const uint2 v = __ldcs(p4 + i);
        c8[i] = make_uint4(v.x, v.y, 0u, 0u);
    }
}
__global__ void c8_to_p4_kernel(const uint4* __restrict__ c8, uint2* __restrict__ p4, long long n_px) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
        const uint4 v = __ldcs(c8 + i);
        p4[i] = make_uint2(v.x, v.y);
    }
}

static size_t wgrad_smem(int CGi, int CGo, int d, int ks) {
    const int dd = ks == 3 ? d : 0;
    const size_t TW = kStripTileT + 2 * dd;
    const size_t z_slot = (size_t)CGo * kStripTileT * 16;
    const size_t x_slot = ((size_t)CGi * TW * 16 + 127) & ~(size_t)127;
    const size_t xring = ks == 3 ? 2 * dd + 3 : 2;
    size_t s = 1024 + 4096 + kWgZSlots * z_slot + xring * x_slot + 4 * TW * 16;   // + the padded N groups read past the last X slot
    // the A operand spans 16 channel groups (32 KB) from the start of a dZ slot, the ones operand 2 groups: keep both inside
    s = std::max(s, (size_t)1024 + 4096 + (kWgZSlots - 1) * z_slot + 16 * (size_t)kStripTileT * 16 + 1024);
    return s;
}

template <int NPAD, int KS>
static int launch_wgrad(const void* x, const void* dz, const WgradParams& p, dim3 grid, cudaStream_t stream) {
    const size_t smem = wgrad_smem(p.CGi, p.CGo, p.d, KS);
    TT_REQUIRE(smem <= 227 * 1024, "wgrad: %zu bytes of shared memory", smem);
    static size_t configured = 0;
    if (smem > configured) {
        TT_CUDA_CHECK(cudaFuncSetAttribute(wgrad_same_kernel<NPAD, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int dd = KS == 3 ? p.d : 0;
    CUtensorMap mx, mz;
    int rc = make_row_map(&mx, x, p.B, p.CGi, p.H, p.T, kStripTileT + 2 * dd);
    if (rc) return rc;
    rc = make_row_map(&mz, dz, p.B, p.CGo, p.H, p.T, kStripTileT);
    if (rc) return rc;
    wgrad_same_kernel<NPAD, KS><<<grid, kWgThreads, smem, stream>>>(mx, mz, p);
    TT_CUDA_CHECK(cudaGetLastError());
    tt_count_launches(1);
    return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int64_t tt_wgrad_scratch_floats(int B, int H, int T) {
    // upper bound over the strip splits tt_conv_wgrad_same chooses
    const long long tiles = (long long)B * ((T + kStripTileT - 1) / kStripTileT);
    const long long strips = std::max<long long>(1, std::min<long long>((2 * 148 + tiles - 1) / tiles, H));
    return tiles * (strips + 1) * kWgMaxTaps * 32 * 32;
}

extern "C" int tt_conv_wgrad_same(const void* x, const void* dz, float* dw, float* db, int B, int Cin, int Cout, int cin_real, int cout_real,
                                  int H, int T, int k, int dilation, float* scratch, void* stream_) {
    TT_REQUIRE(x && dz && dw && scratch, "null argument");
    TT_REQUIRE((Cin == 8 || Cin == 16 || Cin == 32) && (Cout == 8 || Cout == 16 || Cout == 32), "wgrad: padded channel counts must be 8, 16 or 32");
    TT_REQUIRE(cin_real >= 1 && cin_real <= Cin && cout_real >= 1 && cout_real <= Cout, "wgrad: bad real channel counts");
    TT_REQUIRE((k == 3 && dilation >= 1 && dilation <= 3) || k == 1, "wgrad: 3x3 (dilation 1..3) or 1x1");
    if (B <= 0 || H <= 0 || T <= 0) return TT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    WgradParams p;
    p.partial = scratch;
    p.B = B; p.H = H; p.T = T; p.CGi = Cin / 8; p.CGo = Cout / 8; p.d = dilation;
    const long long tiles = (long long)B * ((T + kStripTileT - 1) / kStripTileT);
    const long long strips = std::max<long long>(1, std::min<long long>((2 * 148 + tiles - 1) / tiles, H));
    p.rows_per_strip = (int)((H + strips - 1) / strips);
    dim3 grid((T + kStripTileT - 1) / kStripTileT, (H + p.rows_per_strip - 1) / p.rows_per_strip, B);
    const int n_ctas = (int)(grid.x * grid.y * grid.z);
    const int npad = Cin == 32 ? 32 : 16;
    int rc;
    if (k == 3) rc = npad == 32 ? launch_wgrad<32, 3>(x, dz, p, grid, stream) : launch_wgrad<16, 3>(x, dz, p, grid, stream);
    else rc = npad == 32 ? launch_wgrad<32, 1>(x, dz, p, grid, stream) : launch_wgrad<16, 1>(x, dz, p, grid, stream);
    if (rc) return rc;
    const int total = (k * k + 1) * cout_real * cin_real;
    wgrad_reduce_kernel<<<(total + 127) / 128, 128, 0, stream>>>(scratch, n_ctas, npad, k, cout_real, cin_real, dw, db);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

static int ew_grid(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }

extern "C" int tt_elu_bwd_bf16(const void* gy, const void* a, void* dz, int64_t n, void* stream) {
    TT_REQUIRE(gy && a && dz && n % 8 == 0, "elu_bwd_bf16: null argument or element count not a multiple of 8");
    if (n <= 0) return TT_OK;
    elu_bwd_bf16_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)gy, (const uint4*)a, (uint4*)dz, n / 8);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_res_out_bwd_bf16(const void* gy, const void* y, const void* x, void* dz, int64_t n, void* stream) {
    TT_REQUIRE(gy && y && x && dz && n % 8 == 0, "res_out_bwd_bf16: null argument or element count not a multiple of 8");
    if (n <= 0) return TT_OK;
    res_out_bwd_bf16_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)gy, (const uint4*)y, (const uint4*)x, (uint4*)dz, n / 8);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_p4_to_c8(const void* p4, void* c8, int64_t n_pixels, void* stream) {
    TT_REQUIRE(p4 && c8, "null argument");
    if (n_pixels <= 0) return TT_OK;
    p4_to_c8_kernel<<<ew_grid(n_pixels), 256, 0, (cudaStream_t)stream>>>((const uint2*)p4, (uint4*)c8, n_pixels);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}

extern "C" int tt_c8_to_p4(const void* c8, void* p4, int64_t n_pixels, void* stream) {
    TT_REQUIRE(p4 && c8, "null argument");
    if (n_pixels <= 0) return TT_OK;
    c8_to_p4_kernel<<<ew_grid(n_pixels), 256, 0, (cudaStream_t)stream>>>((const uint4*)c8, (uint2*)p4, n_pixels);
    tt_count_launches(1);
    TT_CUDA_CHECK(cudaGetLastError());
    return TT_OK;
}
