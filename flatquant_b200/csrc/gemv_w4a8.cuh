// Decode GEMV: y[m, n] = (sum_k W[n,k] * X[m,k]) * s_x[m] * s_w[n] (+bias) for M <= 16 tokens.
//
// With a handful of tokens the FlatLinear is a pure stream of the packed 4-bit weight (4 op/B), and a
// tcgen05.mma would spend its >= 128-cycle issue slot on a 128 x 16 tile; the integer dot-product pipe (dp4a) at one
// weight row per lane keeps up with HBM instead:
//   * lane = one weight row, warp w of the CTA = K-slice w (16 slices), so a CTA covers 32 rows x all of K and
//     128-bit loads walk each row sequentially (both halves of every 32-byte sector are consumed back to back);
//   * nibbles are expanded exactly as in the tensor-core kernel ((p<<4)&0xF0F0F0F0, p&0xF0F0F0F0 -> 16*w as int8 x 4),
//     which is why the activation operand format (K-permuted int8) is shared between the two kernels;
//   * the M activation rows are read through L1 with warp-uniform addresses (broadcast), accumulators are M int32
//     registers per lane, the 8 K-slices are summed through shared memory, the x16 is removed by an exact shift.
// Replaces the same reference calls as gemm_w4a8.cuh (deploy/kernels/gemm.cu:7-46 + quant.cu:66-101) at decode shapes.
#pragma once
#include "ptx.cuh"

namespace flatq {

constexpr int kGemvWarps = 16;  // K-slices per CTA: 512 threads x 4 loads in flight = 32 KB of weight stream per CTA
constexpr int kGemvRows = 32;

struct GemvParams {
    const uint8_t* w;     // [N, K/2] packed
    const int8_t* x;      // [M, K] int8 codes, K-permuted
    const float* s_x;     // [M]
    const float* s_w;     // [N]
    const void* bias;     // [N] out dtype or null
    void* out;            // [M, N] f16 / bf16
    int M, N, K;
    int bf16_out;
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MT>
__global__ void __launch_bounds__(kGemvWarps * 32) gemv_w4a8_kernel(const GemvParams p) {
    __shared__ int32_t red[kGemvWarps][MT][kGemvRows + 1];
    grid_dep_launch();
    grid_dep_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * kGemvRows + lane;
    const bool n_ok = n < p.N;
    // K-slice of this warp, in units of 32 weights (16 packed bytes)
    const int chunks = p.K / 32;
    const int c0 = (int)((long long)chunks * warp / kGemvWarps), c1 = (int)((long long)chunks * (warp + 1) / kGemvWarps);
    const uint8_t* wrow = p.w + (size_t)(n_ok ? n : 0) * (p.K / 2);
    int32_t acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
        const uint4 wv = ldg_nc_v4(wrow + (size_t)c * 16);
        const uint32_t pw[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo[j] = (pw[j] << 4) & 0xF0F0F0F0u;  // weights 0,2,4,6 of the 8-group, as 16*w
            hi[j] = pw[j] & 0xF0F0F0F0u;         // weights 1,3,5,7
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m < p.M) {
                const int8_t* xr = p.x + (size_t)m * p.K + (size_t)c * 32;  // warp-uniform address: L1 broadcast
                const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(xr));
                const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(xr) + 1);
                int32_t a = acc[m];
                a = __dp4a((int)lo[0], (int)x0.x, a); a = __dp4a((int)hi[0], (int)x0.y, a);
                a = __dp4a((int)lo[1], (int)x0.z, a); a = __dp4a((int)hi[1], (int)x0.w, a);
                a = __dp4a((int)lo[2], (int)x1.x, a); a = __dp4a((int)hi[2], (int)x1.y, a);
                a = __dp4a((int)lo[3], (int)x1.z, a); a = __dp4a((int)hi[3], (int)x1.w, a);
                acc[m] = a;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) red[warp][m][lane] = acc[m];
    __syncthreads();
    // warp w finishes tokens m = w, w + 16, ...
    for (int m = warp; m < p.M && m < MT; m += kGemvWarps) {
        int32_t s = 0;
#pragma unroll
        for (int w = 0; w < kGemvWarps; ++w) s += red[w][m][lane];
        if (n_ok) {
            float y = __fmul_rn(__fmul_rn((float)(s >> 4), p.s_x[m]), p.s_w[n]);  // exact shift: every term is 16*w*x
            if (p.bias) {
                y = __fadd_rn(y, p.bf16_out ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                            : __half2float(reinterpret_cast<const __half*>(p.bias)[n]));
            }
            if (p.bf16_out) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)m * p.N + n] = __float2bfloat16_rn(y);
            else reinterpret_cast<__half*>(p.out)[(size_t)m * p.N + n] = __float2half_rn(y);
        }
    }
}

}  // namespace flatq
