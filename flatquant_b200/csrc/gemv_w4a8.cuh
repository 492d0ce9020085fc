// Decode GEMV: y[m, n] = (sum_k W[n,k] * X[m,k]) * s_x[m] * s_w[n] (+bias) for M <= 2 tokens.
//
// With one or two tokens the FlatLinear is a pure stream of the packed 4-bit weight (4 op/B), and a
// tcgen05.mma would spend its ~45-cycle issue floor on a 128 x 16 tile.  The integer dot-product pipe (dp4a) keeps
// up with HBM instead:
//   * a warp owns one weight row at a time and reads it with fully coalesced 128-bit loads (512 contiguous bytes per
//     instruction, up to 4 in flight per lane); rows are strided over all warps of a persistent grid;
//   * nibbles are expanded exactly as in the tensor-core kernel ((p<<4)&0xF0F0F0F0, p&0xF0F0F0F0 -> 16*w as int8 x 4),
//     which is why the activation operand format (K-permuted int8) is shared between the two kernels;
//   * the M activation rows are staged once per CTA in shared memory; per-lane partial sums are combined with five
//     shuffles per row; the x16 is removed by an exact arithmetic shift.
// (A first version with one row per lane was limited to ~1.8 TB/s by uncoalesced L1 tag traffic.  3 <= M <= 32 is the
//  weight-streaming mma.sync kernel of gemm_w4a8_skinny.cuh, 33 <= M <= 64 the split-K tcgen05 route.)
// Replaces the same reference calls as gemm_w4a8.cuh (deploy/kernels/gemm.cu:7-46 + quant.cu:66-101) at decode shapes.
#pragma once
#include "ptx.cuh"

namespace flatq {

constexpr int kGemvWarps = 8;

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
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MT>
__global__ void __launch_bounds__(kGemvWarps * 32) gemv_w4a8_kernel(const GemvParams p) {
    extern __shared__ __align__(16) uint8_t xs[];  // [MT][K] int8
    grid_dep_launch();
    grid_dep_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < MT * p.K / 16; i += blockDim.x) {
        const int m = i / (p.K / 16);
        reinterpret_cast<uint4*>(xs)[i] = m < p.M ? reinterpret_cast<const uint4*>(p.x)[i] : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int chunks = p.K / 32;                       // 16-byte packed chunks per row
    const int gw = blockIdx.x * kGemvWarps + warp, nw = gridDim.x * kGemvWarps;
    for (int n = gw; n < p.N; n += nw) {
        const uint8_t* wrow = p.w + (size_t)n * (p.K / 2);
        int32_t acc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = 0;
#pragma unroll 4
        for (int c = lane; c < chunks; c += 32) {
            const uint4 wv = ldg_nc_v4(wrow + (size_t)c * 16);
            const uint32_t pw[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const uint4 x0 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32);
                const uint4 x1 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32 + 16);
                const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                int32_t a = acc[m];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a = __dp4a((int)((pw[j] << 4) & 0xF0F0F0F0u), (int)xw[2 * j], a);   // weights 0,2,4,6 (x16)
                    a = __dp4a((int)(pw[j] & 0xF0F0F0F0u), (int)xw[2 * j + 1], a);      // weights 1,3,5,7 (x16)
                }
                acc[m] = a;
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
        }
        if (lane < MT && lane < p.M) {
            int32_t s = acc[0];
#pragma unroll
            for (int m = 1; m < MT; ++m) if (lane == m) s = acc[m];
            float y = __fmul_rn(__fmul_rn((float)(s >> 4), p.s_x[lane]), p.s_w[n]);  // exact shift: every term is 16*w*x
            if (p.bias) {
                y = __fadd_rn(y, p.bf16_out ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                            : __half2float(reinterpret_cast<const __half*>(p.bias)[n]));
            }
            if (p.bf16_out) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)lane * p.N + n] = __float2bfloat16_rn(y);
            else reinterpret_cast<__half*>(p.out)[(size_t)lane * p.N + n] = __float2half_rn(y);
        }
    }
}

// The same GEMV with the weight stream started BEFORE the dependency wait.  In a decode step the kernel in front of a
// GEMV is the one-CTA transform + quantise kernel (6-10 us on one SM); launched with PDL, this kernel is resident on the
// other SMs meanwhile, and its weights are static.  A warp owns R rows at a time and keeps all R x CPL 16-byte chunks of
// them in registers (CPL = K / 1024 chunks per lane and row); the first R rows are requested before griddepcontrol.wait,
// so that once the activation codes exist only the dp4a pass is left.  R x CPL = 14..16 loads in flight per lane also
// lifts the steady-state stream of the longer GEMVs (gate/up) over the 4 in flight of the generic kernel.
// Same arithmetic, same operand order: bit-identical to gemv_w4a8_kernel.
// SILU (R == 2, fp16, no bias): the weight rows are interleaved [gate_0, up_0, gate_1, ...] (flatq_gemm_w4a8_silu); a warp
// owns the PAIR (2u, 2u + 1) and stores half(silu_f32(half(g))) * half(u) into column u of an [M, N / 2] output -- the
// rounding points of the GEMM followed by silu_mul_il_f16_kernel, without the launch.
template <int MT, int CPL, int R, int MINB, bool SILU>
__global__ void __launch_bounds__(kGemvWarps * 32, MINB) gemv_w4a8_pre_kernel(const GemvParams p) {
    static_assert(!SILU || R == 2, "SiLU pairs: two rows per warp and pass");
    extern __shared__ __align__(16) uint8_t xs[];  // [MT][K] int8
    grid_dep_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kGemvWarps + warp, nw = gridDim.x * kGemvWarps;
    const size_t row_bytes = (size_t)p.K / 2;
    const int units = SILU ? p.N / 2 : p.N;              // what a warp strides over: row pairs or rows
    auto row_of = [&](int u0, int r) { return SILU ? 2 * u0 + r : u0 + r * nw; };
    uint4 wb[R][CPL];
    auto load_rows = [&](int u0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = row_of(u0, r);
            if (n < p.N) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) wb[r][j] = ldg_nc_v4(p.w + (size_t)n * row_bytes + (size_t)(lane + 32 * j) * 16);
            }
        }
    };
    if (gw < units) load_rows(gw);
    grid_dep_wait();
    for (int i = threadIdx.x; i < MT * p.K / 16; i += blockDim.x) {
        const int m = i / (p.K / 16);
        reinterpret_cast<uint4*>(xs)[i] = m < p.M ? reinterpret_cast<const uint4*>(p.x)[i] : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int stride = (SILU ? 1 : R) * nw;
    for (int u0 = gw; u0 < units; u0 += stride) {
        int32_t acc[R][MT];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int m = 0; m < MT; ++m) acc[r][m] = 0;
            if (row_of(u0, r) >= p.N) break;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const uint32_t pw[4] = {wb[r][j].x, wb[r][j].y, wb[r][j].z, wb[r][j].w};
                const int c = lane + 32 * j;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const uint4 x0 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32);
                    const uint4 x1 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32 + 16);
                    const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    int32_t a = acc[r][m];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        a = __dp4a((int)((pw[q] << 4) & 0xF0F0F0F0u), (int)xw[2 * q], a);
                        a = __dp4a((int)(pw[q] & 0xF0F0F0F0u), (int)xw[2 * q + 1], a);
                    }
                    acc[r][m] = a;
                }
            }
        }
        if (u0 + stride < units) load_rows(u0 + stride);   // the next rows are in flight under the reduction and the epilogue
        float yv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = row_of(u0, r);
            yv[r] = 0.f;
            if (n >= p.N) break;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][m] += __shfl_xor_sync(0xffffffffu, acc[r][m], o);
            }
            if (lane < MT && lane < p.M) {
                int32_t s = acc[r][0];
#pragma unroll
                for (int m = 1; m < MT; ++m) if (lane == m) s = acc[r][m];
                float y = __fmul_rn(__fmul_rn((float)(s >> 4), p.s_x[lane]), p.s_w[n]);
                if (SILU) {
                    yv[r] = y;
                } else {
                    if (p.bias) {
                        y = __fadd_rn(y, p.bf16_out ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                                    : __half2float(reinterpret_cast<const __half*>(p.bias)[n]));
                    }
                    if (p.bf16_out) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)lane * p.N + n] = __float2bfloat16_rn(y);
                    else reinterpret_cast<__half*>(p.out)[(size_t)lane * p.N + n] = __float2half_rn(y);
                }
            }
        }
        if (SILU && lane < MT && lane < p.M) {
            const float g = __half2float(__float2half_rn(yv[0]));
            const __half act = __float2half_rn(g / (1.f + expf(-g)));
            reinterpret_cast<__half*>(p.out)[(size_t)lane * (p.N / 2) + u0] = __hmul(act, __float2half_rn(yv[R - 1]));
        }
    }
}

}  // namespace flatq
