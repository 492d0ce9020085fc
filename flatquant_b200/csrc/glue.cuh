// Elementwise glue between the FlatLinears of a decoder layer (the callers either side of the hot
// path): HF RMSNorm, HF rotary embedding, SiLU(gate) * up.  In the reference these are HF transformers' torch
// expressions (modeling_llama.py: LlamaRMSNorm.forward, apply_rotary_pos_emb, LlamaMLP.forward) -- 5-7 elementwise
// launches each, with fp32 temporaries of the whole activation for the norm.  One HBM pass each here, 128-bit accesses.
// Each kernel reproduces the ROUNDING POINTS of the torch expression it replaces, so that swapping it in does not move
// 4-bit codes downstream:
//   rmsnorm   y = w * half(float(x) * rsqrt(mean(float(x)^2) + eps))          (fp32 statistics, two fp16 roundings)
//   rope      y = x * cos + rotate_half(x) * sin, every op rounded to fp16     (__hmul / __hmul / __hadd, no FMA)
//   silu_mul  y = half(silu_f32(float(g))) * u, the product rounded to fp16
// Opt-in (FLATQ_FUSED_GLUE) until their first run on hardware: written after round 1's GPU minutes were spent.
#pragma once
#include "ptx.cuh"

namespace flatq {

// ---- RMSNorm: one CTA of 256 threads per row; the row stays in registers between the two passes for K <= 8192
constexpr int kNormThreads = 256, kNormVecMax = 4;   // 256 threads x 4 x 8 halves

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = lane < (kNormThreads >> 5) ? s_red[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    t = __shfl_sync(0xffffffffu, t, 0);
    __syncthreads();   // s_red may be reused by the next row
    return t;
}

// weight may be null (the weight-less norm of deploy/nn/normalization.py: the scale is folded into the next transform).
// res != null: the residual add of the decoder layer in the same pass -- h = x + res rounded to fp16 (torch's `h + y`),
// written to h_out, and the norm is taken of that h (saves one elementwise launch and one 4 B/element HBM round trip
// per sub-layer).
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_f16_kernel(const __half* __restrict__ x, const __half* __restrict__ w, __half* __restrict__ y, int rows, int K,
                   float mean_dim, float eps, const __half* __restrict__ res = nullptr, __half* __restrict__ h_out = nullptr) {
    __shared__ float s_red[kNormThreads / 32];
    grid_dep_launch();   // PDL: the next kernel of a decode step becomes resident while this one runs
    grid_dep_wait();
    const int nvec = K / 8;   // uint4 vectors per row
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)row * K);
        const uint4* rsrc = res ? reinterpret_cast<const uint4*>(res + (size_t)row * K) : nullptr;
        uint4* hdst = h_out ? reinterpret_cast<uint4*>(h_out + (size_t)row * K) : nullptr;
        auto fetch = [&](int idx) {   // x (+ res, rounded to fp16 like torch's add; stored back as the new residual stream)
            uint4 a = src[idx];
            if (rsrc) {
                const uint4 b = rsrc[idx];
                __half2* ah = reinterpret_cast<__half2*>(&a);
                const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                for (int j = 0; j < 4; ++j) ah[j] = __hadd2_rn(ah[j], bh[j]);
                if (hdst) hdst[idx] = a;
            }
            return a;
        };
        uint4 v[kNormVecMax];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kNormVecMax; ++i) {
            const int idx = threadIdx.x + i * kNormThreads;
            v[i] = make_uint4(0, 0, 0, 0);
            if (idx < nvec) {
                v[i] = fetch(idx);
                const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
        }
        for (int idx = threadIdx.x + kNormVecMax * kNormThreads; idx < nvec; idx += kNormThreads) {   // K > 8192: re-read later
            const uint4 t = fetch(idx);
            const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
        }
        const float total = block_sum_256(ss, s_red);
        const float rs = rsqrtf(total / mean_dim + eps);
        uint4* dst = reinterpret_cast<uint4*>(y + (size_t)row * K);
        auto emit = [&](int idx, const uint4& in) {
            const __half2* h = reinterpret_cast<const __half2*>(&in);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
            uint4 wv = make_uint4(0, 0, 0, 0);
            if (w) wv = reinterpret_cast<const uint4*>(w)[idx];
            const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                const __half2 n = __floats2half2_rn(f.x * rs, f.y * rs);   // .to(input_dtype)
                oh[j] = w ? __hmul2(wh[j], n) : n;                         // weight * normalised, rounded to fp16
            }
            dst[idx] = o;
        };
#pragma unroll
        for (int i = 0; i < kNormVecMax; ++i) {
            const int idx = threadIdx.x + i * kNormThreads;
            if (idx < nvec) emit(idx, v[i]);
        }
        // K > 8192: the tail is re-read -- from h_out when the residual was added (it already holds x + res)
        for (int idx = threadIdx.x + kNormVecMax * kNormThreads; idx < nvec; idx += kNormThreads) emit(idx, (rsrc && hdst) ? hdst[idx] : (rsrc ? fetch(idx) : src[idx]));
    }
}

// ---- rotary embedding, in place: x [tokens, heads, 128], cos / sin [tokens, 128] (HF: emb = cat(freqs, freqs)).
// 8 threads per (token, head) row: thread i owns dims [8i, 8i+8) of the first half and the same dims + 64.
__global__ void __launch_bounds__(256)
rope_f16_kernel(__half* __restrict__ x, const __half* __restrict__ cos_t, const __half* __restrict__ sin_t, int tokens, int heads) {
    grid_dep_launch();
    grid_dep_wait();
    const size_t n_units = (size_t)tokens * heads * 8;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < n_units; u += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(u & 7);
        const size_t row = u >> 3;
        const size_t token = row / heads;
        uint4* p = reinterpret_cast<uint4*>(x + row * 128);
        const uint4 lo = p[i], hi = p[8 + i];
        const uint4 c_lo = reinterpret_cast<const uint4*>(cos_t + token * 128)[i], c_hi = reinterpret_cast<const uint4*>(cos_t + token * 128)[8 + i];
        const uint4 s_lo = reinterpret_cast<const uint4*>(sin_t + token * 128)[i], s_hi = reinterpret_cast<const uint4*>(sin_t + token * 128)[8 + i];
        uint4 o_lo, o_hi;
        const __half2* xl = reinterpret_cast<const __half2*>(&lo);
        const __half2* xh = reinterpret_cast<const __half2*>(&hi);
        const __half2* cl = reinterpret_cast<const __half2*>(&c_lo);
        const __half2* ch = reinterpret_cast<const __half2*>(&c_hi);
        const __half2* sl = reinterpret_cast<const __half2*>(&s_lo);
        const __half2* sh = reinterpret_cast<const __half2*>(&s_hi);
        __half2* ol = reinterpret_cast<__half2*>(&o_lo);
        __half2* oh = reinterpret_cast<__half2*>(&o_hi);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // rotate_half(x) = cat(-x[64:], x[:64]):  out_lo = x_lo * cos_lo + (-x_hi) * sin_lo ; out_hi = x_hi * cos_hi + x_lo * sin_hi
            ol[j] = __hadd2_rn(__hmul2_rn(xl[j], cl[j]), __hmul2_rn(__hneg2(xh[j]), sl[j]));   // _rn: never contracted into HFMA2
            oh[j] = __hadd2_rn(__hmul2_rn(xh[j], ch[j]), __hmul2_rn(xl[j], sh[j]));
        }
        p[i] = o_lo;
        p[8 + i] = o_hi;
    }
}

// ---- SiLU(gate) * up over the fused [gate; up] GEMM output: gu [rows, 2 * inter] -> out [rows, inter]
__global__ void __launch_bounds__(256)
silu_mul_f16_kernel(const __half* __restrict__ gu, __half* __restrict__ out, int rows, int inter) {
    grid_dep_launch();
    grid_dep_wait();
    const int nvec = inter / 8;
    const size_t total = (size_t)rows * nvec;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < total; u += (size_t)gridDim.x * blockDim.x) {
        const size_t row = u / nvec;
        const int c = (int)(u % nvec);
        const uint4 g = reinterpret_cast<const uint4*>(gu + row * 2 * inter)[c];
        const uint4 v = reinterpret_cast<const uint4*>(gu + row * 2 * inter + inter)[c];
        const __half2* gh = reinterpret_cast<const __half2*>(&g);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(gh[j]);
            const __half2 s = __floats2half2_rn(f.x / (1.f + expf(-f.x)), f.y / (1.f + expf(-f.y)));   // F.silu: fp32 math, fp16 result
            oh[j] = __hmul2(s, vh[j]);
        }
        reinterpret_cast<uint4*>(out + row * inter)[c] = o;
    }
}

// The same over the row-INTERLEAVED layout gu [rows, 2 * inter] = [g0, u0, g1, u1, ...] (the weight layout
// flatq_gemm_w4a8_silu takes; this kernel is its small-M leg, where the GEMM runs split-K or as a GEMV).
__global__ void __launch_bounds__(256)
silu_mul_il_f16_kernel(const __half* __restrict__ gu, __half* __restrict__ out, size_t total_vec) {
    grid_dep_launch();
    grid_dep_wait();
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < total_vec; u += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(gu)[2 * u], b = reinterpret_cast<const uint4*>(gu)[2 * u + 1];
        const __half2 p[8] = {*reinterpret_cast<const __half2*>(&a.x), *reinterpret_cast<const __half2*>(&a.y),
                              *reinterpret_cast<const __half2*>(&a.z), *reinterpret_cast<const __half2*>(&a.w),
                              *reinterpret_cast<const __half2*>(&b.x), *reinterpret_cast<const __half2*>(&b.y),
                              *reinterpret_cast<const __half2*>(&b.z), *reinterpret_cast<const __half2*>(&b.w)};
        uint4 o;
        __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float g = __low2float(p[j]);
            oh[j] = __hmul(__float2half_rn(g / (1.f + expf(-g))), __high2half(p[j]));
        }
        reinterpret_cast<uint4*>(out)[u] = o;
    }
}

}  // namespace flatq
