// HBM-bound byte/integer kernels of the path: nibble pack/unpack, per-token quantise, dequantise.
// All are one pass, 128-bit global accesses, grid-stride over vectors.
#pragma once
#include "ptx.cuh"

namespace flatq {

__device__ __forceinline__ int clamp_i4(int v) { return max(-8, min(7, v)); }

// [rows, cols] int8 codes in [-8,7] -> [rows, cols/2] bytes; element 2j -> bits[3:0], 2j+1 -> bits[7:4]
// (deploy/functional/quantization.py:49-56).  One thread: 16 codes in, 8 bytes out.
__global__ void pack_i4_kernel(const int8_t* __restrict__ q, uint8_t* __restrict__ out, size_t n_vec16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(q)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t x = w[2 * h + j];  // 4 codes -> 2 bytes
                const uint32_t b0 = (x & 0xF) | ((x >> 4) & 0xF0);
                const uint32_t b1 = ((x >> 16) & 0xF) | ((x >> 20) & 0xF0);
                r |= (b0 | (b1 << 8)) << (16 * j);
            }
            o[h] = r;
        }
        reinterpret_cast<uint2*>(out)[i] = make_uint2(o[0], o[1]);
    }
}

// bytes -> sign-extended int32 codes (deploy/functional/quantization.py:59-82).  One thread: 4 bytes in,
// 8 int32 out.
__global__ void unpack_i4_kernel(const uint8_t* __restrict__ p, int32_t* __restrict__ out, size_t n_words) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(p)[i];
        int32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (static_cast<int32_t>(w << (28 - 4 * j))) >> 28;
        int4* o = reinterpret_cast<int4*>(out + i * 8);
        o[0] = make_int4(v[0], v[1], v[2], v[3]);
        o[1] = make_int4(v[4], v[5], v[6], v[7]);
    }
}

// packed [rows, K/2] -> int8 [rows, K] holding 16*code, K-permuted [0,2,4,6,1,3,5,7] inside each
// group of 8: the B-operand format of gemm_w4a8 (acc_shift = 8).  One thread: 16 B in, 32 B out.
__global__ void expand_i4_to_i8x16_kernel(const uint8_t* __restrict__ p, uint8_t* __restrict__ out, size_t n_vec16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(p)[i];
        uint4 o0, o1;
        o0.x = (v.x << 4) & 0xF0F0F0F0u; o0.y = v.x & 0xF0F0F0F0u;
        o0.z = (v.y << 4) & 0xF0F0F0F0u; o0.w = v.y & 0xF0F0F0F0u;
        o1.x = (v.z << 4) & 0xF0F0F0F0u; o1.y = v.z & 0xF0F0F0F0u;
        o1.z = (v.w << 4) & 0xF0F0F0F0u; o1.w = v.w & 0xF0F0F0F0u;
        reinterpret_cast<uint4*>(out)[2 * i] = o0;
        reinterpret_cast<uint4*>(out)[2 * i + 1] = o1;
    }
}

// deploy.sym_quant: q = clamp(rn(half(x / scale[row])), -8, 7), packed (deploy/kernels/quant.cu:13-63).
// Division and the half rounding of the quotient follow __hdiv exactly (fp32 quotient of two halves
// rounds to the correctly rounded half).  One thread: 8 halves in, 4 bytes out.
__global__ void sym_quant_f16_kernel(const __half* __restrict__ x, const __half* __restrict__ scale,
                                     uint8_t* __restrict__ out, int rows, int cols) {
    const int vec_per_row = cols >> 3;
    const size_t n_vec = (size_t)rows * vec_per_row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / vec_per_row);
        const float s = __half2float(scale[row]);
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            const int q0 = clamp_i4(__half2int_rn(__float2half_rn(__fdiv_rn(f.x, s))));
            const int q1 = clamp_i4(__half2int_rn(__float2half_rn(__fdiv_rn(f.y, s))));
            o |= ((q0 & 0xF) | ((q1 & 0xF) << 4)) << (8 * j);
        }
        reinterpret_cast<uint32_t*>(out)[i] = o;
    }
}

// y = acc * s_row[m] * s_col[n] (+bias) in fp32, rounded once to fp16 -- the exact form of
// deploy/nn/linear.py:48-54.  mode 1 reproduces the reference kernel's arithmetic bit for bit
// (deploy/kernels/quant.cu:4-10,66-85: half(int(acc/10)) * 10 in half precision).
template <typename OutT>
__global__ void sym_dequant_kernel(const int32_t* __restrict__ acc, const __half* __restrict__ s_row,
                                   const __half* __restrict__ s_col, const OutT* __restrict__ bias,
                                   OutT* __restrict__ out, int rows, int cols, int mode) {
    const int vec_per_row = cols >> 2;
    const size_t n_vec = (size_t)rows * vec_per_row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / vec_per_row);
        const int col = (int)(i % vec_per_row) * 4;
        const int4 a = reinterpret_cast<const int4*>(acc)[i];
        const int av[4] = {a.x, a.y, a.z, a.w};
        OutT o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (mode == 1) {
                int t = (int)((float)av[j] / 10.0f);
                t = max(-65176, min(65176, t));
                const __half xe = __int2half_rn(t);
                const __half y = __hmul(__hmul(__hmul(s_row[row], s_col[col + j]), xe), __float2half(10.0f));
                o[j] = static_cast<OutT>(__half2float(y));
            } else {
                float y = __fmul_rn(__fmul_rn((float)av[j], __half2float(s_row[row])), __half2float(s_col[col + j]));
                if (bias) y = __fadd_rn(y, static_cast<float>(bias[col + j]));
                o[j] = static_cast<OutT>(y);
            }
        }
        reinterpret_cast<uint2*>(out)[i] = *reinterpret_cast<const uint2*>(o);
    }
}

// split-K finalize: y = acc * s_x[m] * s_w[n] (+bias) from the int32 workspace the GEMM CTAs accumulated into
template <typename OutT>
__global__ void splitk_finalize_kernel(const int32_t* __restrict__ acc, const float* __restrict__ s_x,
                                       const float* __restrict__ s_w, const OutT* __restrict__ bias, OutT* __restrict__ out,
                                       int rows, int cols) {
    const size_t n = (size_t)rows * cols;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        float y = __fmul_rn(__fmul_rn((float)acc[i], s_x[r]), s_w[c]);
        if (bias) y = __fadd_rn(y, static_cast<float>(bias[c]));
        out[i] = static_cast<OutT>(y);
    }
}

// fp16 -> fp32 scale conversion for the fused-epilogue GEMM entry point
__global__ void half_to_float_kernel(const __half* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __half2float(in[i]);
}

}  // namespace flatq
