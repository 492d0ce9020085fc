// Online Kronecker transform + per-token symmetric INT4 quantisation (v0: warp-level HMMA).
//
// Replaces the Triton kernels deploy/kernels/kron_matmul.py:24-189 and block_matmul.py:29-104 and the
// torch path flatquant/flat_utils.py:6-17 + flatquant/quant_utils.py:77-119:
//     Z = L^T . X . R   per token (X = token reshaped [n1, n2] row-major)
//     hi = max(max Z, 0) * c_max,  lo = min(min Z, 0) * c_min,  s = max(|lo|, hi) / 7  (1 if 0)
//     q = clamp(rint(Z / s), -8, 7)
// Products are evaluated as (L^T X) R with the intermediate rounded to fp16, like the reference's
// fused kernel (kron_matmul.py:62-76); Z is rounded to fp16 (the dtype the reference's transform
// returns) and the scale / division / rounding are done in fp32, which is what the reference's
// LAC path does after type promotion (quant_utils.py:96-107).
//
// One warp owns a 16-row slab of a token; n1/16 warps form a token, several tokens per CTA.
// Outputs (each optional): fp16 Z, packed nibbles [tokens, K/2] (deploy format), int8 codes
// [tokens, K] in the K-permuted operand format of gemm_w4a8, per-token scale as fp32 and fp16.
#pragma once
#include "ptx.cuh"

namespace flatq {

struct KronParams {
    const __half* x;      // [tokens, n1*n2]
    const __half* left;   // [n1, n1] row-major, or nullptr (identity)
    const __half* right;  // [n2, n2] row-major, or nullptr (identity)
    int tokens, n1, n2, p1;  // p1 = n1 rounded up to 16
    int tokens_per_cta;
    int do_quant;
    float c_max, c_min;   // multiplicative clip factors (already sigmoid-ed / clip ratio)
    __half* z_out;        // [tokens, K] or nullptr
    uint8_t* xq4;         // [tokens, K/2] or nullptr
    int8_t* xq8;          // [tokens, K] permuted, or nullptr
    float* sx_f32;        // [tokens] or nullptr
    __half* sx_f16;       // [tokens] or nullptr
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// smem bytes needed (host + device agree through this)
__host__ __device__ inline size_t kron_smem_bytes(int n1, int n2, int p1, int tpc, bool has_l, bool has_r) {
    size_t b = 0;
    if (has_l) b += (size_t)p1 * (p1 + 8) * 2;
    if (has_r) b += (size_t)n2 * (n2 + 8) * 2;
    b += (size_t)tpc * p1 * (n2 + 8) * 2;         // X tiles; the int8 + nibble staging (1.5 B/elem) aliases them
    b += (size_t)tpc * 16 * 2 * sizeof(float);    // max/min partials (<=16 warps per token)
    return b + 16;
}

template <int KT2>  // n2 == 16*KT2
__global__ void __launch_bounds__(KT2 <= 8 ? 512 : 256)
kron_quant_mma_kernel(const KronParams p) {
    constexpr int P2 = 16 * KT2;
    constexpr int LDX = P2 + 8;  // halves; +16 B keeps ldmatrix rows on distinct banks
    const int P1 = p.p1, LDL = P1 + 8;
    const bool has_l = p.left != nullptr, has_r = p.right != nullptr;
    extern __shared__ __align__(16) uint8_t smem[];
    __half* Ls = reinterpret_cast<__half*>(smem);
    __half* Rs = Ls + (has_l ? (size_t)P1 * LDL : 0);
    __half* Xs = Rs + (has_r ? (size_t)P2 * LDX : 0);
    const int tpc = p.tokens_per_cta;
    // staging for the quantised rows reuses the X tiles: every warp is past its last ldmatrix of X when the
    // block-wide barrier after the max/min exchange has been crossed
    int8_t* q8s = reinterpret_cast<int8_t*>(Xs);
    uint8_t* q4s = reinterpret_cast<uint8_t*>(q8s + (size_t)tpc * P1 * P2);
    float* red = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Xs + (size_t)tpc * P1 * LDX) + 15) & ~uintptr_t(15));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int wpt = P1 >> 4;
    const int tl = warp / wpt, slab = warp % wpt;
    const int K = p.n1 * p.n2;

    // ---- factors -> smem, zero padded
    if (has_l) {
        for (int i = threadIdx.x; i < P1 * LDL; i += blockDim.x) {
            const int r = i / LDL, c = i % LDL;
            Ls[i] = (r < p.n1 && c < p.n1) ? p.left[r * p.n1 + c] : __float2half(0.f);
        }
    }
    if (has_r) {
        for (int i = threadIdx.x; i < P2 * LDX; i += blockDim.x) {
            const int r = i / LDX, c = i % LDX;
            Rs[i] = (c < P2) ? p.right[r * P2 + c] : __float2half(0.f);
        }
    }

    grid_dep_launch();  // PDL: factors are static weights; activations are read after the wait
    grid_dep_wait();
    const int mi = lane >> 3, r8 = lane & 7;
    for (int tb = blockIdx.x * tpc; tb < p.tokens; tb += gridDim.x * tpc) {
        __syncthreads();
        // ---- X tiles -> smem (16-byte vectors; rows >= n1 zero)
        {
            const int vec_per_row = P2 / 8;
            const int total = tpc * P1 * vec_per_row;
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int t = i / (P1 * vec_per_row);
                const int rem = i % (P1 * vec_per_row);
                const int r = rem / vec_per_row, v = rem % vec_per_row;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (tb + t < p.tokens && r < p.n1)
                    val = *reinterpret_cast<const uint4*>(p.x + (size_t)(tb + t) * K + (size_t)r * P2 + v * 8);
                *reinterpret_cast<uint4*>(Xs + ((size_t)t * P1 + r) * LDX + v * 8) = val;
            }
        }
        __syncthreads();

        const int token = tb + tl;
        const bool valid = token < p.tokens;  // warp-uniform
        const __half* Xt = Xs + (size_t)tl * P1 * LDX;
        const int m0 = slab * 16;
        uint32_t ufrag[KT2][4];
        if (has_l) {
            // U[16 x P2] = L^T[slab rows, :] . X       A = L^T via transposed ldmatrix of row-major L
#pragma unroll
            for (int kt2 = 0; kt2 < KT2; ++kt2) {
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k0 = 0; k0 < P1; k0 += 16) {
                    uint32_t a[4], b[4];
                    ldsm_x4_t(smem_u32(Ls + (size_t)(k0 + (mi >> 1) * 8 + r8) * LDL + m0 + (mi & 1) * 8), a);
                    ldsm_x4_t(smem_u32(Xt + (size_t)(k0 + (mi & 1) * 8 + r8) * LDX + kt2 * 16 + (mi >> 1) * 8), b);
                    hmma16816(c0, a, b[0], b[1]);
                    hmma16816(c1, a, b[2], b[3]);
                }
                ufrag[kt2][0] = pack_h2(c0[0], c0[1]);
                ufrag[kt2][1] = pack_h2(c0[2], c0[3]);
                ufrag[kt2][2] = pack_h2(c1[0], c1[1]);
                ufrag[kt2][3] = pack_h2(c1[2], c1[3]);
            }
        } else {
#pragma unroll
            for (int kt2 = 0; kt2 < KT2; ++kt2)
                ldsm_x4(smem_u32(Xt + (size_t)(m0 + (mi & 1) * 8 + r8) * LDX + kt2 * 16 + (mi >> 1) * 8), ufrag[kt2]);
        }

        // Z[16 x P2] = U . R ; z[nt][0] = (row g, cols 8nt+2tig,+1), z[nt][1] = (row g+8, same cols)
        uint32_t z[2 * KT2][2];
        if (has_r) {
#pragma unroll
            for (int nt2 = 0; nt2 < KT2; ++nt2) {
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kt = 0; kt < KT2; ++kt) {
                    uint32_t b[4];
                    ldsm_x4_t(smem_u32(Rs + (size_t)(kt * 16 + (mi & 1) * 8 + r8) * LDX + nt2 * 16 + (mi >> 1) * 8), b);
                    hmma16816(c0, ufrag[kt], b[0], b[1]);
                    hmma16816(c1, ufrag[kt], b[2], b[3]);
                }
                z[2 * nt2][0] = pack_h2(c0[0], c0[1]);
                z[2 * nt2][1] = pack_h2(c0[2], c0[3]);
                z[2 * nt2 + 1][0] = pack_h2(c1[0], c1[1]);
                z[2 * nt2 + 1][1] = pack_h2(c1[2], c1[3]);
            }
        } else {
#pragma unroll
            for (int kt = 0; kt < KT2; ++kt) {
                z[2 * kt][0] = ufrag[kt][0];
                z[2 * kt][1] = ufrag[kt][1];
                z[2 * kt + 1][0] = ufrag[kt][2];
                z[2 * kt + 1][1] = ufrag[kt][3];
            }
        }

        const int row0 = m0 + g, row1 = m0 + g + 8;
        if (p.z_out && valid) {
#pragma unroll
            for (int nt = 0; nt < 2 * KT2; ++nt) {
                if (row0 < p.n1)
                    *reinterpret_cast<uint32_t*>(p.z_out + (size_t)token * K + (size_t)row0 * P2 + nt * 8 + 2 * tig) = z[nt][0];
                if (row1 < p.n1)
                    *reinterpret_cast<uint32_t*>(p.z_out + (size_t)token * K + (size_t)row1 * P2 + nt * 8 + 2 * tig) = z[nt][1];
            }
        }
        if (!p.do_quant) continue;  // uniform across the CTA

        // ---- per-token max / min
        float vmax = 0.f, vmin = 0.f;  // hi = max(amax, 0), lo = min(amin, 0)
#pragma unroll
        for (int nt = 0; nt < 2 * KT2; ++nt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&z[nt][h]));
                vmax = fmaxf(vmax, fmaxf(f.x, f.y));
                vmin = fminf(vmin, fminf(f.x, f.y));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        if (lane == 0) { red[(tl * 16 + slab) * 2] = vmax; red[(tl * 16 + slab) * 2 + 1] = vmin; }
        __syncthreads();
        for (int s = 0; s < wpt; ++s) {
            vmax = fmaxf(vmax, red[(tl * 16 + s) * 2]);
            vmin = fminf(vmin, red[(tl * 16 + s) * 2 + 1]);
        }
        const float hi = vmax * p.c_max, lo = vmin * p.c_min;
        const float amax = fmaxf(fabsf(lo), hi);
        const float scale = amax == 0.f ? 1.f : __fdiv_rn(amax, 7.0f);
        if (valid && slab == 0 && lane == 0) {
            if (p.sx_f32) p.sx_f32[token] = scale;
            if (p.sx_f16) p.sx_f16[token] = __float2half_rn(scale);
        }

        // ---- quantise into the staging tiles
        int8_t* q8t = q8s + (size_t)tl * P1 * P2;
        uint8_t* q4t = q4s + (size_t)tl * P1 * (P2 / 2);
#pragma unroll
        for (int nt = 0; nt < 2 * KT2; ++nt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = h ? row1 : row0;
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&z[nt][h]));
                const int qa = max(-8, min(7, __float2int_rn(__fdiv_rn(f.x, scale))));
                const int qb = max(-8, min(7, __float2int_rn(__fdiv_rn(f.y, scale))));
                // column pair (8nt+2tig, 8nt+2tig+1): one packed byte; permuted int8 slots tig and 4+tig
                q4t[row * (P2 / 2) + nt * 4 + tig] = (uint8_t)((qa & 0xF) | ((qb & 0xF) << 4));
                q8t[row * P2 + nt * 8 + tig] = (int8_t)qa;
                q8t[row * P2 + nt * 8 + 4 + tig] = (int8_t)qb;
            }
        }
        __syncthreads();
        // ---- coalesced copy-out of the rows < n1 (contiguous per token)
        if (valid) {
            const int tid_t = slab * 32 + lane, nthr_t = wpt * 32;
            if (p.xq8) {
                const uint4* src = reinterpret_cast<const uint4*>(q8t);
                uint4* dst = reinterpret_cast<uint4*>(p.xq8 + (size_t)token * K);
                for (int i = tid_t; i < K / 16; i += nthr_t) dst[i] = src[i];
            }
            if (p.xq4) {
                const uint4* src = reinterpret_cast<const uint4*>(q4t);
                uint4* dst = reinterpret_cast<uint4*>(p.xq4 + (size_t)token * (K / 2));
                for (int i = tid_t; i < K / 32; i += nthr_t) dst[i] = src[i];
            }
        }
    }
}

// Plain per-token quantiser for already-transformed (or untransformed) fp16 rows: the
// deploy.nn.Quantizer path (deploy/nn/quantization.py:13-36) fused into one pass.
// One warp per token row; row cached in registers when K <= 32*8*VEC.
__global__ void __launch_bounds__(256)
rowquant_kernel(const __half* __restrict__ x, int tokens, int K, float c_max, float c_min, int half_math,
                uint8_t* __restrict__ xq4, int8_t* __restrict__ xq8, float* __restrict__ sx_f32,
                __half* __restrict__ sx_f16) {
    grid_dep_launch();
    grid_dep_wait();
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp_g; t < tokens; t += nwarps) {
        const uint4* row = reinterpret_cast<const uint4*>(x + (size_t)t * K);
        float vmax = 0.f, vmin = 0.f;
        for (int i = lane; i < K / 8; i += 32) {
            const uint4 v = row[i];
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                vmax = fmaxf(vmax, fmaxf(f.x, f.y));
                vmin = fminf(vmin, fminf(f.x, f.y));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        const float amax = fmaxf(fabsf(vmin * c_min), vmax * c_max);
        float scale = amax == 0.f ? 1.f : __fdiv_rn(amax, 7.0f);
        if (half_math) scale = __half2float(__float2half_rn(scale));  // deploy path stores/uses an fp16 scale
        if (lane == 0) {
            if (sx_f32) sx_f32[t] = scale;
            if (sx_f16) sx_f16[t] = __float2half_rn(scale);
        }
        for (int i = lane; i < K / 8; i += 32) {
            const uint4 v = row[i];  // second pass hits L1/L2
            const __half2* h = reinterpret_cast<const __half2*>(&v);
            uint32_t o4 = 0, lo8 = 0, hi8 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                float d0 = __fdiv_rn(f.x, scale), d1 = __fdiv_rn(f.y, scale);
                if (half_math) { d0 = __half2float(__float2half_rn(d0)); d1 = __half2float(__float2half_rn(d1)); }
                const int qa = max(-8, min(7, __float2int_rn(d0)));
                const int qb = max(-8, min(7, __float2int_rn(d1)));
                o4 |= ((qa & 0xF) | ((qb & 0xF) << 4)) << (8 * j);
                lo8 |= (uint32_t)(qa & 0xFF) << (8 * j);
                hi8 |= (uint32_t)(qb & 0xFF) << (8 * j);
            }
            if (xq4) reinterpret_cast<uint32_t*>(xq4 + (size_t)t * (K / 2))[i] = o4;
            if (xq8) reinterpret_cast<uint2*>(xq8 + (size_t)t * K)[i] = make_uint2(lo8, hi8);
        }
    }
}

}  // namespace flatq
