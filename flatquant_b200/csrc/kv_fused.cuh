// everything between the q/k/v FlatLinears and the attention, in ONE launch.
//
// The reference does this with ~25 launches per layer (deploy/transformers/modeling_llama.py:125-137 rotary embedding,
// kv_cache.py:263-353: K . T_k (128 x 128 matmul or FHT), asym_quantize_and_pack_i4 (~10 torch kernels each for K and V,
// :11-51), repeat_interleave for GQA, init/append_kv_i4; the query transform Q . T_k^-T at decode), round 1 of this
// project with eight (RoPE q, RoPE k, three transform launches, two quantise+scatter launches, plus torch copies of the
// strided q/k/v views).  Here the rows (token, head) of the fused qkv GEMM output are read ONCE, in place, strided:
//
//   q rows:  RoPE (HF rotate_half, every op rounded to fp16 like apply_rotary_pos_emb) -> . T_k^-T -> q_out fp16
//   k rows:  RoPE -> . T_k -> [k_out fp16 for the prefill attention] -> per-(token, head) asymmetric 4-bit
//            (s = (max-min).clamp(1e-5)/15, zero = -min, the reference's fp16 op-by-op rounding) -> page scatter
//   v rows:  asymmetric 4-bit -> page scatter
//
// One CTA = 128 rows of one kind (8 warps x 16 rows).  The 128 x 128 products run on mma.sync m16n8k16 (fp32
// accumulate, one fp16 rounding of the result -- the arithmetic of the right-only transform kernel it replaces,
// kron_quant.cuh): per token this is 40 rows x 16 K MACs, three orders of magnitude below the tensor-core budget, and
// the kernel is bound by the HBM/L2 pass over q, k and v.  Bytes per token (Llama-3-8B, 32 + 8 + 8 heads): read
// 48 x 256 B, write 32 x 256 B (q) + 16 x 68 B (packed K, V + params) [+ 8 x 256 B of k_out at prefill].
#pragma once
#include "kron_quant.cuh"

namespace flatq {

struct KvFusedParams {
    const __half* q; const __half* k; const __half* v;   // row (token, head) of X at X + token * ld_X + head * 128
    long long ld_q, ld_k, ld_v;                          // token strides in halves (the fused qkv output: all equal)
    int tokens, q_heads, kv_heads;
    const __half* cos_t; const __half* sin_t;            // [positions, 128] tables (HF: emb = cat(freqs, freqs)), or null: no RoPE
    const int* pos;                                      // [tokens] position of every token (device)
    const __half* t_q; const __half* t_k;                // [128, 128] row-major: q_t = q . t_q, k_t = k . t_k; null = identity
    __half* q_out;                                       // [tokens, q_heads, 128]
    __half* k_out;                                       // [tokens, kv_heads, 128] or null
    uint8_t* kv_data; __half* kv_param; const int* slot; // paged INT4 cache (qattention.cuh layout), slot[token]
    int layer, n_layers, page_size;
    int q_blocks, k_blocks;                              // CTAs [0, q_blocks): q rows; [q_blocks, q_blocks + k_blocks): k; rest: v
};

constexpr int kKvfRows = 128, kKvfThreads = 256, kKvfLd = 128 + 8;   // smem row stride in halves (ldmatrix bank spread)
constexpr int kKvfSmem = (128 * kKvfLd + 8 * 16 * kKvfLd) * 2;   // T + per-warp A tiles

// the reference's per-row asymmetric quantiser on 4 consecutive values per lane (kv_quant_kernel, bit-exact)
__device__ __forceinline__ void kvf_quant_row_warp(const __half (&xv)[4], int lane, __half& scale, __half& zero, uint16_t& packed) {
    __half2 h0 = __halves2half2(xv[0], xv[1]), h1 = __halves2half2(xv[2], xv[3]);
    __half2 mx = __hmax2(h0, h1), mn = __hmin2(h0, h1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t a = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mx), o);
        uint32_t b = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mn), o);
        mx = __hmax2(mx, *reinterpret_cast<__half2*>(&a));
        mn = __hmin2(mn, *reinterpret_cast<__half2*>(&b));
    }
    const __half xmax = __hmax(__low2half(mx), __high2half(mx)), xmin = __hmin(__low2half(mn), __high2half(mn));
    __half range = __hsub(xmax, xmin);
    const __half eps = __float2half_rn(1e-5f);
    if (__hlt(range, eps)) range = eps;
    scale = __float2half_rn(__half2float(range) / 15.0f);
    zero = __hneg(xmin);
    const float fs = __half2float(scale);
    uint32_t qv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half t = __hadd(xv[j], zero);
        const __half u = __float2half_rn(__fdiv_rn(__half2float(t), fs));
        qv[j] = (uint32_t)max(0, min(15, __half2int_rn(u)));
    }
    packed = (uint16_t)(qv[0] | (qv[1] << 4) | (qv[2] << 8) | (qv[3] << 12));
    (void)lane;
}

__global__ void __launch_bounds__(kKvfThreads)
kv_fused_kernel(const KvFusedParams p) {
    extern __shared__ __align__(16) uint8_t kvf_smem[];
    __half* Ts = reinterpret_cast<__half*>(kvf_smem);                  // [128][kKvfLd] the block's transform matrix
    __half* As = Ts + 128 * kKvfLd;                                     // [8 warps][16][kKvfLd] RoPE'd rows, then fp16 results
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kind = (int)blockIdx.x < p.q_blocks ? 0 : ((int)blockIdx.x < p.q_blocks + p.k_blocks ? 1 : 2);
    const int blk = kind == 0 ? blockIdx.x : (kind == 1 ? blockIdx.x - p.q_blocks : blockIdx.x - p.q_blocks - p.k_blocks);
    const int heads = kind == 0 ? p.q_heads : p.kv_heads;
    const long long n_rows = (long long)p.tokens * heads;
    const __half* src = kind == 0 ? p.q : (kind == 1 ? p.k : p.v);
    const long long ld = kind == 0 ? p.ld_q : (kind == 1 ? p.ld_k : p.ld_v);
    const __half* T = kind == 0 ? p.t_q : (kind == 1 ? p.t_k : nullptr);

    if (T) {   // static weights: before the dependency wait
        for (int i = threadIdx.x; i < 128 * 16; i += kKvfThreads) {
            const int r = i >> 4, c = (i & 15) * 8;
            *reinterpret_cast<uint4*>(Ts + r * kKvfLd + c) = __ldg(reinterpret_cast<const uint4*>(T + r * 128 + c));
        }
    }
    grid_dep_launch();
    grid_dep_wait();
    __syncthreads();

    const long long row0 = (long long)blk * kKvfRows + warp * 16;
    __half* Aw = As + warp * 16 * kKvfLd;
    if (kind == 2) {
        // ---- V rows: quantise as they are (one row per iteration, 4 values per lane), scatter
        for (int r = 0; r < 16; ++r) {
            const long long row = row0 + r;
            if (row >= n_rows) break;
            const int token = (int)(row / heads), head = (int)(row % heads);
            const uint2 raw = *reinterpret_cast<const uint2*>(src + token * ld + head * 128 + lane * 4);
            const __half xv[4] = {__ushort_as_half((uint16_t)(raw.x & 0xFFFF)), __ushort_as_half((uint16_t)(raw.x >> 16)),
                                  __ushort_as_half((uint16_t)(raw.y & 0xFFFF)), __ushort_as_half((uint16_t)(raw.y >> 16))};
            __half scale, zero;
            uint16_t packed;
            kvf_quant_row_warp(xv, lane, scale, zero, packed);
            const int s = p.slot[token], page = s / p.page_size, off = s % p.page_size;
            const size_t idx = ((((size_t)page * p.n_layers + p.layer) * 2 + 1) * p.kv_heads + head) * p.page_size + off;
            reinterpret_cast<uint16_t*>(p.kv_data + idx * 64)[lane] = packed;
            if (lane == 0) reinterpret_cast<__half2*>(p.kv_param)[idx] = __halves2half2(scale, zero);
        }
        return;
    }

    // ---- q / k rows: RoPE into the warp's A tile
    for (int r = 0; r < 16; ++r) {
        const long long row = row0 + r;
        uint2 out = make_uint2(0, 0);
        if (row < n_rows) {
            const int token = (int)(row / heads), head = (int)(row % heads);
            const uint2 raw = *reinterpret_cast<const uint2*>(src + token * ld + head * 128 + lane * 4);
            out = raw;
            if (p.cos_t) {
                const int ps = p.pos[token];
                const uint2 c = __ldg(reinterpret_cast<const uint2*>(p.cos_t + (size_t)ps * 128 + lane * 4));
                const uint2 s = __ldg(reinterpret_cast<const uint2*>(p.sin_t + (size_t)ps * 128 + lane * 4));
                // rotate_half: lanes 0-15 hold dims [0, 64) and pair with lane + 16 (dims + 64): rot = -x_hi ; lanes 16-31: rot = x_lo
                uint2 other;
                other.x = __shfl_xor_sync(0xffffffffu, raw.x, 16);
                other.y = __shfl_xor_sync(0xffffffffu, raw.y, 16);
                const __half2* xh = reinterpret_cast<const __half2*>(&raw);
                const __half2* oh = reinterpret_cast<const __half2*>(&other);
                const __half2* ch = reinterpret_cast<const __half2*>(&c);
                const __half2* sh = reinterpret_cast<const __half2*>(&s);
                __half2* rh = reinterpret_cast<__half2*>(&out);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __half2 rot = lane < 16 ? __hneg2(oh[j]) : oh[j];
                    rh[j] = __hadd2_rn(__hmul2_rn(xh[j], ch[j]), __hmul2_rn(rot, sh[j]));   // _rn: never contracted into HFMA2
                }
            }
        }
        *reinterpret_cast<uint2*>(Aw + r * kKvfLd + lane * 4) = out;
    }
    __syncwarp();

    // ---- Z[16 x 128] = A . T  (mma.sync m16n8k16, fp32 accumulate, one fp16 rounding); T == null: Z = A
    const int g = lane >> 2, tig = lane & 3, mi = lane >> 3, r8 = lane & 7;
    if (T) {
        uint32_t afrag[8][4];
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
            ldsm_x4(smem_u32(Aw + (size_t)((mi & 1) * 8 + r8) * kKvfLd + kt * 16 + (mi >> 1) * 8), afrag[kt]);
        __syncwarp();   // every lane has its A fragments: the tile may be overwritten with the results
#pragma unroll
        for (int nt2 = 0; nt2 < 8; ++nt2) {
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kt = 0; kt < 8; ++kt) {
                uint32_t b[4];
                ldsm_x4_t(smem_u32(Ts + (size_t)(kt * 16 + (mi & 1) * 8 + r8) * kKvfLd + nt2 * 16 + (mi >> 1) * 8), b);
                hmma16816(c0, afrag[kt], b[0], b[1]);
                hmma16816(c1, afrag[kt], b[2], b[3]);
            }
            // rows g / g + 8, columns nt2 * 16 + {2 tig, 2 tig + 1} and + 8
            *reinterpret_cast<uint32_t*>(Aw + g * kKvfLd + nt2 * 16 + 2 * tig) = pack_h2(c0[0], c0[1]);
            *reinterpret_cast<uint32_t*>(Aw + (g + 8) * kKvfLd + nt2 * 16 + 2 * tig) = pack_h2(c0[2], c0[3]);
            *reinterpret_cast<uint32_t*>(Aw + g * kKvfLd + nt2 * 16 + 8 + 2 * tig) = pack_h2(c1[0], c1[1]);
            *reinterpret_cast<uint32_t*>(Aw + (g + 8) * kKvfLd + nt2 * 16 + 8 + 2 * tig) = pack_h2(c1[2], c1[3]);
        }
        __syncwarp();
    }

    // ---- rows back out: 4 values per lane per row (coalesced 256-byte rows)
    for (int r = 0; r < 16; ++r) {
        const long long row = row0 + r;
        if (row >= n_rows) break;
        const int token = (int)(row / heads), head = (int)(row % heads);
        const uint2 val = *reinterpret_cast<const uint2*>(Aw + r * kKvfLd + lane * 4);
        if (kind == 0) {
            *reinterpret_cast<uint2*>(p.q_out + (size_t)row * 128 + lane * 4) = val;
            continue;
        }
        if (p.k_out) *reinterpret_cast<uint2*>(p.k_out + (size_t)row * 128 + lane * 4) = val;
        const __half xv[4] = {__ushort_as_half((uint16_t)(val.x & 0xFFFF)), __ushort_as_half((uint16_t)(val.x >> 16)),
                              __ushort_as_half((uint16_t)(val.y & 0xFFFF)), __ushort_as_half((uint16_t)(val.y >> 16))};
        __half scale, zero;
        uint16_t packed;
        kvf_quant_row_warp(xv, lane, scale, zero, packed);
        const int s = p.slot[token], page = s / p.page_size, off = s % p.page_size;
        const size_t idx = ((((size_t)page * p.n_layers + p.layer) * 2 + 0) * p.kv_heads + head) * p.page_size + off;
        reinterpret_cast<uint16_t*>(p.kv_data + idx * 64)[lane] = packed;
        if (lane == 0) reinterpret_cast<__half2*>(p.kv_param)[idx] = __halves2half2(scale, zero);
    }
}

}  // namespace flatq
