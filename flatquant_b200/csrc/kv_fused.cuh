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
    int rows_per_cta;                                    // 16 rows per warp x 2 / 4 / 8 warps (decode steps: few rows, many small CTAs)
};

constexpr int kKvfRows = 128, kKvfThreads = 256, kKvfLd = 128 + 8;   // smem row stride in halves (ldmatrix bank spread)
constexpr int kKvfSmem = (128 * kKvfLd + 8 * 16 * kKvfLd) * 2;   // T + per-warp A tiles

// The reference's per-row asymmetric quantiser (kv_quant_kernel of qattention.cuh, bit-exact: fp16 op-by-op rounding) for
// FOUR rows of a warp at once, 4 consecutive values per lane and row: the four max / min butterflies are independent, so
// their shuffles overlap; rows are packed two per half2 register for the butterfly.
__device__ __forceinline__ void kvf_quant_rows4(const uint2 (&raw)[4], __half (&scale)[4], __half (&zero)[4], uint16_t (&packed)[4]) {
    __half mxs[4], mns[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const __half2 h0 = *reinterpret_cast<const __half2*>(&raw[r].x), h1 = *reinterpret_cast<const __half2*>(&raw[r].y);
        const __half2 a = __hmax2(h0, h1), b = __hmin2(h0, h1);
        mxs[r] = __hmax(__low2half(a), __high2half(a));
        mns[r] = __hmin(__low2half(b), __high2half(b));
    }
    __half2 mx[2] = {__halves2half2(mxs[0], mxs[1]), __halves2half2(mxs[2], mxs[3])};
    __half2 mn[2] = {__halves2half2(mns[0], mns[1]), __halves2half2(mns[2], mns[3])};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t a = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mx[i]), o);
            const uint32_t b = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mn[i]), o);
            mx[i] = __hmax2(mx[i], *reinterpret_cast<const __half2*>(&a));
            mn[i] = __hmin2(mn[i], *reinterpret_cast<const __half2*>(&b));
        }
    }
    const __half eps = __float2half_rn(1e-5f);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const __half xmax = (r & 1) ? __high2half(mx[r >> 1]) : __low2half(mx[r >> 1]);
        const __half xmin = (r & 1) ? __high2half(mn[r >> 1]) : __low2half(mn[r >> 1]);
        __half range = __hsub(xmax, xmin);
        if (__hlt(range, eps)) range = eps;
        scale[r] = __float2half_rn(__half2float(range) / 15.0f);
        zero[r] = __hneg(xmin);
        const float fs = __half2float(scale[r]), inv = __frcp_rn(fs);
        const __half xv[4] = {__ushort_as_half((uint16_t)(raw[r].x & 0xFFFF)), __ushort_as_half((uint16_t)(raw[r].x >> 16)),
                              __ushort_as_half((uint16_t)(raw[r].y & 0xFFFF)), __ushort_as_half((uint16_t)(raw[r].y >> 16))};
        uint32_t qv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = __half2float(__hadd(xv[j], zero[r]));
            // t / fs, correctly rounded (residual-corrected reciprocal: the recurrence the division routine itself uses;
            // operands are normal fp16 values, no over/underflow) at 3 FMA-pipe instructions instead of a division call
            const float q0 = t * inv;
            const __half u = __float2half_rn(fmaf(fmaf(-fs, q0, t), inv, q0));
            qv[j] = (uint32_t)max(0, min(15, __half2int_rn(u)));
        }
        packed[r] = (uint16_t)(qv[0] | (qv[1] << 4) | (qv[2] << 8) | (qv[3] << 12));
    }
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
    const int n_rows = p.tokens * heads;   // < 2^31 (checked on the host)
    const __half* src = kind == 0 ? p.q : (kind == 1 ? p.k : p.v);
    const long long ld = kind == 0 ? p.ld_q : (kind == 1 ? p.ld_k : p.ld_v);
    const __half* T = kind == 0 ? p.t_q : (kind == 1 ? p.t_k : nullptr);

    grid_dep_launch();   // PDL trigger first: the attention kernel behind this one becomes resident early
    if (T) {   // static weights: before the dependency wait; the thread's 8 loads are issued before the first store
#pragma unroll 1
        for (int i0 = threadIdx.x; i0 < 128 * 16; i0 += 8 * blockDim.x) {   // blockDim = 64 / 128 / 256: whole batches of 8
            uint4 tv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) tv[u] = __ldg(reinterpret_cast<const uint4*>(T) + i0 + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * blockDim.x, r = i >> 4, c = (i & 15) * 8;
                *reinterpret_cast<uint4*>(Ts + r * kKvfLd + c) = tv[u];
            }
        }
    }
    grid_dep_wait();
    __syncthreads();

    // The warp's 16 rows are walked in groups of four (loads of a group in flight together, shuffle chains of its rows
    // overlapping); the loops over the groups are NOT unrolled: the first version of this kernel was 11 k straight-line
    // instructions and was bound by instruction fetch (ncu: 32 % of the stalls "no instruction").
    const int row0 = blk * p.rows_per_cta + warp * 16;   // rows_per_cta = 16 x the CTA's warps
    __half* Aw = As + warp * 16 * kKvfLd;
    const int t_first = row0 / heads, h_first = row0 - t_first * heads;   // (token, head) advance by increment

    // ---- phase 1: rows -> (RoPE) -> the warp's tile in shared memory
    {
        int t = t_first, h = h_first;
#pragma unroll 1
        for (int r0 = 0; r0 < 16; r0 += 4) {
            uint2 raw[4], cs[4], sn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                raw[u] = cs[u] = sn[u] = make_uint2(0, 0);
                if (row0 + r0 + u < n_rows) {
                    raw[u] = *reinterpret_cast<const uint2*>(src + t * ld + h * 128 + lane * 4);
                    if (kind != 2 && p.cos_t) {
                        const int ps = p.pos[t];
                        cs[u] = __ldg(reinterpret_cast<const uint2*>(p.cos_t + (size_t)ps * 128 + lane * 4));
                        sn[u] = __ldg(reinterpret_cast<const uint2*>(p.sin_t + (size_t)ps * 128 + lane * 4));
                    }
                }
                if (++h == heads) { h = 0; ++t; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint2 out = raw[u];
                if (kind != 2 && p.cos_t) {
                    // rotate_half: lanes 0-15 hold dims [0, 64) and pair with lane + 16 (dims + 64): rot = -x_hi ; lanes 16-31: rot = x_lo
                    uint2 other;
                    other.x = __shfl_xor_sync(0xffffffffu, raw[u].x, 16);
                    other.y = __shfl_xor_sync(0xffffffffu, raw[u].y, 16);
                    const __half2* xh = reinterpret_cast<const __half2*>(&raw[u]);
                    const __half2* oh = reinterpret_cast<const __half2*>(&other);
                    const __half2* ch = reinterpret_cast<const __half2*>(&cs[u]);
                    const __half2* sh = reinterpret_cast<const __half2*>(&sn[u]);
                    __half2* rh = reinterpret_cast<__half2*>(&out);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const __half2 rot = lane < 16 ? __hneg2(oh[j]) : oh[j];
                        rh[j] = __hadd2_rn(__hmul2_rn(xh[j], ch[j]), __hmul2_rn(rot, sh[j]));   // _rn: never contracted into HFMA2
                    }
                }
                *reinterpret_cast<uint2*>(Aw + (r0 + u) * kKvfLd + lane * 4) = out;
            }
        }
    }
    __syncwarp();

    // ---- phase 2: Z[16 x 128] = A . T  (mma.sync m16n8k16, fp32 accumulate, one fp16 rounding); T == null: Z = A
    const int g = lane >> 2, tig = lane & 3, mi = lane >> 3, r8 = lane & 7;
    if (T) {
        uint32_t afrag[8][4];
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
            ldsm_x4(smem_u32(Aw + (size_t)((mi & 1) * 8 + r8) * kKvfLd + kt * 16 + (mi >> 1) * 8), afrag[kt]);
        __syncwarp();   // every lane has its A fragments: the tile may be overwritten with the results
#pragma unroll 2
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

    // ---- phase 3: rows back out, 4 values per lane per row (coalesced 256-byte rows); k / v: quantise + page scatter
    {
        int t = t_first, h = h_first;
#pragma unroll 1
        for (int r0 = 0; r0 < 16; r0 += 4) {
            uint2 val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) val[u] = *reinterpret_cast<const uint2*>(Aw + (r0 + u) * kKvfLd + lane * 4);
            __half scale[4], zero[4];
            uint16_t packed[4];
            if (kind != 0) kvf_quant_rows4(val, scale, zero, packed);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = row0 + r0 + u;
                if (row < n_rows) {
                    if (kind == 0) {
                        *reinterpret_cast<uint2*>(p.q_out + (size_t)row * 128 + lane * 4) = val[u];
                    } else {
                        if (kind == 1 && p.k_out) *reinterpret_cast<uint2*>(p.k_out + (size_t)row * 128 + lane * 4) = val[u];
                        const int s = p.slot[t], page = s / p.page_size, off = s - page * p.page_size;
                        const size_t idx = ((((size_t)page * p.n_layers + p.layer) * 2 + (kind - 1)) * p.kv_heads + h) * p.page_size + off;
                        reinterpret_cast<uint16_t*>(p.kv_data + idx * 64)[lane] = packed[u];
                        if (lane == 0) reinterpret_cast<__half2*>(p.kv_param)[idx] = __halves2half2(scale[u], zero[u]);
                    }
                }
                if (++h == heads) { h = 0; ++t; }
            }
        }
    }
}

}  // namespace flatq
