// qattention side path: online Hadamard transform, 4-bit asymmetric KV-cache quantisation into a
// paged cache, and single-query (decode) attention over the packed INT4 cache.  All HBM-bound SIMT kernels.
//
//   fht_kernel             replaces third-party/fast-hadamard-transform/csrc/fast_hadamard_transform_cuda.cu:153-229
//                          (out = scale * x . H_n per row, n = 2^k): radix-8 butterflies in registers, shared-memory
//                          exchanges between passes, fp32 arithmetic, 16-byte global accesses.
//   kv_quant_kernel        replaces deploy/transformers/kv_cache.py:11-51 (`asym_quantize_and_pack_i4`, ~10 torch
//                          kernels) + AppendPagedKVCache{Prefill,Decode}Kernel (deploy/kernels/include/flashinfer/
//                          page.cuh:119-219): per (token, head): s = (max-min).clamp(1e-5)/15, zero = -min,
//                          q = clamp(rint((x+zero)/s), 0, 15) with the reference's fp16 op-by-op rounding, packed
//                          two per byte and scattered into the page table, params (s, zero) as half2.
//   decode_attn_i4_kernel  replaces BatchDecodeWithPagedKVCacheKernel (decode.cuh:480-689; dequant quantization.cuh:
//                          59-84: x_hat = nibble*s - zero): softmax(q.K^T/sqrt(d)).V, online softmax in fp32.
// Cache layout (kv_cache.py:180-193): data  [pages, layers, 2 (K/V), heads, page_size, head_dim/2] u8
//                                     param [pages, layers, 2, heads, page_size, 2] f16 (scale, zero)
#pragma once
#include "ptx.cuh"

namespace flatq {

// ------------------------------------------------------------------------------------------------ FHT
// One CTA per row.  smem holds the row as fp32 with one pad word per 32 (bank-conflict-free strided access).
__device__ __forceinline__ int fht_pad(int i) { return i + (i >> 5); }

template <int RADIX_BITS>
__device__ __forceinline__ void fht_pass(float* s, int n, int shift, int tid, int nthreads) {
    constexpr int R = 1 << RADIX_BITS;
    const int groups = n >> RADIX_BITS;
    for (int t = tid; t < groups; t += nthreads) {
        const int base = ((t >> shift) << (shift + RADIX_BITS)) | (t & ((1 << shift) - 1));
        float v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = s[fht_pad(base + (j << shift))];
#pragma unroll
        for (int h = 1; h < R; h <<= 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (!(j & h)) {
                    const float a = v[j], b = v[j | h];
                    v[j] = a + b;
                    v[j | h] = a - b;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) s[fht_pad(base + (j << shift))] = v[j];
    }
}

// `tpr` threads per row, blockDim / tpr rows per CTA (short rows -- the 128-wide K / Q head transform -- would otherwise
// leave a CTA with one warp and a serial walk over its rows: 290 us for 65536 rows in round 2's first measurement).
__global__ void __launch_bounds__(1024) fht_kernel(const __half* __restrict__ x, __half* __restrict__ out, int rows, int n,
                                                  int log_n, float scale, int tpr) {
    extern __shared__ float fht_all[];
    const int rpc = blockDim.x / tpr, sub = threadIdx.x / tpr, tid = threadIdx.x % tpr, nt = tpr;
    float* fht_s = fht_all + (size_t)sub * (n + n / 32 + 1);
    for (int row0 = blockIdx.x * rpc; row0 < rows; row0 += gridDim.x * rpc) {   // uniform trip count: barriers are CTA-wide
        const int row = row0 + sub;
        const bool ok = row < rows;
        const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)row * n);
        if (ok) {
            for (int i = tid; i < n / 8; i += nt) {
                const uint4 v = src[i];
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    fht_s[fht_pad(i * 8 + 2 * j)] = f.x;
                    fht_s[fht_pad(i * 8 + 2 * j + 1)] = f.y;
                }
            }
        }
        __syncthreads();
        int shift = 0;
        for (; shift + 3 <= log_n; shift += 3) { fht_pass<3>(fht_s, n, shift, tid, nt); __syncthreads(); }
        if (log_n - shift == 2) { fht_pass<2>(fht_s, n, shift, tid, nt); __syncthreads(); }
        if (log_n - shift == 1) { fht_pass<1>(fht_s, n, shift, tid, nt); __syncthreads(); }
        if (ok) {
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * n);
            for (int i = tid; i < n / 8; i += nt) {
                uint4 v;
                __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    h[j] = __floats2half2_rn(fht_s[fht_pad(i * 8 + 2 * j)] * scale, fht_s[fht_pad(i * 8 + 2 * j + 1)] * scale);
                dst[i] = v;
            }
        }
        __syncthreads();
    }
}

// Short rows (n = 8 LANES <= 256: the 128-wide K / Q head transform of the Hadamard baseline mode): no shared memory and
// no barrier -- a row lives in LANES lanes of a warp, 8 consecutive elements (one 16-byte access) per lane; butterflies
// over index bits 0-2 in registers, over the higher bits by shuffles.  Same butterfly order (bit 0 upwards) and the same
// fp32 operations as fht_kernel: bit-identical results.
template <int LANES>
__global__ void __launch_bounds__(256) fht_warp_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long rows, float scale) {
    constexpr int RPW = 32 / LANES;   // rows per warp
    const int lane = threadIdx.x & 31, sub = lane / LANES, l = lane % LANES;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r0 = ((((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * RPW; r0 < rows; r0 += warps * RPW) {   // warp-uniform
        const long long row = r0 + sub;
        const bool ok = row < rows;
        float v[8];
        {
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (ok) raw = reinterpret_cast<const uint4*>(x + row * (8 * LANES))[l];
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                v[2 * j] = f.x;
                v[2 * j + 1] = f.y;
            }
        }
#pragma unroll
        for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (!(j & h)) {
                    const float a = v[j], b = v[j | h];
                    v[j] = a + b;
                    v[j | h] = a - b;
                }
            }
        }
#pragma unroll
        for (int m = 1; m < LANES; m <<= 1) {
            const bool upper = (l & m) != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float o = __shfl_xor_sync(0xffffffffu, v[j], m);
                v[j] = upper ? o - v[j] : v[j] + o;   // (a + b) in the lower lane, (a - b) in the upper one
            }
        }
        if (ok) {
            uint4 o;
            __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j] * scale, v[2 * j + 1] * scale);
            reinterpret_cast<uint4*>(out + row * (8 * LANES))[l] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------ FHT, n = K * 2^m
// matmul_hadU_cuda for the sizes that are not a power of two (deploy/functional/online_trans.py:144-151; the reference
// runs hadamard_transform over the 2^m axis and then `hadK @ input`, a cuBLAS batched GEMM with an fp16 round trip
// through HBM in between; third-party/fast-hadamard-transform's own fused forms are fast_hadamard_transform_cuda.cu:
// 37-127 / 276-441, the 12N/20N/28N/40N kernels).  One CTA per row: the row lives in shared memory as fp32 [K][N2],
// radix-8 butterflies over every N2 segment, then the dense K x K factor over the K axis.  The intermediate is rounded
// to fp16 after the scale (what hadamard_transform returns to the reference's matmul), the K-sum accumulates in fp32.
//   out[k', j] = sum_k hadK[k', k] * half(scale * (x_k . H_N2)[j])
template <int RADIX_BITS>
__device__ __forceinline__ void fht_seg_pass(float* s, int segs, int log_n2, int shift, int tid, int nthreads) {
    constexpr int R = 1 << RADIX_BITS;
    const int gps = 1 << (log_n2 - RADIX_BITS);       // groups per segment
    const int groups = segs * gps;
    for (int t = tid; t < groups; t += nthreads) {
        const int seg = t >> (log_n2 - RADIX_BITS), tt = t & (gps - 1);
        const int base = (seg << log_n2) + (((tt >> shift) << (shift + RADIX_BITS)) | (tt & ((1 << shift) - 1)));
        float v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = s[fht_pad(base + (j << shift))];
#pragma unroll
        for (int h = 1; h < R; h <<= 1) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (!(j & h)) {
                    const float a = v[j], b = v[j | h];
                    v[j] = a + b;
                    v[j | h] = a - b;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) s[fht_pad(base + (j << shift))] = v[j];
    }
}

__global__ void __launch_bounds__(1024) fht_kn_kernel(const __half* __restrict__ x, __half* __restrict__ out,
                                                     const __half* __restrict__ hadK, int rows, int K, int log_n2, float scale) {
    extern __shared__ float fht_s[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n2 = 1 << log_n2, n = K * n2;
    __half* hk = reinterpret_cast<__half*>(fht_s + fht_pad(n) + 1);   // [K][K]
    for (int i = tid; i < K * K; i += nt) hk[i] = hadK[i];
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)row * n);
        for (int i = tid; i < n / 8; i += nt) {
            const uint4 v = src[i];
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                fht_s[fht_pad(i * 8 + 2 * j)] = f.x;
                fht_s[fht_pad(i * 8 + 2 * j + 1)] = f.y;
            }
        }
        __syncthreads();
        int shift = 0;
        for (; shift + 3 <= log_n2; shift += 3) { fht_seg_pass<3>(fht_s, K, log_n2, shift, tid, nt); __syncthreads(); }
        if (log_n2 - shift == 2) { fht_seg_pass<2>(fht_s, K, log_n2, shift, tid, nt); __syncthreads(); }
        if (log_n2 - shift == 1) { fht_seg_pass<1>(fht_s, K, log_n2, shift, tid, nt); __syncthreads(); }
        for (int i = tid; i < n; i += nt) fht_s[fht_pad(i)] = __half2float(__float2half_rn(fht_s[fht_pad(i)] * scale));
        __syncthreads();
        // dense factor: a work item = 4 output rows k' x 2 adjacent columns j (a warp's lanes share k': hadK reads broadcast)
        const int pairs = n2 >> 1, items = (K >> 2) * pairs;
        __half* dst = out + (size_t)row * n;
        for (int w = tid; w < items; w += nt) {
            const int kc = (w / pairs) << 2, j = (w % pairs) << 1;
            float acc[4][2] = {};
            for (int k = 0; k < K; ++k) {
                const float v0 = fht_s[fht_pad((k << log_n2) + j)], v1 = fht_s[fht_pad((k << log_n2) + j + 1)];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float hv = __half2float(hk[(kc + r) * K + k]);
                    acc[r][0] = fmaf(hv, v0, acc[r][0]);
                    acc[r][1] = fmaf(hv, v1, acc[r][1]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<__half2*>(dst + ((kc + r) << log_n2) + j) = __floats2half2_rn(acc[r][0], acc[r][1]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ KV quantise (+ append)
// One warp per (token, head) row of D = 128 halves (4 per lane).  slot[token] = page * page_size + offset gives the
// destination inside the paged cache (host/torch computes it from kv_indptr / kv_indices / last_page_offset).
struct KvQuantParams {
    const __half* x;       // [tokens, heads, D]
    int tokens, heads, D;  // D == 128
    // contiguous outputs (may be null)
    uint8_t* q_out;        // [tokens, heads, D/2]
    __half* param_out;     // [tokens, heads, 2]
    // paged outputs (may be null)
    uint8_t* kv_data;      // page layout, see file header
    __half* kv_param;
    const int* slot;       // [tokens]
    int layer, n_layers, kv_select /*0 = K, 1 = V*/, page_size;
};

__global__ void __launch_bounds__(256) kv_quant_kernel(const KvQuantParams p) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int rows = p.tokens * p.heads;
    for (int r = wg; r < rows; r += nw) {
        const uint2 raw = reinterpret_cast<const uint2*>(p.x + (size_t)r * p.D)[lane];
        const __half2 h0 = *reinterpret_cast<const __half2*>(&raw.x), h1 = *reinterpret_cast<const __half2*>(&raw.y);
        __half2 mx = __hmax2(h0, h1), mn = __hmin2(h0, h1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint32_t a = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mx), o);
            uint32_t b = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&mn), o);
            mx = __hmax2(mx, *reinterpret_cast<__half2*>(&a));
            mn = __hmin2(mn, *reinterpret_cast<__half2*>(&b));
        }
        const __half xmax = __hmax(__low2half(mx), __high2half(mx)), xmin = __hmin(__low2half(mn), __high2half(mn));
        // the reference's fp16 tensor ops, one rounding each: (xmax - xmin).clamp(min=1e-5) / 15 ; zero = -xmin
        __half range = __hsub(xmax, xmin);
        const __half eps = __float2half_rn(1e-5f);
        if (__hlt(range, eps)) range = eps;
        const __half scale = __float2half_rn(__half2float(range) / 15.0f);
        const __half zero = __hneg(xmin);
        const float fs = __half2float(scale);
        uint32_t q[4];
        const __half xv[4] = {__low2half(h0), __high2half(h0), __low2half(h1), __high2half(h1)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half t = __hadd(xv[j], zero);                                  // h(x + zero)
            const __half u = __float2half_rn(__fdiv_rn(__half2float(t), fs));      // h(t / scale)
            q[j] = (uint32_t)max(0, min(15, __half2int_rn(u)));                    // round half even, clamp
        }
        const uint16_t packed = (uint16_t)(q[0] | (q[1] << 4) | (q[2] << 8) | (q[3] << 12));  // elements 4l..4l+3
        const int token = r / p.heads, head = r % p.heads;
        if (p.q_out) reinterpret_cast<uint16_t*>(p.q_out + (size_t)r * (p.D / 2))[lane] = packed;
        if (p.param_out && lane == 0) reinterpret_cast<__half2*>(p.param_out)[r] = __halves2half2(scale, zero);
        if (p.kv_data) {
            const int s = p.slot[token], page = s / p.page_size, off = s % p.page_size;
            const size_t idx = ((((size_t)page * p.n_layers + p.layer) * 2 + p.kv_select) * p.heads + head) * p.page_size + off;
            reinterpret_cast<uint16_t*>(p.kv_data + idx * (p.D / 2))[lane] = packed;
            if (lane == 0) reinterpret_cast<__half2*>(p.kv_param)[idx] = __halves2half2(scale, zero);
        }
    }
}

// ------------------------------------------------------------------------------------------------ decode attention
// grid (batch, q_heads); 128 threads = 4 warps, each warp takes every 4th token; lane owns 4 of the 128 dims.
struct DecodeParams {
    const __half* q;         // [batch, q_heads, D]
    __half* o;               // [batch, q_heads, D]
    const uint8_t* kv_data;  // paged, see header
    const __half* kv_param;
    const int* kv_indptr;    // [batch + 1] page ranges into kv_indices
    const int* kv_indices;   // page ids
    const int* last_page_len;  // [batch] valid entries in the last page (1..page_size)
    int batch, q_heads, kv_heads, D, page_size, layer, n_layers;
    float sm_scale;
};

__global__ void __launch_bounds__(128) decode_attn_i4_kernel(const DecodeParams p) {
    const int b = blockIdx.x, hq = blockIdx.y, hk = hq / (p.q_heads / p.kv_heads);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_m[4], s_l[4], s_acc[4][128];
    const uint2 qraw = reinterpret_cast<const uint2*>(p.q + ((size_t)b * p.q_heads + hq) * p.D)[lane];
    const float2 q01 = __half22float2(*reinterpret_cast<const __half2*>(&qraw.x));
    const float2 q23 = __half22float2(*reinterpret_cast<const __half2*>(&qraw.y));
    const float qv[4] = {q01.x * p.sm_scale, q01.y * p.sm_scale, q23.x * p.sm_scale, q23.y * p.sm_scale};
    float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int pg0 = p.kv_indptr[b], pg1 = p.kv_indptr[b + 1];
    const int n_tok = (pg1 - pg0 - 1) * p.page_size + p.last_page_len[b];
    for (int t = warp; t < n_tok; t += 4) {
        const int page = p.kv_indices[pg0 + t / p.page_size], off = t % p.page_size;
        const size_t kidx = ((((size_t)page * p.n_layers + p.layer) * 2 + 0) * p.kv_heads + hk) * p.page_size + off;
        const size_t vidx = kidx + (size_t)p.kv_heads * p.page_size;
        const uint16_t kp = reinterpret_cast<const uint16_t*>(p.kv_data + kidx * (p.D / 2))[lane];
        const uint16_t vp = reinterpret_cast<const uint16_t*>(p.kv_data + vidx * (p.D / 2))[lane];
        const float2 kpar = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[kidx]);
        const float2 vpar = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[vidx]);
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) dot += qv[j] * ((float)((kp >> (4 * j)) & 0xF) * kpar.x - kpar.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float m_new = fmaxf(m, dot);
        const float corr = __expf(m - m_new), pexp = __expf(dot - m_new);
        l = l * corr + pexp;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = acc[j] * corr + pexp * ((float)((vp >> (4 * j)) & 0xF) * vpar.x - vpar.y);
        m = m_new;
    }
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
    for (int j = 0; j < 4; ++j) s_acc[warp][lane * 4 + j] = acc[j];
    __syncthreads();
    if (warp == 0) {
        float mm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
        float ll = 0.f, out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float c = s_m[w] == -INFINITY ? 0.f : __expf(s_m[w] - mm);
            ll += s_l[w] * c;
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] += s_acc[w][lane * 4 + j] * c;
        }
        const float inv = ll > 0.f ? 1.f / ll : 0.f;
        uint2 ov;
        *reinterpret_cast<__half2*>(&ov.x) = __floats2half2_rn(out[0] * inv, out[1] * inv);
        *reinterpret_cast<__half2*>(&ov.y) = __floats2half2_rn(out[2] * inv, out[3] * inv);
        reinterpret_cast<uint2*>(p.o + ((size_t)b * p.q_heads + hq) * p.D)[lane] = ov;
    }
}

// ------------------------------------------------------------------------------------------------ split-KV decode attention
// The kernel above walks a sequence with one CTA per (batch, query head), one warp per token and no loads in flight
// beyond the current token: at batch 1 that is 32 CTAs on 148 SMs, each serialising ~seq/4 dependent-load round trips,
// and with grouped-query attention every query head of a group re-reads the same K/V rows.  This variant
//   * gives a CTA ONE kv head and GH of its query heads (K/V bytes are read once per group, not once per query head),
//   * splits the sequence over `splits` CTAs (flash-decoding) so that batch x kv_heads x splits fills the chip,
//   * reads 8 bytes (16 dims) per lane: 8 lanes cover a 64-byte row, a warp covers 4 consecutive tokens = 256
//     contiguous bytes of a page per load instruction, and the loads of the next TWO chunks are in flight while the
//     current chunk's arithmetic runs (a ring of three register buffers),
//   * keeps, per 8-lane token group, an online-softmax state (m, l, acc[16 dims]) for each of the GH heads in registers;
//     the 16 groups of a CTA are merged through shared memory, the splits by decode_attn_merge_kernel.
// Partial results: part[((b * q_heads + hq) * splits + s) * 130 + {0..127: unnormalised output, 128: max, 129: sum}], fp32.
// Same arithmetic as above (x_hat = nibble * scale - zero in fp32, scores scaled by sm_scale, fp32 online softmax) up to
// the order of the fp32 sums.  Bytes per (token, kv head): 2 * (64 + 4), read once per head group of GH.
constexpr int kDecSplitThreads = 128, kDecGroups = kDecSplitThreads / 8, kDecPartStride = 130;

struct DecKvChunk { uint2 k, v; float2 kpar, vpar; };

__device__ __forceinline__ void dec_load_chunk(const DecodeParams& p, int pg0, int hk, int t, int lane8, DecKvChunk& c) {
    const int page = p.kv_indices[pg0 + t / p.page_size], off = t % p.page_size;
    const size_t kidx = ((((size_t)page * p.n_layers + p.layer) * 2 + 0) * p.kv_heads + hk) * p.page_size + off;
    const size_t vidx = kidx + (size_t)p.kv_heads * p.page_size;
    c.k = reinterpret_cast<const uint2*>(p.kv_data + kidx * (p.D / 2))[lane8];
    c.v = reinterpret_cast<const uint2*>(p.kv_data + vidx * (p.D / 2))[lane8];
    c.kpar = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[kidx]);
    c.vpar = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[vidx]);
}

// 16 nibbles of a uint2 (element 2j in the low nibble of byte j) -> x_hat[16] = nibble * scale - zero
__device__ __forceinline__ void dec_dequant16(const uint2 raw, const float2 par, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        x[j] = fmaf((float)((raw.x >> (4 * j)) & 0xF), par.x, -par.y);
        x[8 + j] = fmaf((float)((raw.y >> (4 * j)) & 0xF), par.x, -par.y);
    }
}

template <int GH>  // query heads of one kv head handled by a CTA (1, 2 or 4)
__global__ void __launch_bounds__(kDecSplitThreads)
decode_attn_i4_split_kernel(const DecodeParams p, int splits, float* __restrict__ part) {
    grid_dep_launch();   // PDL: the merge kernel becomes resident under this one
    grid_dep_wait();
    const int split = blockIdx.x, b = blockIdx.z;
    const int G = p.q_heads / p.kv_heads, head_groups = G / GH;
    const int hk = blockIdx.y / head_groups, hq0 = hk * G + (blockIdx.y % head_groups) * GH;
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;   // 16-dim slice of the row; token group 0..15
    // per token group: GH x 128 partial outputs, stored [d][lane8] (dim 16 * lane8 + d at d * 8 + lane8) and padded by 8
    // words per group, so that the 4 groups x 8 lanes of a warp write 32 distinct banks
    constexpr int kGrpStride = GH * 128 + 8;
    __shared__ float s_m[kDecGroups][GH], s_l[kDecGroups][GH];
    __shared__ float s_acc[kDecGroups * kGrpStride];

    float qv[GH][16];
#pragma unroll
    for (int g = 0; g < GH; ++g) {
        const uint4* qp = reinterpret_cast<const uint4*>(p.q + ((size_t)b * p.q_heads + hq0 + g) * p.D + 16 * lane8);
        const uint4 lo = qp[0], hi = qp[1];
        const __half2* h0 = reinterpret_cast<const __half2*>(&lo);
        const __half2* h1 = reinterpret_cast<const __half2*>(&hi);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(h0[j]), c = __half22float2(h1[j]);
            qv[g][2 * j] = a.x * p.sm_scale; qv[g][2 * j + 1] = a.y * p.sm_scale;
            qv[g][8 + 2 * j] = c.x * p.sm_scale; qv[g][8 + 2 * j + 1] = c.y * p.sm_scale;
        }
    }
    float m[GH], l[GH], acc[GH][16];
#pragma unroll
    for (int g = 0; g < GH; ++g) {
        m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[g][d] = 0.f;
    }

    const int pg0 = p.kv_indptr[b], pg1 = p.kv_indptr[b + 1];
    const int n_tok = (pg1 - pg0 - 1) * p.page_size + p.last_page_len[b];
    const int per = (n_tok + splits - 1) / splits;           // tokens per split
    const int t0 = split * per, t1 = min(n_tok, t0 + per);   // empty when t0 >= n_tok
    const int n_iter = t1 > t0 ? (t1 - t0 + kDecGroups - 1) / kDecGroups : 0;   // uniform over the CTA: shuffles stay converged
    // two chunks in flight beyond the one being consumed (a ring of three register buffers)
    DecKvChunk cur, nxt, nx2;
    cur.k = cur.v = nxt.k = nxt.v = nx2.k = nx2.v = make_uint2(0, 0);
    cur.kpar = cur.vpar = nxt.kpar = nxt.vpar = nx2.kpar = nx2.vpar = make_float2(0.f, 0.f);
    if (t0 + grp < t1) dec_load_chunk(p, pg0, hk, t0 + grp, lane8, cur);
    if (t0 + kDecGroups + grp < t1) dec_load_chunk(p, pg0, hk, t0 + kDecGroups + grp, lane8, nxt);
    for (int it = 0; it < n_iter; ++it) {
        const int t = t0 + it * kDecGroups + grp;
        const bool valid = t < t1;
        if (t + 2 * kDecGroups < t1) dec_load_chunk(p, pg0, hk, t + 2 * kDecGroups, lane8, nx2);
        float xh[16];
        dec_dequant16(cur.k, cur.kpar, xh);
        float dot[GH];
#pragma unroll
        for (int g = 0; g < GH; ++g) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) a = fmaf(qv[g][d], xh[d], a);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            dot[g] = a;
        }
        dec_dequant16(cur.v, cur.vpar, xh);
        if (valid) {
#pragma unroll
            for (int g = 0; g < GH; ++g) {
                const float m_new = fmaxf(m[g], dot[g]);
                const float corr = __expf(m[g] - m_new), pexp = __expf(dot[g] - m_new);   // m = -inf at first: corr = 0
                l[g] = l[g] * corr + pexp;
#pragma unroll
                for (int d = 0; d < 16; ++d) acc[g][d] = fmaf(acc[g][d], corr, pexp * xh[d]);
                m[g] = m_new;
            }
        }
        cur = nxt;
        nxt = nx2;
    }

    // ---- merge the CTA's 16 token groups (thread j owns output dim j), publish one partial per (query head, split)
#pragma unroll
    for (int g = 0; g < GH; ++g) {
        if (lane8 == 0) { s_m[grp][g] = m[g]; s_l[grp][g] = l[g]; }
#pragma unroll
        for (int d = 0; d < 16; ++d) s_acc[grp * kGrpStride + g * 128 + d * 8 + lane8] = acc[g][d];
    }
    __syncthreads();
    const int j = threadIdx.x, jj = (j & 15) * 8 + (j >> 4);   // where dim j sits inside a group's [d][lane8] block
#pragma unroll
    for (int g = 0; g < GH; ++g) {
        float mm = -INFINITY;
#pragma unroll
        for (int r = 0; r < kDecGroups; ++r) mm = fmaxf(mm, s_m[r][g]);
        float ll = 0.f, out = 0.f;
#pragma unroll
        for (int r = 0; r < kDecGroups; ++r) {
            const float c = s_m[r][g] == -INFINITY ? 0.f : __expf(s_m[r][g] - mm);
            ll = fmaf(s_l[r][g], c, ll);
            out = fmaf(s_acc[r * kGrpStride + g * 128 + jj], c, out);
        }
        float* dst = part + (((size_t)b * p.q_heads + hq0 + g) * splits + split) * kDecPartStride;
        dst[j] = out;
        if (j == 0) { dst[128] = mm; dst[129] = ll; }
    }
}

// grid = batch * q_heads, 128 threads (thread j = output dim j): combine the splits' partial softmax states
__global__ void __launch_bounds__(128) decode_attn_merge_kernel(const float* __restrict__ part, int splits, __half* __restrict__ o) {
    grid_dep_launch();
    grid_dep_wait();
    const float* src = part + (size_t)blockIdx.x * splits * kDecPartStride;
    const int j = threadIdx.x;
    float mm = -INFINITY;
    for (int s = 0; s < splits; ++s) mm = fmaxf(mm, src[s * kDecPartStride + 128]);
    float ll = 0.f, out = 0.f;
    for (int s = 0; s < splits; ++s) {
        const float ms = src[s * kDecPartStride + 128];
        const float c = ms == -INFINITY ? 0.f : __expf(ms - mm);
        ll = fmaf(src[s * kDecPartStride + 129], c, ll);
        out = fmaf(src[s * kDecPartStride + j], c, out);
    }
    o[(size_t)blockIdx.x * 128 + j] = __float2half_rn(ll > 0.f ? out / ll : 0.f);
}

}  // namespace flatq
