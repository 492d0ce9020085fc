// Split-KV decode attention over the paged INT4 cache on the tensor cores (round 2).
//
// Replaces BatchDecodeWithPagedKVCacheKernel (deploy/kernels/include/flashinfer/decode.cuh:480-689, bound as batch_decode_i4,
// bindings.cpp:118-170) for the grouped-query case, like decode_attn_i4_split_kernel (qattention.cuh) -- whose scalar online
// softmax (one exp and 2 x 16 FMAs per token, head and lane) is issue-bound at 155 / 440 GB/s of the KV bytes (bsz 1 / 16,
// seq 2048).  Here a CTA still owns one kv head of one sequence and one split of its tokens, but the G query heads of the
// group are the rows of an m16n8k16 tile and the arithmetic is two small GEMMs per 32-token tile and warp:
//
//   S' [16 x 32] = Q [16 x 128] . N_k^T          N_k = the K nibbles as fp16 integers 0..15 (exact)
//   S  = kscale_t S' - kzero_t qsum              x_hat = nibble * scale - zero, folded:  q . k_hat = s (q . n) - z sum(q)
//   online softmax over the tile (fp32), P = exp(S - m)
//   O [16 x 128] += (P * vscale_t) . N_v   and   c += sum_t P_t vzero_t          (o = O - c at the end)
//
// so the per-element work is one nibble -> fp16 conversion (PRMT + HADD2, two values per instruction) and the mma itself;
// scale / zero enter once per (token, head row).  The 128 dims are kept in the order the conversion produces them
// (inside every 8: 0,2,4,6,1,3,5,7): Q is staged with the same permutation (a dot product does not care) and the output
// is un-permuted when it is written.  Partial results and the merge over splits are those of the SIMT kernel
// (part[..][130], decode_attn_merge_kernel).  Same arithmetic up to fp32 summation order and the fp16 rounding of P.
//
// FRONT = true (flatq_kv_decode_fused): the kernel in front of the attention in a decode step -- kv_fused_kernel: RoPE ->
// Q . T_k^-T, K . T_k -> 4-bit quantise -> page scatter of the ONE new token per sequence -- is folded into this kernel's
// prologue, so that a decode layer has one launch less on its latency chain.  Every CTA rotates and transforms the G query
// rows of its kv head itself (the same m16n8k16 sequence as kv_fused.cuh: bit-identical q_t; T_k^-T and T_k are static and
// are copied to shared memory before the dependency wait, aliased onto the K / V tile buffers); the CTA whose split holds
// the newest token also rotates, transforms, quantises and stores that token's K row and quantises and stores its V row
// (kvf_quant_rows4: the reference's fp16 op-by-op quantiser) before it enters the token loop, where it reads them back
// like any other token.
#pragma once
#include "kron_quant.cuh"
#include "kv_fused.cuh"
#include "qattention.cuh"

namespace flatq {

constexpr int kDmWarps = 4, kDmThreads = kDmWarps * 32, kDmTile = 32, kDmLd = 128 + 8;
// per warp: K tile + V tile (fp16 nibbles) + per-token (kscale, kzero, vscale, vzero); per CTA: Q tile [16][kDmLd] + qsum[16]
constexpr int kDmWarpBytes = 2 * kDmTile * kDmLd * 2 + 4 * kDmTile * 4;
constexpr int kDmSmem = kDmWarps * kDmWarpBytes + 16 * kDmLd * 2 + 16 * 4 + kDmWarps * 16 * 2 * 4 /* m, l */ + 16;

// 8 nibbles of a word -> four half2 in the order (n0,n2) (n4,n6) (n1,n3) (n5,n7), each exactly the integer value
__device__ __forceinline__ void dm_nibbles8(uint32_t w, uint32_t (&h)[4]) {
    const uint32_t lo = w & 0x0F0F0F0Fu, hi = (w >> 4) & 0x0F0F0F0Fu;
    const uint32_t magic = 0x64646464u;                       // fp16 0x64xx = 1024 + xx
    const __half2 k1024 = __float2half2_rn(1024.f);
    uint32_t t[4] = {__byte_perm(lo, magic, 0x5140), __byte_perm(lo, magic, 0x7362), __byte_perm(hi, magic, 0x5140),
                     __byte_perm(hi, magic, 0x7362)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&t[i]), k1024);
        h[i] = *reinterpret_cast<const uint32_t*>(&v);
    }
}
// position of original dim d (0..127) in the permuted order
__host__ __device__ constexpr int dm_perm(int d) { return (d & ~7) + ((d & 1) << 2) + ((d & 7) >> 1); }

struct DecodeFront {         // the new token's pre-RoPE rows (strided views of the qkv GEMM output) and what kv_fused needs
    const __half* q; const __half* k; const __half* v;   // row (sequence b, head h) at X + b * ld_X + h * 128
    long long ld_q, ld_k, ld_v;
    const __half* cos_t; const __half* sin_t;            // [positions, 128]
    const int* pos;                                      // [batch] position of the new token
    const __half* t_q; const __half* t_k;                // [128, 128] row-major: q_t = q . t_q, k_t = k . t_k
    const int* slot;                                     // [batch] cache slot (page * page_size + offset) of the new token
    uint8_t* kv_data; __half* kv_param;                  // the cache, writable
};

// one row (128 halves, 4 per lane) through HF's rotate_half RoPE -- the arithmetic of kv_fused_kernel, phase 1
__device__ __forceinline__ uint2 dm_rope_row(uint2 raw, uint2 cs, uint2 sn, int lane) {
    uint2 other, out;
    other.x = __shfl_xor_sync(0xffffffffu, raw.x, 16);
    other.y = __shfl_xor_sync(0xffffffffu, raw.y, 16);
    const __half2* xh = reinterpret_cast<const __half2*>(&raw);
    const __half2* oh = reinterpret_cast<const __half2*>(&other);
    const __half2* ch = reinterpret_cast<const __half2*>(&cs);
    const __half2* sh = reinterpret_cast<const __half2*>(&sn);
    __half2* rh = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const __half2 rot = lane < 16 ? __hneg2(oh[j]) : oh[j];
        rh[j] = __hadd2_rn(__hmul2_rn(xh[j], ch[j]), __hmul2_rn(rot, sh[j]));
    }
    return out;
}

template <bool FRONT>
__global__ void __launch_bounds__(kDmThreads)
decode_attn_i4_mma_kernel(const DecodeParams p, const DecodeFront f, int splits, float* __restrict__ part) {
    extern __shared__ __align__(16) uint8_t dm_smem[];
    grid_dep_launch();   // PDL: the merge kernel becomes resident under this one
    if (FRONT) {         // static transform matrices -> shared memory (the tile buffers are not in use yet)
        static_assert(2 * 128 * kDmLd * 2 <= kDmWarps * kDmWarpBytes, "T_q and T_k alias the K / V tile buffers");
        __half* Tq = reinterpret_cast<__half*>(dm_smem);
        __half* Tk = Tq + 128 * kDmLd;
#pragma unroll 1
        for (int i0 = threadIdx.x; i0 < 128 * 16; i0 += 8 * kDmThreads) {
            uint4 a[8], c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] = __ldg(reinterpret_cast<const uint4*>(f.t_q) + i0 + u * kDmThreads);
                c[u] = __ldg(reinterpret_cast<const uint4*>(f.t_k) + i0 + u * kDmThreads);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kDmThreads, r = i >> 4, col = (i & 15) * 8;
                *reinterpret_cast<uint4*>(Tq + r * kDmLd + col) = a[u];
                *reinterpret_cast<uint4*>(Tk + r * kDmLd + col) = c[u];
            }
        }
    }
    grid_dep_wait();     // q (and, without FRONT, this step's cache rows) come from the kernel in front
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int G = p.q_heads / p.kv_heads, hq0 = hk * G;       // G <= 16 (host)
    __half* Ks = reinterpret_cast<__half*>(dm_smem + warp * kDmWarpBytes);
    __half* Vs = Ks + kDmTile * kDmLd;
    float* par = reinterpret_cast<float*>(Vs + kDmTile * kDmLd);     // [4][32]: kscale, kzero, vscale, vzero
    __half* Qs = reinterpret_cast<__half*>(dm_smem + kDmWarps * kDmWarpBytes);   // [16][kDmLd], permuted dims, x sm_scale
    float* qsum = reinterpret_cast<float*>(Qs + 16 * kDmLd);                      // [16] sum over dims of the scaled q row
    float* s_ml = qsum + 16;                                                      // [warps][16][2]

    const int pg0 = p.kv_indptr[b], pg1 = p.kv_indptr[b + 1];
    const int n_tok = (pg1 - pg0 - 1) * p.page_size + p.last_page_len[b];
    const int per = (n_tok + splits - 1) / splits;
    const int t0 = split * per, t1 = min(n_tok, t0 + per);

    if (FRONT) {
        const __half* Tq = reinterpret_cast<const __half*>(dm_smem);
        const __half* Tk = Tq + 128 * kDmLd;
        const int gq = lane >> 2, tq = lane & 3, mq = lane >> 3, rq = lane & 7;
        const int ps = f.pos[b];
        const uint2 cs = __ldg(reinterpret_cast<const uint2*>(f.cos_t + (size_t)ps * 128 + lane * 4));
        const uint2 sn = __ldg(reinterpret_cast<const uint2*>(f.sin_t + (size_t)ps * 128 + lane * 4));
        // ---- the group's query rows: RoPE -> Qs (natural dim order; rows >= G zero)
        for (int r = warp; r < 16; r += kDmWarps) {
            uint2 out = make_uint2(0, 0);
            if (r < G) out = dm_rope_row(*reinterpret_cast<const uint2*>(f.q + (size_t)b * f.ld_q + (size_t)(hq0 + r) * 128 + lane * 4), cs, sn, lane);
            *reinterpret_cast<uint2*>(Qs + r * kDmLd + lane * 4) = out;
        }
        __syncthreads();
        uint32_t af[8][4];
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
            ldsm_x4(smem_u32(Qs + (size_t)((mq & 1) * 8 + rq) * kDmLd + kt * 16 + (mq >> 1) * 8), af[kt]);
        __syncthreads();   // every warp holds its A fragments: Qs may be reused
        // ---- q_t = q . T_q: warp w owns output columns [32 w, 32 w + 32) (k order as kv_fused_kernel: bit-identical)
        uint32_t qt[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int nt2 = 2 * warp + i;
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kt = 0; kt < 8; ++kt) {
                uint32_t bb[4];
                ldsm_x4_t(smem_u32(Tq + (size_t)(kt * 16 + (mq & 1) * 8 + rq) * kDmLd + nt2 * 16 + (mq >> 1) * 8), bb);
                hmma16816(c0, af[kt], bb[0], bb[1]);
                hmma16816(c1, af[kt], bb[2], bb[3]);
            }
            qt[i][0] = pack_h2(c0[0], c0[1]); qt[i][1] = pack_h2(c0[2], c0[3]);
            qt[i][2] = pack_h2(c1[0], c1[1]); qt[i][3] = pack_h2(c1[2], c1[3]);
        }
        // ---- the CTA whose split holds the newest token writes that token's K and V rows into the cache
        if (t0 <= n_tok - 1 && n_tok - 1 < t1) {   // CTA-uniform
            const int sl = f.slot[b], page = sl / p.page_size, off = sl - page * p.page_size;
            const size_t kidx = ((((size_t)page * p.n_layers + p.layer) * 2 + 0) * p.kv_heads + hk) * p.page_size + off;
            const size_t vidx = kidx + (size_t)p.kv_heads * p.page_size;
            if (warp == 0) {
                const uint2 kr0 = dm_rope_row(*reinterpret_cast<const uint2*>(f.k + (size_t)b * f.ld_k + (size_t)hk * 128 + lane * 4), cs, sn, lane);
                *reinterpret_cast<uint2*>(Qs + lane * 4) = kr0;   // row 0 of the A tile; rows 1..15 hold finite leftovers (rows are independent)
            } else if (warp == 1) {
                uint2 raw4[4] = {*reinterpret_cast<const uint2*>(f.v + (size_t)b * f.ld_v + (size_t)hk * 128 + lane * 4), make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
                __half sc4[4], ze4[4];
                uint16_t pk4[4];
                kvf_quant_rows4(raw4, sc4, ze4, pk4);
                reinterpret_cast<uint16_t*>(f.kv_data + vidx * 64)[lane] = pk4[0];
                if (lane == 0) reinterpret_cast<__half2*>(f.kv_param)[vidx] = __halves2half2(sc4[0], ze4[0]);
            }
            __syncthreads();
            uint32_t kf[8][4];
#pragma unroll
            for (int kt = 0; kt < 8; ++kt)
                ldsm_x4(smem_u32(Qs + (size_t)((mq & 1) * 8 + rq) * kDmLd + kt * 16 + (mq >> 1) * 8), kf[kt]);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int nt2 = 2 * warp + i;
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    uint32_t bb[4];
                    ldsm_x4_t(smem_u32(Tk + (size_t)(kt * 16 + (mq & 1) * 8 + rq) * kDmLd + nt2 * 16 + (mq >> 1) * 8), bb);
                    hmma16816(c0, kf[kt], bb[0], bb[1]);
                    hmma16816(c1, kf[kt], bb[2], bb[3]);
                }
                if (gq == 0) {   // row 0 of the tile = the K row
                    *reinterpret_cast<uint32_t*>(Qs + nt2 * 16 + 2 * tq) = pack_h2(c0[0], c0[1]);
                    *reinterpret_cast<uint32_t*>(Qs + nt2 * 16 + 8 + 2 * tq) = pack_h2(c1[0], c1[1]);
                }
            }
            __syncthreads();
            if (warp == 0) {
                uint2 raw4[4] = {*reinterpret_cast<const uint2*>(Qs + lane * 4), make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
                __half sc4[4], ze4[4];
                uint16_t pk4[4];
                kvf_quant_rows4(raw4, sc4, ze4, pk4);
                reinterpret_cast<uint16_t*>(f.kv_data + kidx * 64)[lane] = pk4[0];
                if (lane == 0) reinterpret_cast<__half2*>(f.kv_param)[kidx] = __halves2half2(sc4[0], ze4[0]);
            }
            __threadfence();   // the token loop of this CTA reads the two rows back through global memory
            __syncthreads();
        }
        // ---- q_t -> Qs, scaled and in the permuted dim order (what the plain prologue does with the q_t it is handed)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = gq + (j & 1) * 8, col = (2 * warp + i) * 16 + (j >> 1) * 8 + 2 * tq;
                const float2 v2 = __half22float2(*reinterpret_cast<const __half2*>(&qt[i][j]));
                Qs[row * kDmLd + dm_perm(col)] = __float2half_rn(v2.x * p.sm_scale);
                Qs[row * kDmLd + dm_perm(col + 1)] = __float2half_rn(v2.y * p.sm_scale);
            }
        }
    } else {
        // ---- Q rows of the group -> shared (fp16, scaled, permuted); rows >= G are zero
        for (int i = threadIdx.x; i < 16 * 16; i += kDmThreads) {
            const int r = i >> 4, d0 = (i & 15) * 8;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (r < G) v = *reinterpret_cast<const uint4*>(p.q + ((size_t)b * p.q_heads + hq0 + r) * 128 + d0);
            const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
            for (int u = 0; u < 8; ++u) Qs[r * kDmLd + dm_perm(d0 + u)] = __float2half_rn(__half2float(h[u]) * p.sm_scale);
        }
    }
    __syncthreads();
    for (int r = warp; r < 16; r += kDmWarps) {   // sum of the ROUNDED q values: the same numbers the mma multiplies
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) s += __half2float(Qs[r * kDmLd + lane * 4 + u]);
#pragma unroll
        for (int o_ = 16; o_ > 0; o_ >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o_);
        if (lane == 0) qsum[r] = s;
    }
    __syncthreads();

    const int g = lane >> 2, tig = lane & 3, mi = lane >> 3, r8 = lane & 7;
    uint32_t qf[8][4];
#pragma unroll
    for (int kt = 0; kt < 8; ++kt)
        ldsm_x4(smem_u32(Qs + (size_t)((mi & 1) * 8 + r8) * kDmLd + kt * 16 + (mi >> 1) * 8), qf[kt]);
    const float qs0 = qsum[g], qs1 = qsum[g + 8];

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, cz0 = 0.f, cz1 = 0.f;   // rows g and g + 8 (partial over tig)

    // raw nibbles + params of one 32-token tile: lane = (token in a group of 8, 16-byte quarter of the row), 4 passes
    uint4 kr[4], vr[4];
    float4 prm[4];
    auto load_tile = [&](int tb) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int t = tb + it * 8 + (lane >> 2);
            kr[it] = vr[it] = make_uint4(0, 0, 0, 0);
            prm[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < t1) {
                const int page = p.kv_indices[pg0 + t / p.page_size], off = t % p.page_size;
                const size_t kidx = ((((size_t)page * p.n_layers + p.layer) * 2 + 0) * p.kv_heads + hk) * p.page_size + off;
                const size_t vidx = kidx + (size_t)p.kv_heads * p.page_size;
                kr[it] = reinterpret_cast<const uint4*>(p.kv_data + kidx * 64)[lane & 3];
                vr[it] = reinterpret_cast<const uint4*>(p.kv_data + vidx * 64)[lane & 3];
                if ((lane & 3) == 0) {
                    const float2 kp = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[kidx]);
                    const float2 vp = __half22float2(reinterpret_cast<const __half2*>(p.kv_param)[vidx]);
                    prm[it] = make_float4(kp.x, kp.y, vp.x, vp.y);
                }
            }
        }
    };
    if (t0 + warp * kDmTile < t1) load_tile(t0 + warp * kDmTile);
    for (int tb = t0 + warp * kDmTile; tb < t1; tb += kDmWarps * kDmTile) {
        __syncwarp();   // the previous tile's ldmatrix reads are done before the tiles are overwritten
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int tl = it * 8 + (lane >> 2), c0 = (lane & 3) * 32;
            const uint32_t kw[4] = {kr[it].x, kr[it].y, kr[it].z, kr[it].w}, vw[4] = {vr[it].x, vr[it].y, vr[it].z, vr[it].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t h[4];
                dm_nibbles8(kw[w], h);
                *reinterpret_cast<uint4*>(Ks + tl * kDmLd + c0 + w * 8) = make_uint4(h[0], h[1], h[2], h[3]);
                dm_nibbles8(vw[w], h);
                *reinterpret_cast<uint4*>(Vs + tl * kDmLd + c0 + w * 8) = make_uint4(h[0], h[1], h[2], h[3]);
            }
            if ((lane & 3) == 0) {
                par[tl] = prm[it].x; par[32 + tl] = prm[it].y; par[64 + tl] = prm[it].z; par[96 + tl] = prm[it].w;
            }
        }
        __syncwarp();
        // the next tile's global loads fly under this tile's mma work (the registers were consumed by the conversion)
        if (tb + kDmWarps * kDmTile < t1) load_tile(tb + kDmWarps * kDmTile);

        // ---- S' = Q . N_k^T : 4 n-tiles of 8 tokens
        float sc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
#pragma unroll
            for (int np = 0; np < 2; ++np) {   // a pair of n-tiles per ldmatrix.x4
                uint32_t bf[4];
                ldsm_x4(smem_u32(Ks + (size_t)(np * 16 + (mi >> 1) * 8 + r8) * kDmLd + kt * 16 + (mi & 1) * 8), bf);
                hmma16816(sc[2 * np], qf[kt], bf[0], bf[1]);
                hmma16816(sc[2 * np + 1], qf[kt], bf[2], bf[3]);
            }
        }
        // ---- S = kscale S' - kzero qsum, masked; tile max per row
        float tmax0 = -INFINITY, tmax1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tl = nt * 8 + 2 * tig + e;
                const float ks = par[tl], kz = par[32 + tl];
                const bool ok = tb + tl < t1;
                sc[nt][e] = ok ? fmaf(ks, sc[nt][e], -kz * qs0) : -INFINITY;
                sc[nt][2 + e] = ok ? fmaf(ks, sc[nt][2 + e], -kz * qs1) : -INFINITY;
                tmax0 = fmaxf(tmax0, sc[nt][e]);
                tmax1 = fmaxf(tmax1, sc[nt][2 + e]);
            }
        }
        tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 1));
        tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 2));
        tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 1));
        tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 2));
        const float mn0 = fmaxf(m0, tmax0), mn1 = fmaxf(m1, tmax1);      // finite: every tile holds >= 1 valid token
        const float corr0 = __expf(m0 - mn0), corr1 = __expf(m1 - mn1);  // exp(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1;
        l0 *= corr0; l1 *= corr1; cz0 *= corr0; cz1 *= corr1;
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
        // ---- P = exp(S - m); A fragments of (P * vscale); l and the zero-point term
        uint32_t pa[2][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float pv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tl = nt * 8 + 2 * tig + e;
                const float vs = par[64 + tl], vz = par[96 + tl];
                const float p0 = __expf(sc[nt][e] - m0), p1 = __expf(sc[nt][2 + e] - m1);   // masked: exp(-inf) = 0
                l0 += p0; l1 += p1;
                cz0 = fmaf(p0, vz, cz0); cz1 = fmaf(p1, vz, cz1);
                pv[e] = p0 * vs; pv[2 + e] = p1 * vs;
            }
            // C layout of n-tile nt -> A layout of k-step nt / 2: a0/a1 (k 0..7) from even nt, a2/a3 (k 8..15) from odd nt
            pa[nt >> 1][(nt & 1) * 2] = pack_h2(pv[0], pv[1]);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(pv[2], pv[3]);
        }
        // ---- O += (P vscale) . N_v : 2 k-steps of 16 tokens x 16 n-tiles of 8 dims
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                uint32_t bf[4];
                ldsm_x4_t(smem_u32(Vs + (size_t)(ks * 16 + (mi & 1) * 8 + r8) * kDmLd + np * 16 + (mi >> 1) * 8), bf);
                hmma16816(o[2 * np], pa[ks], bf[0], bf[1]);
                hmma16816(o[2 * np + 1], pa[ks], bf[2], bf[3]);
            }
        }
    }

    // ---- the warp's state -> shared (reusing its K / V tiles as fp32 [16][128] + the (m, l) table), then the CTA merge
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    cz0 += __shfl_xor_sync(0xffffffffu, cz0, 1); cz0 += __shfl_xor_sync(0xffffffffu, cz0, 2);
    cz1 += __shfl_xor_sync(0xffffffffu, cz1, 1); cz1 += __shfl_xor_sync(0xffffffffu, cz1, 2);
    __syncwarp();
    float* ow = reinterpret_cast<float*>(Ks);   // 16 x 128 fp32 = 8 KB <= the K tile (8.7 KB)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int dpos = nt * 8 + 2 * tig + e;              // permuted position -> original dim
            const int d = (dpos & ~7) + (((dpos & 7) & 3) << 1) + ((dpos & 7) >> 2);
            ow[g * 128 + d] = o[nt][e] - cz0;
            ow[(g + 8) * 128 + d] = o[nt][2 + e] - cz1;
        }
    }
    if (tig == 0) {
        s_ml[(warp * 16 + g) * 2] = m0; s_ml[(warp * 16 + g) * 2 + 1] = l0;
        s_ml[(warp * 16 + g + 8) * 2] = m1; s_ml[(warp * 16 + g + 8) * 2 + 1] = l1;
    }
    __syncthreads();
    const int j = threadIdx.x;   // output dim
    for (int r = 0; r < G; ++r) {
        float mm = -INFINITY;
#pragma unroll
        for (int w = 0; w < kDmWarps; ++w) mm = fmaxf(mm, s_ml[(w * 16 + r) * 2]);
        float ll = 0.f, out = 0.f;
#pragma unroll
        for (int w = 0; w < kDmWarps; ++w) {
            const float mw = s_ml[(w * 16 + r) * 2];
            const float c = mw == -INFINITY ? 0.f : __expf(mw - mm);
            ll = fmaf(s_ml[(w * 16 + r) * 2 + 1], c, ll);
            out = fmaf(reinterpret_cast<const float*>(dm_smem + w * kDmWarpBytes)[r * 128 + j], c, out);
        }
        float* dst = part + (((size_t)b * p.q_heads + hq0 + r) * splits + split) * kDecPartStride;
        dst[j] = out;
        if (j == 0) { dst[128] = mm; dst[129] = ll; }
    }
}

}  // namespace flatq
