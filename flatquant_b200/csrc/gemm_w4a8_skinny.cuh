// Decode-batch FlatLinear GEMM: y[m, n] = (sum_k W[n,k] X[m,k]) * s_x[m] * s_w[n] (+bias) for 3 <= M <= 32 tokens.
//
// At these sizes the op is a pure stream of the packed 4-bit weight (4 M op per weight byte: 64 op/B at 16 tokens, an
// order of magnitude under the ridge), and the tcgen05 kernel is the wrong shape for it: a 128-row CTA tile leaves N / 128 = 32..224
// CTAs each walking K through a few-stage TMA ring -- too few bytes in flight to fill HBM -- and the split-K variant
// needs a memset, int32 atomics and a finalize launch (measured, 16 tokens: 16-26 us per linear whose weights stream in
// 2-9 us).  This kernel is organised like the GEMV instead -- no shared-memory operand pipeline, every byte of the
// weight requested as early as possible -- with the integer tensor-core instruction doing the arithmetic:
//   * a CTA owns 16 weight rows at a time (persistent grid over the row tiles); its 8 warps split K in interleaved 128-element super-steps (warp w: steps w, w+8, ...),
//     so a row's 8 x 64 B reads of a round are contiguous;
//   * per super-step a lane (g = lane / 4, t = lane % 4) loads ONE 16-byte chunk of row g and of row g + 8 (4 lanes cover
//     64 contiguous bytes of a row) and expands its four packed words exactly as the other kernels do
//     ((p << 4) & 0xF0F0F0F0, p & 0xF0F0F0F0 -> 16 w as int8 x 4); word i of the chunk is the A fragment of
//     mma.sync.m16n8k32.s8 number i.  The K order inside an MMA is therefore "lane t contributes the 8 codes of word i
//     of ITS chunk" -- any order works as long as B agrees, and B (the K-permuted int8 activation codes, read straight
//     from L2: 32 contiguous bytes per lane and 8-token block) is indexed to match;
//   * the first super-steps' weights are requested BEFORE griddepcontrol.wait (static data), so under PDL the stream
//     starts while the transform + quantise kernel in front is still running;
//   * the 8 warps' int32 partial tiles meet in shared memory; thread e of the CTA finishes output (row e % 16, token
//     e / 16): exact >> 4, fp32 scale product, bias, one rounding -- the epilogue of gemm_w4a8.cuh, bit for bit (integer
//     accumulation is order-independent).
//   * SILU: rows interleaved [gate_j, up_j]; neighbouring threads of the epilogue exchange and store
//     half(silu_f32(half(g))) * half(u) into an [M, N / 2] output (flatq_gemm_w4a8_silu).
// Replaces the same reference calls as gemm_w4a8.cuh (deploy/kernels/gemm.cu:7-46 + quant.cu:66-101) at decode batches.
#pragma once
#include "gemv_w4a8.cuh"

namespace flatq {

constexpr int kSkinnyWarps = 8;
constexpr int kSkinnyRows = 16;
constexpr int kSkinnyPre = 4;      // super-steps whose weights are requested before the dependency wait

__device__ __forceinline__ void imma16832(int32_t (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NB = 8-token blocks (M <= 8 NB).  XS: the activation codes are staged once per (persistent) CTA in shared memory --
// rows padded by 16 bytes so that the quarter-warp's eight 16-byte reads (2 tokens x 4 K chunks) fall into distinct bank
// groups -- instead of being re-read from L2 in every super-step of every row tile (whose latency, four dependent round
// trips per tile, was what bounded the first version: 22 us for the 58.7 MB gate/up weight).  Used whenever
// 8 NB (K + 16) bytes fit (host: skinny_xs_bytes).
template <int NB>
__host__ __device__ constexpr int skinny_red_bytes() { return kSkinnyWarps * NB * 8 * (kSkinnyRows + 1) * 4; }

template <int NB, bool SILU, bool XS>
__global__ void __launch_bounds__(kSkinnyWarps * 32, 2) gemm_w4a8_skinny_kernel(const GemvParams p) {
    extern __shared__ __align__(16) uint8_t sk_smem[];
    int32_t (*red)[NB * 8][kSkinnyRows + 1] = reinterpret_cast<int32_t (*)[NB * 8][kSkinnyRows + 1]>(sk_smem);
    uint8_t* xs = sk_smem + skinny_red_bytes<NB>();
    const int xs_ld = p.K + 16;
    grid_dep_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int steps = p.K / 128, tiles = (p.N + kSkinnyRows - 1) / kSkinnyRows;

    uint4 wa[kSkinnyPre], wb[kSkinnyPre];
    auto load_w = [&](int tile, int s0) {   // ragged last tile: clamped loads, masked stores
        const int r_lo = min(tile * kSkinnyRows + g, p.N - 1), r_hi = min(tile * kSkinnyRows + g + 8, p.N - 1);
        const uint8_t* w_lo = p.w + (size_t)r_lo * (p.K / 2) + t * 16;
        const uint8_t* w_hi = p.w + (size_t)r_hi * (p.K / 2) + t * 16;
#pragma unroll
        for (int u = 0; u < kSkinnyPre; ++u) {
            const int s = s0 + u * kSkinnyWarps;
            if (s < steps) {
                wa[u] = ldg_nc_v4(w_lo + (size_t)s * 64);
                wb[u] = ldg_nc_v4(w_hi + (size_t)s * 64);
            }
        }
    };
    int tile = blockIdx.x;
    if (tile < tiles) load_w(tile, warp);
    grid_dep_wait();
    if (XS) {
        const int cpr = p.K / 16;   // 16-byte chunks per row
        for (int i = threadIdx.x; i < NB * 8 * cpr; i += kSkinnyWarps * 32) {
            const int m = i / cpr, c = i - m * cpr;
            const uint4 v = m < p.M ? reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K)[c] : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(xs + (size_t)m * xs_ld + (size_t)c * 16) = v;
        }
        __syncthreads();
    }

    // Persistent over the row tiles: the first chunk of the NEXT tile's weights is requested before this tile's partial
    // sums go through shared memory, so the stream does not pause for the reduction and the epilogue.
    for (; tile < tiles; tile += gridDim.x) {
        const int n0 = tile * kSkinnyRows;
        int32_t c[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) c[nb][0] = c[nb][1] = c[nb][2] = c[nb][3] = 0;

        for (int s0 = warp; s0 < steps; s0 += kSkinnyPre * kSkinnyWarps) {
            if (s0 != warp) load_w(tile, s0);
#pragma unroll
            for (int u = 0; u < kSkinnyPre; ++u) {
                const int s = s0 + u * kSkinnyWarps;
                if (s >= steps) break;
                uint4 xa[NB], xb[NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const int m = nb * 8 + g;
                    if (XS) {
                        const uint4* xp = reinterpret_cast<const uint4*>(xs + (size_t)m * xs_ld + (size_t)s * 128 + t * 32);
                        xa[nb] = xp[0];
                        xb[nb] = xp[1];
                    } else if (m < p.M) {
                        const uint4* xp = reinterpret_cast<const uint4*>(p.x + (size_t)m * p.K + (size_t)s * 128 + t * 32);
                        xa[nb] = xp[0];
                        xb[nb] = xp[1];
                    } else {
                        xa[nb] = xb[nb] = make_uint4(0, 0, 0, 0);
                    }
                }
                const uint32_t pa[4] = {wa[u].x, wa[u].y, wa[u].z, wa[u].w}, pb[4] = {wb[u].x, wb[u].y, wb[u].z, wb[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t a0 = (pa[i] << 4) & 0xF0F0F0F0u, a2 = pa[i] & 0xF0F0F0F0u;   // codes 0,2,4,6 / 1,3,5,7 of the word (x16)
                    const uint32_t a1 = (pb[i] << 4) & 0xF0F0F0F0u, a3 = pb[i] & 0xF0F0F0F0u;
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        const uint32_t xw[8] = {xa[nb].x, xa[nb].y, xa[nb].z, xa[nb].w, xb[nb].x, xb[nb].y, xb[nb].z, xb[nb].w};
                        imma16832(c[nb], a0, a1, a2, a3, xw[2 * i], xw[2 * i + 1]);
                    }
                }
            }
        }
        if (tile + (int)gridDim.x < tiles) load_w(tile + gridDim.x, warp);
        // C fragment: c0 (row g, token 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1) of the 8-token block
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            red[warp][nb * 8 + 2 * t][g] = c[nb][0];
            red[warp][nb * 8 + 2 * t + 1][g] = c[nb][1];
            red[warp][nb * 8 + 2 * t][g + 8] = c[nb][2];
            red[warp][nb * 8 + 2 * t + 1][g + 8] = c[nb][3];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < NB * 8 * kSkinnyRows; e += kSkinnyWarps * 32) {   // uniform trip count per warp
            const int row = e & 15, m = e >> 4, n = n0 + row;
            int32_t s = 0;
#pragma unroll
            for (int w = 0; w < kSkinnyWarps; ++w) s += red[w][m][row];
            const bool ok = m < p.M && n < p.N;
            float y = 0.f;
            if (ok) {
                y = __fmul_rn(__fmul_rn((float)(s >> 4), p.s_x[m]), p.s_w[n]);   // exact shift: every term is 16 w x
                if (p.bias) {
                    y = __fadd_rn(y, p.bf16_out ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                                : __half2float(reinterpret_cast<const __half*>(p.bias)[n]));
                }
            }
            if (SILU) {   // (gate, up) = rows (2j, 2j + 1): neighbouring lanes
                const __half h = __float2half_rn(y);
                const uint32_t up = __shfl_xor_sync(0xffffffffu, (uint32_t)__half_as_ushort(h), 1);
                if (ok && !(row & 1)) {
                    const float gf = __half2float(h);
                    const __half act = __float2half_rn(gf / (1.f + expf(-gf)));
                    reinterpret_cast<__half*>(p.out)[(size_t)m * (p.N / 2) + (n >> 1)] = __hmul(act, __ushort_as_half((uint16_t)up));
                }
            } else if (ok) {
                if (p.bf16_out) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)m * p.N + n] = __float2bfloat16_rn(y);
                else reinterpret_cast<__half*>(p.out)[(size_t)m * p.N + n] = __float2half_rn(y);
            }
        }
        __syncthreads();   // `red` is rewritten by the next tile
    }
}

}  // namespace flatq
