// W4 x A8 GEMM on CTA PAIRS (tcgen05.mma.cta_group::2) for prefill shapes (M > 128).
//
// Same arithmetic, operand formats and epilogue as gemm_w4a8.cuh (which replaces deploy/kernels/gemm.cu:7-46 and
// deploy/kernels/quant.cu:66-101); what changes is who feeds the tensor cores.  On B200 the GEMM runs against the
// board power limit (profiles/r01_streamk_power_bound.md), so the bytes moved per MAC matter as much as idle cycles.
// The single-CTA tile is 128 weight rows x 256 tokens: per 256-deep K stage an SM pulls 16 KB of packed weights and
// 64 KB of int8 activations out of L2 and writes them to shared memory.  Here two SMs of a TPC share one
// 256-weight-row x 256-token tile:
//
//   * each CTA expands ITS 128 weight rows into its own TMEM (A operand) exactly as before;
//   * each CTA loads only HALF of the activation tile (128 tokens x 256 B = 32 KB per stage); the MMA unit of either
//     SM reads the other half from the peer's shared memory (B operand of a cta_group::2 MMA: N/2 rows per CTA);
//   * one thread of the even CTA issues UMMA M=256 (128 lanes in each CTA) x N=256 x K=32 for the pair.
//
//   => 48 KB instead of 80 KB from L2 and into SMEM per SM and stage, 128 KB of SMEM free for a 4-deep weight ring.
//
// Synchronisation (s = stage):
//   full[s]   (leader's)  <- 64 KB of X bytes from both CTAs' TMA (complete_tx on the leader) + 8 expander warps
//                            (4 local arrivals, 4 from the peer through mapa / shared::cluster)
//   empty[s]  (both)      <- tcgen05.commit.cta_group::2 multicast: slot s of BOTH CTAs may be refilled
//   wfull/wempty          <- per CTA, as in the single-CTA kernel
//   full_acc  (both)      <- commit multicast;   empty_acc (leader's) <- 16 epilogue warps of each CTA
#pragma once
#include "gemm_w4a8.cuh"

namespace flatq {

constexpr int k2smStages = 4;                    // X smem ring == TMEM A ring (64 columns per stage)
constexpr int k2smStagesW = 4;                   // packed-W smem ring
constexpr int k2smTokens = 256;                  // UMMA N: tokens per tile
constexpr int k2smXHalf = 128 * 128;             // one TMA box: this CTA's 128 tokens x 128 B
constexpr int k2smXStage = 2 * k2smXHalf;        // 256 K-bytes per token row, as two swizzle atoms: 32 KB
constexpr int k2smBytes = k2smStages * k2smXStage + k2smStagesW * kWStageBytes + 1024 /*align*/ + 512 /*bars*/ + 1024 /*scales*/;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_w4a8_2sm_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                     const GemmParams p) {
    constexpr int TN = k2smTokens;
    extern __shared__ uint8_t smem_raw[];
    // the pair addresses each other's shared memory by offset: both CTAs must lay it out identically
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xs = smem;
    uint8_t* ws = xs + k2smStages * k2smXStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ws + k2smStagesW * kWStageBytes);
    uint64_t* full = bars;                       // [k2smStages]   (used in the leader)
    uint64_t* empty = full + k2smStages;         // [k2smStages]
    uint64_t* wfull = empty + k2smStages;        // [k2smStagesW]
    uint64_t* wempty = wfull + k2smStagesW;      // [k2smStagesW]
    uint64_t* full_acc = wempty + k2smStagesW;   // [1]
    uint64_t* empty_acc = full_acc + 1;          // [1]            (used in the leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_acc + 1);
    float* sx_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [256]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();     // 0: leader (issues the MMAs), 1: peer

    if (threadIdx.x == 0) {
        for (int i = 0; i < k2smStages; ++i) { mbar_init(&full[i], 1 + 8); mbar_init(&empty[i], 1); }
        for (int i = 0; i < k2smStagesW; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 4); }
        mbar_init(full_acc, 1);
        mbar_init(empty_acc, 32);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_x);
    if (warp == 2 && lane == 0) tma_prefetch_desc(&tm_w);
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before anyone arrives remotely; TMEM allocated
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    grid_dep_launch();
    // Programmatic dependent launch: only what the previous kernel produced (activation codes, scales, the output
    // buffer) needs the dependency.  The packed weights are static, so the W producer starts streaming them at once
    // and the first stages are already in shared memory when the transform kernel's tail retires.
    if (warp != 2) grid_dep_wait();

    const int nb_w = (p.N + kGemmBlockW - 1) / kGemmBlockW;
    const int nb_w2 = (nb_w + 1) / 2;            // pairs of weight blocks (an odd last block pairs with TMA zero fill)
    const int nb_m = (p.M + TN - 1) / TN;
    const int KB = (p.K + kGemmBlockK - 1) / kGemmBlockK;
    const int n_tiles = nb_w2 * nb_m;
    const int cid = (int)(blockIdx.x >> 1), ncl = (int)(gridDim.x >> 1);
    // tile order: token blocks fastest (see gemm_w4a8.cuh)

    if (warp < 4) {
      reg_dealloc<kRegsCtl>();
      if (warp == 0) {
        // ================================ X producer (this CTA's 128 tokens) ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = cid; tile < n_tiles; tile += ncl) {
                const int m0 = (tile % nb_m) * TN + (int)rank * 128;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % k2smStages, ph = (it / k2smStages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * k2smXStage);  // both CTAs' halves
                    tma_load_2d_pair(xs + s * k2smXStage, &tm_x, &full[s], kb * kGemmBlockK, m0);
                    tma_load_2d_pair(xs + s * k2smXStage + k2smXHalf, &tm_x, &full[s], kb * kGemmBlockK + 128, m0);
                }
            }
        }
      } else if (warp == 2) {
        // ================================ W producer (this CTA's 128 weight rows) ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = cid; tile < n_tiles; tile += ncl) {
                const int n0 = ((tile / nb_m) * 2 + (int)rank) * kGemmBlockW;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % k2smStagesW, ph = (it / k2smStagesW) & 1;
                    mbar_wait(&wempty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&wfull[s], kWStageBytes);
                    tma_load_2d(ws + s * kWStageBytes, &tm_w, &wfull[s], kb * (kGemmBlockK / 2), n0);
                }
            }
        }
      } else if (warp == 1 && rank == 0) {
        // ================================ MMA issuer (leader CTA only) ================================
        constexpr uint32_t idesc = make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, 2 * kGemmBlockW, TN);
        uint32_t it = 0, tile_iter = 0;
        for (int tile = cid; tile < n_tiles; tile += ncl, ++tile_iter) {
            mbar_wait(empty_acc, (tile_iter & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % k2smStages, ph = (it / k2smStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t xaddr = smem_u32(xs + s * k2smXStage);
                    const uint32_t a_tmem = tmem + kTmemColsA0 + s * 64;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t bdesc =
                            make_smem_desc_kmajor(xaddr + (kk >> 2) * k2smXHalf + (kk & 3) * 32, 1024, kLayoutSw128);
                        mma_i8_ts_pair(tmem, a_tmem + kk * 8, bdesc, idesc, (kb | kk) != 0);
                    }
                    tc_commit_pair(&empty[s], 0b11);
                    if (kb == KB - 1) tc_commit_pair(full_acc, 0b11);
                }
                __syncwarp();
            }
        }
      }
    } else if (warp < 12) {
        // ================================ nibble expanders ================================
        if constexpr (kRegsExpand < 72) reg_dealloc<kRegsExpand>();
        const int grp = (warp - 4) >> 2;  // handles stage iterations with it % 2 == grp
        const int q = warp & 3;           // TMEM lane quadrant this warp may touch
        const uint32_t row = q * 32 + lane;
        const uint32_t ws_addr = smem_u32(ws) + row * 128 + ((row & 7) << 4);  // see gemm_w4a8.cuh (XOR swizzle)
        uint32_t it = 0;
        for (int tile = cid; tile < n_tiles; tile += ncl) {
            for (int kb = 0; kb < KB; ++kb, ++it) {
                if ((it & 1) != (uint32_t)grp) continue;
                const uint32_t sw = it % k2smStagesW, phw = (it / k2smStagesW) & 1;
                const uint32_t s = it % k2smStages, ph = (it / k2smStages) & 1;
                mbar_wait(&wfull[sw], phw);
                uint4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = ld_shared_v4((ws_addr + sw * kWStageBytes) ^ (c << 4));
                mbar_wait(&empty[s], ph ^ 1);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + kTmemColsA0 + s * 64;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    uint32_t a[16];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint4 u = v[2 * h + c];
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            a[c * 8 + 2 * j] = (w4[j] << 4) & 0xF0F0F0F0u;
                            a[c * 8 + 2 * j + 1] = w4[j] & 0xF0F0F0F0u;
                        }
                    }
                    tmem_st_x16(taddr + h * 16, a);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&wempty[sw]);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&full[s], 0);   // the leader's barrier, from either CTA
            }
        }
    } else {
        // ================================ epilogue (this CTA's 128 weight rows x 256 tokens) ================================
        reg_alloc<kRegsEpi>();
        const int ew = warp - 12;
        const int q = warp & 3;
        const int half = ew >> 2;
        constexpr int COLS = TN / 4;
        const int c0 = half * COLS;
        const bool fold = p.acc_shift == 4 && p.K <= 16384;
        uint32_t tile_iter = 0;
        for (int tile = cid; tile < n_tiles; tile += ncl, ++tile_iter) {
            const int n0 = ((tile / nb_m) * 2 + (int)rank) * kGemmBlockW;
            const int m0 = (tile % nb_m) * TN;
            const int n = n0 + q * 32 + lane;
            const bool n_ok = n < p.N;
            float sw_n = 0.f, bias_n = 0.f;
            if (MODE != kEpiI32 && n_ok) {
                sw_n = p.s_w[n];
                if (p.bias) {
                    bias_n = MODE == kEpiF16 ? __half2float(reinterpret_cast<const __half*>(p.bias)[n])
                                             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n]);
                }
            }
            if (MODE != kEpiI32) {
                named_bar_sync(1 + half, 128);
                if (q == 0) {
#pragma unroll
                    for (int i = lane; i < COLS; i += 32) {
                        const int m = m0 + c0 + i;
                        const float v = m < p.M ? p.s_x[m] : 0.f;
                        sx_s[c0 + i] = fold ? v * 0.0625f : v;
                    }
                }
                named_bar_sync(1 + half, 128);
            }
            mbar_wait(full_acc, tile_iter & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem + ((uint32_t)(q * 32) << 16) + c0;
            uint32_t r[COLS];
#pragma unroll
            for (int c = 0; c < COLS / 32; ++c) tmem_ld_x32(t_acc + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(empty_acc, 0);   // accumulator of this CTA drained: tell the leader
            if (MODE == kEpiI32) {
                int32_t* dst = reinterpret_cast<int32_t*>(p.out) + (size_t)(m0 + c0) * p.N + n;
#pragma unroll
                for (int j = 0; j < COLS; ++j) {
                    if (m0 + c0 + j < p.M && n_ok) dst[(size_t)j * p.N] = static_cast<int32_t>(r[j]) >> p.acc_shift;
                }
            } else {
                const bool has_bias = p.bias != nullptr;
                uint16_t* out16 = reinterpret_cast<uint16_t*>(p.out);
                uint32_t idx = (uint32_t)(m0 + c0) * (uint32_t)p.N + (uint32_t)n;
                if (p.silu_pairs) idx = (uint32_t)(m0 + c0) * (uint32_t)(p.N >> 1) + (uint32_t)(n >> 1);
                if (p.rows_per_owner > 0) {   // reduce-scatter in the epilogue: this tile goes to its owner's staging slot
                    const int owner = m0 / p.rows_per_owner;
                    out16 = reinterpret_cast<uint16_t*>(p.peer_out[owner]);
                    idx = (uint32_t)(p.tp_rank * p.rows_per_owner + (m0 - owner * p.rows_per_owner) + c0) * (uint32_t)p.N + (uint32_t)n;
                }
                const int j_end = n_ok ? p.M - (m0 + c0) : 0;
                auto finish = [&](auto fold_c) {
                    constexpr bool FOLD = decltype(fold_c)::value;
#pragma unroll
                    for (int j = 0; j < COLS; j += 4) {
                        const float4 s4 = *reinterpret_cast<const float4*>(&sx_s[c0 + j]);
                        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                        if (MODE == kEpiF16 && p.silu_pairs) {
                            // Lanes (2k, 2k + 1) hold (gate_k, up_k) of every token.  Tokens go two at a time: the even lane
                            // finishes token j + i (it receives that token's `up`), the odd lane token j + i + 1 (it receives
                            // that token's `gate`) -- ONE exchange and ONE SiLU per lane and token pair, all lanes busy.
                            const bool odd = (lane & 1) != 0;
#pragma unroll
                            for (int i = 0; i < 4; i += 2) {
                                uint16_t hh[2];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int32_t a = FOLD ? static_cast<int32_t>(r[j + i + e]) : static_cast<int32_t>(r[j + i + e]) >> p.acc_shift;
                                    hh[e] = __half_as_ushort(__float2half_rn(__fmul_rn(__fmul_rn((float)a, sv[i + e]), sw_n)));
                                }
                                const uint32_t recv = __shfl_xor_sync(0xffffffffu, (uint32_t)(odd ? hh[0] : hh[1]), 1);
                                const uint16_t g16 = odd ? (uint16_t)recv : hh[0], u16 = odd ? hh[1] : (uint16_t)recv;
                                const float gf = __half2float(__ushort_as_half(g16));
                                const __half s = __float2half_rn(gf / (1.f + expf(-gf)));            // F.silu: fp32 math, fp16 result
                                const uint16_t o16 = __half_as_ushort(__hmul(s, __ushort_as_half(u16)));   // * up, rounded to fp16
                                if (j + i + (odd ? 1 : 0) < j_end) out16[idx + (odd ? (uint32_t)(p.N >> 1) : 0u)] = o16;
                                idx += (uint32_t)p.N;   // two tokens of N / 2 columns
                            }
                            continue;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int32_t a = FOLD ? static_cast<int32_t>(r[j + i]) : static_cast<int32_t>(r[j + i]) >> p.acc_shift;
                            // explicit _rn ops: no FMA contraction, so the result equals the oracle's fp32 arithmetic
                            float y = __fmul_rn(__fmul_rn((float)a, sv[i]), sw_n);
                            if (has_bias) y = __fadd_rn(y, bias_n);
                            uint16_t h;
                            if (MODE == kEpiF16) h = __half_as_ushort(__float2half_rn(y));
                            else h = __bfloat16_as_ushort(__float2bfloat16_rn(y));
                            if (j + i < j_end) out16[idx] = h;
                            idx += (uint32_t)p.N;
                        }
                    }
                };
                if (fold) finish(FoldTrue{}); else finish(FoldFalse{});
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    cluster_sync_all();   // the peer may still read this CTA's SMEM / arrive on its barriers until both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<512>(tmem);
    }
}

}  // namespace flatq
