// Online Kronecker transform + per-token INT4 quantisation, second tcgen05 design (round 2): the intermediate product
// never leaves TENSOR MEMORY and several token groups are in flight per CTA.
//
// Replaces the Triton programs of deploy/kernels/kron_matmul.py:24-130,192-266 (a9) and the quantiser of
// flatquant/quant_utils.py:77-119 (a4).  Evaluation order and rounding points are the deployed reference's
// (kron_matmul.py:62-76: (L^T X) first, fp16 intermediate, then . R):
//     Y = L^T . X   (fp32 accumulate, rounded to fp16),   Z = Y . R   (fp32 accumulate, rounded to fp16),
//     s = max(|min(Z,0)*c_min|, max(Z,0)*c_max)/7,  q = clamp(rint(Z/s), -8, 7).
//
// Why a second design: the first one (kron_quant_tc.cuh) was laid out around a "128 cycles per tcgen05.mma whatever N"
// floor that tools/mma_rate2.cu showed to be an artefact of the probe's issue loop -- an M128 x N64 x K16 MMA costs
// ~45 cycles, M128 x N128 64 cycles.  Small-N MMAs are cheap, so the products can be arranged for the shortest
// dependency chain instead of the fewest instructions, and the chain TMA -> MMA -> TMEM read -> SMEM write (transposed)
// -> fence.proxy.async -> MMA of the first design (one group in flight per CTA, 0.36 of the HBM roofline) goes away:
//
//   rows m = (t, i') of a group of T tokens live on the 128 TMEM lanes (token t at lanes [t*128/T, t*128/T + n1))
//   stage 1   D[(t,i'), j]  = sum_(t',i)  A1[(t,i'), (t',i)] . X[(t',i), j]
//             A1 = block-diagonal L^T in SMEM (K-major, built once per CTA), B = the X tile exactly as TMA delivers it
//             ([T*n1 rows][n2 contiguous] = an MN-major operand), M = 128, N = n2, K = T*n1
//   convert   each of 128 threads reads ITS lane's n2 fp32 values, rounds to fp16 and stores them back to TMEM as the
//             packed A operand of stage 2 (tcgen05.ld -> cvt -> tcgen05.st: no shared memory, no proxy fence)
//   stage 2   D[(t,i'), j'] = sum_j  Y[(t,i'), j] . R[j, j']       A = TMEM, B = R as it lies in HBM (MN-major, by TMA)
//   quantise  the thread of row (t, i') holds Z_t[i', 0..n2): per-token max/min over the token's warps, codes, and ONE
//             contiguous n2-byte (int8 GEMM operand) or n2/2-byte (packed nibbles) store per thread -- a warp writes
//             32 consecutive rows = 2-4 KB contiguous, no staging tile.
//
// Several groups are in flight per CTA (TMEM slot = n2 accumulator columns, reused by both stages, + n2/2 columns of
// packed Y; group `it` of a CTA uses slot it % kSlots).  Warps: 0 TMA producer, 1 issues the stage-1 MMAs (+ TMEM
// alloc), 2 issues the stage-2 MMAs -- two issuers, so that stage 1 runs ahead by as many groups as there are free
// slots instead of waiting for the conversion of the previous group --, 3 idle, 4.. worker warpgroups (group `it` is
// converted and quantised by warpgroup it % kWorkerGroups).  One persistent CTA per SM.
#pragma once
#include "kron_quant_tc.cuh"

namespace flatq {

template <int N1, int N2, int T>
struct KronTs {
    static_assert(N2 == 64 || N2 == 128, "n2 must be 64 or 128");
    static_assert(T == 1 || T == 2 || T == 4, "tokens per group");
    static_assert(T * N1 <= 128 && (T * N1) % 16 == 0 && N1 <= 128 / T, "group rows");
    static constexpr int kMS = 128 / T;                      // TMEM lanes per token
    static constexpr int kK1 = T * N1;                       // stage-1 contraction length (rows of the X tile)
    static constexpr int kK1Atoms = (kK1 + 63) / 64;         // 64-element K atoms of A1
    static constexpr int kPanels = N2 / 64;                  // 128-byte panels per X / R row
    static constexpr int kXPanelBytes = kK1 * 128;
    static constexpr int kXStageBytes = kPanels * kXPanelBytes;
    static constexpr int kA1Bytes = kK1Atoms * 128 * 128;    // [atoms][128 rows][128 B]
    static constexpr int kRBytes = kPanels * N2 * 128;       // [panels][n2 rows][128 B]
    static constexpr int kSlotCols = N2 + N2 / 2;
    static constexpr int kSlots = N2 == 64 ? 5 : 2;          // TMEM slots: groups between "stage 1 issued" and "accumulator read out"
    static constexpr int kWorkerGroups = N2 == 64 ? 3 : 2;   // warpgroups doing convert / quantise (group g: pairs it % kWorkerGroups == g)
    static constexpr int kTmemCols = 512;
    static constexpr int kXStages = N2 == 64 ? 4 : 3;
    static constexpr int kThreads = (4 + 4 * kWorkerGroups) * 32;
    static constexpr int kSmemBytes = kXStages * kXStageBytes + kA1Bytes + kRBytes + 1024 /*align*/ + 1024 /*barriers, reductions*/;
};

template <int N1, int N2, int T, bool OUT8>
__global__ void __launch_bounds__(KronTs<N1, N2, T>::kThreads, 1)
kron_quant_ts_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_r, const KronParams p) {
    using C = KronTs<N1, N2, T>;
    constexpr int NS = C::kSlots, XS = C::kXStages, NWG = C::kWorkerGroups;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xs = smem;                                   // [XS][panels][kK1 rows][128 B]
    uint8_t* a1s = xs + XS * C::kXStageBytes;             // block-diagonal L^T, K-major SW128
    uint8_t* rs = a1s + C::kA1Bytes;                      // R, MN-major panels (TMA)
    uint8_t* tail = rs + C::kRBytes;
    uint64_t* full_x = reinterpret_cast<uint64_t*>(tail); // [XS]
    uint64_t* empty_x = full_x + XS;                      // [XS]
    uint64_t* d1_full = empty_x + XS;                     // [NS] stage-1 accumulator complete
    uint64_t* a2_ready = d1_full + NS;                    // [NS] packed Y written (4 worker warps)
    uint64_t* d2_full = a2_ready + NS;                    // [NS] stage-2 accumulator complete
    uint64_t* d_free = d2_full + NS;                      // [NS] accumulator read out (4 worker warps)
    uint64_t* bar_r = d_free + NS;                        // R panels landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_r + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 2); // [NWG][4 warps][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = (p.tokens + T - 1) / T;
    // PDL trigger first: at decode sizes this kernel is one CTA, and the GEMV that follows uses the time to become
    // resident and pull its (static) weight rows into registers (gemv_w4a8.cuh); a prefill-sized dependent cannot become
    // resident before this grid's CTAs retire (shared memory), so the early trigger costs nothing there.
    grid_dep_launch();

    if (threadIdx.x == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(&full_x[i], 1); mbar_init(&empty_x[i], 1); }
        for (int i = 0; i < NS; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&a2_ready[i], 4); mbar_init(&d2_full[i], 1); mbar_init(&d_free[i], 4); }
        mbar_init(bar_r, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x);
        mbar_arrive_expect_tx(bar_r, C::kRBytes);   // the factors are static weights: no dependency on the previous kernel
#pragma unroll
        for (int a = 0; a < C::kPanels; ++a) tma_load_2d(rs + a * (N2 * 128), &tm_r, bar_r, a * 128 /*bytes*/, 0);
    }
    if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
    // block-diagonal L^T: A1[(t, i')][(t', i)] = (t == t') L[i][i'], zero elsewhere (padding rows and columns included).
    // Zero fill with 16-byte stores, then 16-byte coalesced reads of L's rows scattered as 2-byte stores (transposition).
    for (int e = threadIdx.x; e < C::kA1Bytes / 16; e += C::kThreads) st_shared_v4(smem_u32(a1s) + e * 16, make_uint4(0, 0, 0, 0));
    __syncthreads();
    for (int e = threadIdx.x; e < N1 * (N1 / 8); e += C::kThreads) {
        const int i = e / (N1 / 8), ip0 = (e % (N1 / 8)) * 8;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.left + i * N1 + ip0));
        const uint16_t* h = reinterpret_cast<const uint16_t*>(&v);
#pragma unroll
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                *reinterpret_cast<uint16_t*>(a1s + sw128_off(128, t * C::kMS + ip0 + u, t * N1 + i)) = h[u];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    grid_dep_wait();

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
                const uint32_t s = it % XS, ph = (it / XS) & 1;
                mbar_wait(&empty_x[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_x[s], C::kXStageBytes);
#pragma unroll
                for (int a = 0; a < C::kPanels; ++a)
                    tma_load_2d(xs + s * C::kXStageBytes + a * C::kXPanelBytes, &tm_x, &full_x[s], a * 128 /*bytes*/, g * C::kK1);
            }
        }
    } else if (warp == 1) {
        // ================================ stage-1 MMA issuer ================================
        constexpr uint32_t idesc1 = make_idesc(kDF32, kFmtF16, kFmtF16, 0, 1, 128, N2);   // A1 K-major (SMEM), B = X MN-major
        const uint32_t a1addr = smem_u32(a1s);
        uint32_t it = 0;
        for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
            const uint32_t s = it % NS, ph = (it / NS) & 1;
            const uint32_t xi = it % XS, xph = (it / XS) & 1;
            mbar_wait(&full_x[xi], xph);
            mbar_wait(&d_free[s], ph ^ 1);       // the slot's previous accumulator has been read out
            tc_fence_after();
            if (lane == 0) {
                const uint32_t xaddr = smem_u32(xs + xi * C::kXStageBytes);
                const uint32_t d = tmem + s * C::kSlotCols;
#pragma unroll
                for (int kk = 0; kk < C::kK1 / 16; ++kk) {
                    const uint64_t adesc = make_smem_desc_kmajor(a1addr + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 1024, kLayoutSw128);
                    const uint64_t bdesc = make_smem_desc_mnmajor(xaddr + kk * 2048, C::kXPanelBytes, 1024, kLayoutSw128);
                    mma_f16_ss(d, adesc, bdesc, idesc1, kk != 0);
                }
                tc_commit(&empty_x[xi]);
                tc_commit(&d1_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ================================ stage-2 MMA issuer ================================
        constexpr uint32_t idesc2 = make_idesc(kDF32, kFmtF16, kFmtF16, 0, 1, 128, N2);   // A = TMEM, B = R MN-major
        mbar_wait(bar_r, 0);
        const uint32_t raddr = smem_u32(rs);
        uint32_t it = 0;
        for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
            const uint32_t s = it % NS, ph = (it / NS) & 1;
            mbar_wait(&a2_ready[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d = tmem + s * C::kSlotCols, a2 = d + N2;
#pragma unroll
                for (int kk = 0; kk < N2 / 16; ++kk) {
                    const uint64_t bdesc = make_smem_desc_mnmajor(raddr + kk * 2048, N2 * 128, 1024, kLayoutSw128);
                    mma_f16_ts(d, a2 + kk * 8, bdesc, idesc2, kk != 0);
                }
                tc_commit(&d2_full[s]);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ================================ workers: convert, then quantise + store ================================
        const int wg = (warp - 4) >> 2;            // worker warpgroup: groups it % NWG == wg
        const int q = warp & 3;                    // TMEM lane quadrant
        const int row = q * 32 + lane;             // (t, i')
        const int tok_in = row / C::kMS, ip = row % C::kMS;
        const bool row_ok = ip < N1;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float* red_w = red + wg * 8;
        uint32_t it = wg;
        for (int g = blockIdx.x + wg * gridDim.x; g < groups; g += NWG * gridDim.x, it += NWG) {
            const uint32_t s = it % NS, ph = (it / NS) & 1;
            const uint32_t d = lane_addr + s * C::kSlotCols, a2 = d + N2;
            // ---- stage-1 accumulator -> fp16 -> packed A operand of stage 2 (same lanes); all loads in flight at once
            mbar_wait(&d1_full[s], ph);
            tc_fence_after();
            {
                uint32_t r[N2];
#pragma unroll
                for (int c = 0; c < N2; c += 32) tmem_ld_x32(d + c, *reinterpret_cast<uint32_t(*)[32]>(&r[c]));
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < N2; c += 32) {
                    uint32_t h[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) h[j] = pack_h2(__uint_as_float(r[c + 2 * j]), __uint_as_float(r[c + 2 * j + 1]));
                    tmem_st_x16(a2 + c / 2, h);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a2_ready[s]);

            // ---- stage-2 accumulator -> Z (fp16) -> per-token scale -> codes
            mbar_wait(&d2_full[s], ph);
            tc_fence_after();
            uint32_t z[N2 / 2];
            __half2 hmax = __float2half2_rn(0.f), hmin = __float2half2_rn(0.f);
            {
                uint32_t r[N2];
#pragma unroll
                for (int c = 0; c < N2; c += 32) tmem_ld_x32(d + c, *reinterpret_cast<uint32_t(*)[32]>(&r[c]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_free[s]);
#pragma unroll
                for (int j = 0; j < N2 / 2; ++j) {
                    z[j] = pack_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    const __half2 zz = *reinterpret_cast<const __half2*>(&z[j]);
                    hmax = __hmax2(hmax, zz);
                    hmin = __hmin2(hmin, zz);
                }
            }
            float vmax = fmaxf(__low2float(hmax), __high2float(hmax));
            float vmin = fminf(__low2float(hmin), __high2float(hmin));
            if (!row_ok) { vmax = 0.f; vmin = 0.f; }   // padding lanes hold exact zeros anyway
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            }
            if (T < 4) {   // a token spans 4 (T = 1) or 2 (T = 2) warps of this warpgroup
                named_bar_sync(1 + wg, 128);           // the previous group's reads of red_w are done
                if (lane == 0) { red_w[q * 2] = vmax; red_w[q * 2 + 1] = vmin; }
                named_bar_sync(1 + wg, 128);
                if (T == 1) {
                    vmax = fmaxf(fmaxf(red_w[0], red_w[2]), fmaxf(red_w[4], red_w[6]));
                    vmin = fminf(fminf(red_w[1], red_w[3]), fminf(red_w[5], red_w[7]));
                } else {
                    const int w0 = (q >> 1) * 2;
                    vmax = fmaxf(red_w[w0 * 2], red_w[w0 * 2 + 2]);
                    vmin = fminf(red_w[w0 * 2 + 1], red_w[w0 * 2 + 3]);
                }
            }
            const float amax = fmaxf(fabsf(vmin * p.c_min), vmax * p.c_max);
            const float scale = amax == 0.f ? 1.f : __fdiv_rn(amax, 7.0f);
            const float inv_s = __frcp_rn(scale);
            const int token = g * T + tok_in;
            const bool tok_ok = row_ok && token < p.tokens;
            if (tok_ok && ip == 0) {
                if (p.sx_f32) p.sx_f32[token] = scale;
                if (p.sx_f16) p.sx_f16[token] = __float2half_rn(scale);
            }
            const uint64_t inv2 = pack_f32x2(inv_s, inv_s), nscale2 = pack_f32x2(-scale, -scale);
            // clamp BEFORE scaling, two values per instruction in fp16: any bound inside (7 s, 7.5 s) / (-8.5 s, -8 s) gives
            // the same codes as clamping the rounded quotient to [-8, 7] (a bound that overflows fp16 clamps nothing,
            // and then nothing needs clamping: |z| <= 65504 < 8.25 s)
            const __half2 hi2 = __float2half2_rn(7.25f * scale), lo2 = __float2half2_rn(-8.25f * scale);
            auto codes2 = [&](uint32_t zw, uint32_t& even, uint32_t& odd) {   // low byte of each = the int8 code
                const __half2 zc = __hmax2(__hmin2(*reinterpret_cast<const __half2*>(&zw), hi2), lo2);
                const float2 f = __half22float2(zc);
                const uint64_t z2 = pack_f32x2(f.x, f.y);
                uint64_t q0, r, q1, m;
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q0) : "l"(z2), "l"(inv2));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(nscale2), "l"(q0), "l"(z2));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q1) : "l"(r), "l"(inv2), "l"(q0));   // correctly rounded z / s
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(q1), "l"(pack_f32x2(12582912.f, 12582912.f)));
                asm("mov.b64 {%0, %1}, %2;" : "=r"(even), "=r"(odd) : "l"(m));
            };
            if (OUT8) {
                // int8 GEMM operand: inside every 8 consecutive k the order is [0,2,4,6,1,3,5,7] (gemm_w4a8.cuh)
                uint4* dst = reinterpret_cast<uint4*>(p.xq8 + (size_t)token * (N1 * N2) + ip * N2);
#pragma unroll
                for (int c = 0; c < N2 / 16; ++c) {     // 16 codes = 16 bytes per store
                    uint32_t w[4];
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        uint32_t ev[4], od[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) codes2(z[c * 8 + h8 * 4 + j], ev[j], od[j]);
                        w[h8 * 2] = __byte_perm(__byte_perm(ev[0], ev[1], 0x0040), __byte_perm(ev[2], ev[3], 0x0040), 0x5410);
                        w[h8 * 2 + 1] = __byte_perm(__byte_perm(od[0], od[1], 0x0040), __byte_perm(od[2], od[3], 0x0040), 0x5410);
                    }
                    if (tok_ok) dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                // packed nibbles, element 2j low / 2j+1 high (deploy/functional/quantization.py:54-55)
                uint4* dst = reinterpret_cast<uint4*>(p.xq4 + (size_t)token * (N1 * N2 / 2) + ip * (N2 / 2));
#pragma unroll
                for (int c = 0; c < N2 / 32; ++c) {     // 32 codes = 16 bytes per store
                    uint32_t w[4];
#pragma unroll
                    for (int wi = 0; wi < 4; ++wi) {
                        uint32_t b[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t ev, od;
                            codes2(z[c * 16 + wi * 4 + j], ev, od);
                            b[j] = (ev & 0xF) | (od << 4);      // byte j in the low 8 bits (upper bits dropped by the permute)
                        }
                        w[wi] = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
                    }
                    if (tok_ok) dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::kTmemCols>(tmem);
    }
}

}  // namespace flatq
