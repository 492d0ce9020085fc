// W4 x A8 -> INT32 GEMM on tcgen05 (UTCIMMA, kind::i8) for sm_100a.
//
// Replaces deploy/kernels/gemm.cu:7-46 (CUTLASS Sm80 s4 mma.sync GEMM) and, with the dequant
// epilogue, deploy/kernels/quant.cu:66-101 + the bias add of deploy/nn/linear.py:48-54.
//
// B200 has no INT4 tensor-core instruction (tcgen05.mma.kind::i4 does not exist), so the 4-bit
// weight stream is kept packed in HBM/L2/SMEM and expanded to int8 *in registers* on its way
// into TENSOR MEMORY, where it is consumed as the A operand of tcgen05.mma (A may come from
// TMEM, B must come from SMEM).  Roles are therefore swapped w.r.t. the reference GEMM:
//
//      D[n, m] (TMEM, int32)  =  sum_k  W[n, k] (TMEM, int8 = 16*w)  *  X[m, k] (SMEM, int8)
//
//   * UMMA M = 128 weight rows (TMEM lanes), UMMA N = TN tokens (16..128), UMMA K = 32.
//   * Nibble expansion is two logic ops per 8 weights: (p<<4)&0xF0F0F0F0 and p&0xF0F0F0F0 give
//     the even / odd nibbles as int8 values 16*w.  This permutes K inside every group of 8
//     ([0,2,4,6,1,3,5,7]); the activation operand is stored with the same permutation, so the
//     contraction is unchanged.  The factor 16 (or 256 when X is expanded the same way) is
//     removed exactly by an arithmetic shift of the int32 accumulator.
//   * Measured on B200 (tools/mma_rate2.cu, profiles/r02_mma_rate2.log; round 1's "128 cycles whatever N" was an
//     artefact of its probe): one tcgen05.mma (i8, M128 K32) costs N/2 cycles down to N = 128 (64 cycles) and hits
//     an issue floor of ~45 cycles below that, i.e. N >= 128 runs at the 8192 MAC/cycle/SM peak; a tcgen05.commit
//     after every 4 MMAs costs 144-173 cycles/MMA while >= 8 MMAs per commit stay at the flat rate.  Prefill tiles are therefore 128 weight rows x 256 tokens
//     (256 TMEM columns of int32 accumulator, single-buffered, + a 3-stage A ring of 64 columns) and
//     one pipeline stage is 256 K-elements = 8 MMAs behind ONE mbarrier wait and ONE commit.
//   * Stage contents: packed W tile 128 rows x 128 B (SWIZZLE_128B), int8 X tile = two TMA boxes of
//     TN rows x 128 B (SWIZZLE_128B).
//
// Warp roles (28 warps = 7 warpgroups, persistent over output tiles; registers re-partitioned with setmaxnreg):
//   warp 0  X TMA producer     warp 1  MMA issuer (+ TMEM alloc)     warp 2  W TMA producer     warp 3 idle   (40 regs)
//   warps 4-11  nibble expanders (2 groups x 4 lane-quadrants; a group takes every other stage)              (72 regs)
//   warps 12-27 epilogue (4 per lane-quadrant, each a quarter of the token columns)                           (80 regs):
//               copy the RAW accumulator columns into registers (one tcgen05.ld batch), hand the TMEM buffer
//               back to the MMA issuer at once, and only then dequantise and store -- the accumulator is
//               single-buffered at TN = 256, so everything before the release is exposed tensor-pipe idle time.
//
// Barriers per stage s:  full[s]  <- X bytes (expect_tx) + 4 expander warps;   empty[s] <- tcgen05.commit
//                        wfull[s] <- W bytes;                                  wempty[s] <- 4 expander warps
#pragma once
#include "grouped_sched.h"
#include "ptx.cuh"

namespace flatq {

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

struct FoldTrue { static constexpr bool value = true; };
struct FoldFalse { static constexpr bool value = false; };

#define FLATQ_TWAIT(slot, stmt)                          \
    do {                                                \
        if (PROF) {                                     \
            const long long t0_ = clock64();            \
            stmt;                                       \
            tacc[slot] += clock64() - t0_;              \
        } else {                                        \
            stmt;                                       \
        }                                               \
    } while (0)

constexpr int kGemmBlockW = 128;   // weight rows per tile
constexpr int kGemmBlockK = 256;   // K elements per stage
constexpr int kGemmWarps = 28;
#ifndef FLATQ_REGS_EPI
#define FLATQ_REGS_EPI 80
#endif
#ifndef FLATQ_REGS_CTL
#define FLATQ_REGS_CTL 40
#endif
#ifndef FLATQ_REGS_EXP
#define FLATQ_REGS_EXP 72
#endif
constexpr int kRegsCtl = FLATQ_REGS_CTL, kRegsExpand = FLATQ_REGS_EXP, kRegsEpi = FLATQ_REGS_EPI;  
constexpr int kGemmThreads = kGemmWarps * 32;
constexpr int kStagesW = 2;        // packed-W smem ring
constexpr int kTmemColsA0 = 256;   // accumulator(s) in columns [0,256): one 256-wide or two 128-wide buffers
constexpr int kWStageBytes = kGemmBlockW * (kGemmBlockK / 2);  // 16 KB

enum EpilogueMode : int { kEpiI32 = 0, kEpiF16 = 1, kEpiBF16 = 2 };

struct GemmParams {
    int M, N, K;          // tokens, out features, in features
    int acc_shift;        // 4: X holds true int8 codes; 8: X holds 16*code
    void* out;            // [M, N] int32 / f16 / bf16, row-major
    const float* s_x;     // [M]  per-token scale   (dequant modes)
    const float* s_w;     // [N]  per-channel scale (dequant modes)
    const void* bias;     // [N] same dtype as out, or nullptr
    int k_splits;         // > 1: split-K over CTAs, int32 atomics into `out` (MODE == kEpiI32 only; decode shapes)
    // stream-K for the last, partial wave of tiles (16-bit output modes): its tiles' K ranges are dealt evenly to all
    // CTAs; a CTA that continues a tile begun by its predecessor publishes a raw int32 partial, the CTA that began the
    // tile adds the partials of its successors in its epilogue.  Integer sums: the result is bit-identical.
    int streamk;                       // 0 / 1
    int32_t* sk_part;                  // [grid][TN columns][128 rows] int32 partial of CTA c's continued tile
    unsigned long long* sk_flags;      // [grid][16 epilogue warps]: == epoch once that warp's slice is published
    unsigned long long epoch;          // unique per launch (no zeroing of the workspace between launches)
    long long* prof;      // debug: per-CTA role timers (PROF instantiation only)
    int dbg;              // debug (PROF only): 1 skip TMEM stores, 2 skip MMA issue, 4 skip epilogue, 8 no X TMA, 16 no W TMA/LDS, 32 weight-blocks-fastest raster
    // grouped (mixture-of-experts) launch, GROUPED instantiations only (csrc/grouped_sched.h): n_groups GEMMs sharing
    // N and K; M = total rows of the sorted activation, s_w / bias are [n_groups, N], the W tensor map spans the
    // stacked [n_groups * N, K / 2] weight.  Appended so that the plain instantiations' parameter offsets are unchanged.
    const int* grp_off;   // [n_groups + 1] first row of every group (device; written by the router)
    const int* grp_tile;  // [n_groups + 1] tile prefix (device; grouped_prefix_kernel, same stream, just before)
    int n_groups;
    int t_upper;          // host-side bound on the number of tiles: nb_w * (ceil(M / TN) + n_groups)
    // row-parallel (tensor-parallel) launch of the CTA-pair kernel: the epilogue IS the reduce-scatter.  Token rows
    // [o rows_per_owner, (o + 1) rows_per_owner) belong to rank o; every tile is stored through the mapped peer pointer
    // into slot `tp_rank` of its owner's staging buffer [tp slots][rows_per_owner][N] (csrc/peer_reduce.cuh).
    // rows_per_owner == 0: plain launch (out).  rows_per_owner % 256 == 0: a 256-token tile never straddles two owners.
    int rows_per_owner, tp_rank;
    void* peer_out[8];
    // SiLU(gate) * up in the epilogue of the fused [gate; up] GEMM (CTA-pair kernel, 16-bit outputs): the weight rows are
    // interleaved offline (row 2j = gate_j, row 2j + 1 = up_j), so a lane pair holds (gate_j, up_j) of a token; the even
    // lane rounds both to the output type (what the unfused GEMM would have stored), applies HF's act_fn(gate) * up with
    // the rounding points of silu_mul_f16_kernel (glue.cuh) and stores column j of an [M, N / 2] output.
    int silu_pairs;
};

// ---- work schedule: every warp role walks the same list of units of its CTA
struct GemmUnit { int tile, kb0, kb1, kind; };  // kind 0: whole tile (or one atomic split-K slice); 1: stream-K continued
                                                // part (publish partial); 2: stream-K first part (collect partials)
struct GemmSched {
    int G, cta, KB, splits, T;   // grid, this CTA, K stages per tile, uniform split-K factor, output tiles
    int streamk, F;              // stream-K on the last T - F tiles (F = multiple of G tiles run whole)
    int Q;                       // (T - F) * KB: K stages of the stream-K region, CTA c owns [Q c / G, Q (c + 1) / G)
    __device__ __forceinline__ int lo(int c) const { return (int)((unsigned)Q * (unsigned)c / (unsigned)G); }  // Q G < 2^31 (host)
    __device__ __forceinline__ GemmUnit get(int i) const {   // kind < 0: no more work
        GemmUnit u;
        u.tile = 0; u.kb0 = 0; u.kb1 = 0; u.kind = -1;
        if (!streamk) {
            const int unit = cta + i * G;
            if (unit < T * splits) {
                const int sp = unit % splits;
                u.tile = unit / splits;
                u.kb0 = KB * sp / splits;   // KB <= 256
                u.kb1 = KB * (sp + 1) / splits;
                u.kind = 0;
            }
            return u;
        }
        const int l = lo(cta), h = lo(cta + 1);   // share < KB: the range touches at most two tiles
        const int t0 = l / KB, r0 = l - t0 * KB;
        const int b1 = (t0 + 1) * KB;
        const bool has_peer = r0 != 0;                   // continues a tile begun by an earlier CTA
        const bool has_own = has_peer ? h > b1 : h > l;  // begins a tile (finished by later CTAs)
        const int i_own = has_peer ? 1 : 0, i_dp = i_own + (has_own ? 1 : 0);
        if (has_peer && i == 0) {
            u.tile = F + t0; u.kb0 = r0; u.kb1 = (h < b1 ? h : b1) - t0 * KB; u.kind = 1;
        } else if (has_own && i == i_own) {
            const int t1 = has_peer ? t0 + 1 : t0;
            u.tile = F + t1; u.kb0 = 0; u.kb1 = h - t1 * KB; u.kind = 2;
        } else {
            const int tile = cta + (i - i_dp) * G;   // whole tiles of the full waves
            if (tile < F) { u.tile = tile; u.kb0 = 0; u.kb1 = KB; u.kind = 0; }
        }
        return u;
    }
};

__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int TN>
struct GemmSmem {
    static constexpr int kXHalfBytes = TN * 128;            // one TMA box: TN rows x 128 B
    static constexpr int kXStageBytes = 2 * kXHalfBytes;    // 256 K-bytes per row, as two swizzle atoms
    static constexpr int kStages = TN > 128 ? 3 : 4;        // X smem ring == TMEM A ring (64 columns per stage)
    static constexpr int kAccBufs = TN > 128 ? 1 : 2;       // TMEM accumulator buffers
    static constexpr int kBytes = kStages * kXStageBytes + kStagesW * kWStageBytes + 1024 /*align*/ + 512 /*bars*/ + 1024 /*per-token scales of the tile being stored*/;
};

template <int TN, int MODE, bool PROF = false, bool GROUPED = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_w4a8_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                 const GemmParams p) {
    static_assert(TN == 16 || TN == 32 || TN == 64 || TN == 128 || TN == 256, "token tile");
    constexpr int NBUF = GemmSmem<TN>::kAccBufs;
    constexpr int kStages = GemmSmem<TN>::kStages;
    constexpr int XH = GemmSmem<TN>::kXHalfBytes;
    constexpr int XB = GemmSmem<TN>::kXStageBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xs = smem;
    uint8_t* ws = xs + kStages * XB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ws + kStagesW * kWStageBytes);
    uint64_t* full = bars;                     // [kStages]
    uint64_t* empty = full + kStages;          // [kStages]
    uint64_t* wfull = empty + kStages;         // [kStagesW]
    uint64_t* wempty = wfull + kStagesW;       // [kStagesW]
    uint64_t* full_acc = wempty + kStagesW;    // [2]
    uint64_t* empty_acc = full_acc + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_acc + 2);
    float* sx_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [256], 16-B aligned

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    long long tacc[4] = {0, 0, 0, 0};
    const long long t_begin = PROF ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 5); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kStagesW; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(&full_acc[i], 1); mbar_init(&empty_acc[i], 16); }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_x);
    if (warp == 2 && lane == 0) tma_prefetch_desc(&tm_w);
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    grid_dep_launch();  // PDL: the next kernel may begin its prologue as SMs free up
    // everything above overlapped the previous kernel's tail; its outputs are visible after the wait.  The W producer
    // (warp 2) skips it: the packed weights are static, so their first stages stream in while the dependency resolves.
    // (a grouped launch reads its schedule from memory written by the preceding kernels: every role waits)
    if (GROUPED || warp != 2) grid_dep_wait();

    const int nb_w = (p.N + kGemmBlockW - 1) / kGemmBlockW;
    const int nb_m = (p.M + TN - 1) / TN;
    // tile order: token blocks fastest, so the CTAs running concurrently share few weight tiles (each packed W tile
    // is fetched from HBM once and re-read from L2 by its other token blocks) while all of X stays L2-resident
    const bool w_fast = PROF && (p.dbg & 32);  // debug: the other raster order (weight blocks fastest)
    const int KB_all = (p.K + kGemmBlockK - 1) / kGemmBlockK;
    const int splits = p.k_splits > 1 ? p.k_splits : 1;
    GemmSched sc;
    sc.G = (int)gridDim.x; sc.cta = (int)blockIdx.x; sc.KB = KB_all; sc.splits = splits; sc.T = nb_w * nb_m;
    if constexpr (GROUPED) {
        const int t_dev = __ldg(p.grp_tile + p.n_groups);   // tiles of this routing; never more than the host bound
        sc.T = t_dev < p.t_upper ? t_dev : p.t_upper;
    }
    sc.streamk = p.streamk;
    sc.F = sc.T - sc.T % sc.G;
    sc.Q = (sc.T - sc.F) * KB_all;
    // setmaxnreg sits at the top of each warpgroup's own branch so that ptxas allocates every role's code against
    // that role's budget
    if (warp < 4) {
      reg_dealloc<kRegsCtl>();
      if (warp == 0) {
        // ================================ X producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int ui = 0;; ++ui) {
                const GemmUnit u = sc.get(ui);
                if (u.kind < 0) break;
                int m0;
                if constexpr (GROUPED) m0 = grouped_tile(p.grp_off, p.grp_tile, p.n_groups, TN, kGemmBlockW, u.tile).m0;
                else m0 = (w_fast ? u.tile / nb_w : u.tile % nb_m) * TN;
                for (int kb = u.kb0; kb < u.kb1; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    FLATQ_TWAIT(0, mbar_wait(&empty[s], ph ^ 1));
                    if (PROF && (p.dbg & 8)) { mbar_arrive(&full[s]); continue; }
                    mbar_arrive_expect_tx(&full[s], XB);
                    tma_load_2d(xs + s * XB, &tm_x, &full[s], kb * kGemmBlockK, m0);
                    tma_load_2d(xs + s * XB + XH, &tm_x, &full[s], kb * kGemmBlockK + 128, m0);
                }
            }
        }
      } else if (warp == 2) {
        // ================================ W producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int ui = 0;; ++ui) {
                const GemmUnit u = sc.get(ui);
                if (u.kind < 0) break;
                int n0;  // row of the (stacked) packed weight
                if constexpr (GROUPED) {
                    const GroupTile gt = grouped_tile(p.grp_off, p.grp_tile, p.n_groups, TN, kGemmBlockW, u.tile);
                    n0 = gt.group * p.N + gt.n0;
                } else {
                    n0 = (w_fast ? u.tile % nb_w : u.tile / nb_m) * kGemmBlockW;
                }
                for (int kb = u.kb0; kb < u.kb1; ++kb, ++it) {
                    const uint32_t s = it % kStagesW, ph = (it / kStagesW) & 1;
                    FLATQ_TWAIT(0, mbar_wait(&wempty[s], ph ^ 1));
                    if (PROF && (p.dbg & 16)) { mbar_arrive(&wfull[s]); continue; }
                    mbar_arrive_expect_tx(&wfull[s], kWStageBytes);
                    tma_load_2d(ws + s * kWStageBytes, &tm_w, &wfull[s], kb * (kGemmBlockK / 2), n0);
                }
            }
        }
      } else if (warp == 1) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc = make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, kGemmBlockW, TN);
        uint32_t it = 0, tile_iter = 0;
        for (;; ++tile_iter) {
            const GemmUnit u = sc.get((int)tile_iter);
            if (u.kind < 0) break;
            const int kb0 = u.kb0, kb1 = u.kb1;
            const uint32_t b = tile_iter % NBUF, phb = (tile_iter / NBUF) & 1;
            FLATQ_TWAIT(2, mbar_wait(&empty_acc[b], phb ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem + b * 128;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                FLATQ_TWAIT(0, mbar_wait(&full[s], ph));
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t xaddr = smem_u32(xs + s * XB);
                    const uint32_t a_tmem = tmem + kTmemColsA0 + s * 64;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (PROF && (p.dbg & 2)) break;
                        // K sub-block kk: 32 bytes inside swizzle atom (kk / 4)
                        const uint64_t bdesc =
                            make_smem_desc_kmajor(xaddr + (kk >> 2) * XH + (kk & 3) * 32, 1024, kLayoutSw128);
                        mma_i8_ts(d_tmem, a_tmem + kk * 8, bdesc, idesc, ((kb - kb0) | kk) != 0);
                    }
                    tc_commit(&empty[s]);
                    if (kb == kb1 - 1) tc_commit(&full_acc[b]);
                }
                __syncwarp();
            }
        }
      }  // warp 3 idles: it only completes the control warpgroup for setmaxnreg
    } else if (warp < 12) {
        // ================================ nibble expanders ================================
        if constexpr (kRegsExpand < 72) reg_dealloc<kRegsExpand>();
        const int grp = (warp - 4) >> 2;  // handles stage iterations with it % 2 == grp
        const int q = warp & 3;           // TMEM lane quadrant this warp may touch
        const uint32_t row = q * 32 + lane;
        // SWIZZLE_128B: 16-B chunk index ^= (row % 8).  Stage bases are 1024-B aligned, so chunk c of this row lives at
        // (row base | (row % 8) << 4) ^ (c << 4): one XOR with an immediate per load, no per-chunk address registers
        const uint32_t ws_addr = smem_u32(ws) + row * 128 + ((row & 7) << 4);
        uint32_t it = 0;
        for (int ui = 0;; ++ui) {
                const GemmUnit u = sc.get(ui);
                if (u.kind < 0) break;
            for (int kb = u.kb0; kb < u.kb1; ++kb, ++it) {
                if ((it & 1) != (uint32_t)grp) continue;
                const uint32_t sw = it % kStagesW, phw = (it / kStagesW) & 1;
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                FLATQ_TWAIT(0, mbar_wait(&wfull[sw], phw));
                uint4 v[8];  // this row's 128 packed bytes, fetched while the A slot may still be busy
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (PROF && (p.dbg & 16)) v[c] = make_uint4(it, c, lane, 7);
                    else v[c] = ld_shared_v4((ws_addr + sw * kWStageBytes) ^ (c << 4));
                }
                FLATQ_TWAIT(1, mbar_wait(&empty[s], ph ^ 1));
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + kTmemColsA0 + s * 64;
#pragma unroll
                for (int h = 0; h < 4; ++h) {  // 16 TMEM columns (64 weights) per store
                    uint32_t a[16];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint4 u = v[2 * h + c];
                        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            a[c * 8 + 2 * j] = (w4[j] << 4) & 0xF0F0F0F0u;  // nibbles 0,2,4,6 -> 16*w as int8
                            a[c * 8 + 2 * j + 1] = w4[j] & 0xF0F0F0F0u;     // nibbles 1,3,5,7
                        }
                    }
                    if (!(PROF && (p.dbg & 1))) tmem_st_x16(taddr + h * 16, a);
                }
                // every packed byte of this row has been consumed by the ALU ops above: release the W slot
                __syncwarp();
                if (lane == 0) mbar_arrive(&wempty[sw]);
                FLATQ_TWAIT(2, tmem_st_wait());
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
    } else {
        // ================================ epilogue ================================
        reg_alloc<kRegsEpi>();
        const int ew = warp - 12;          // 0..15
        const int q = warp & 3;            // TMEM lane quadrant
        const int half = ew >> 2;          // which quarter of the token columns
        uint32_t tile_iter = 0;
        constexpr int COLS = TN / 4 < 16 ? 16 : TN / 4;   // token columns per epilogue warp (small TN: some quarters idle)
        for (;; ++tile_iter) {
            const GemmUnit u = sc.get((int)tile_iter);
            if (u.kind < 0) break;
            int n0, m0, m_end, ch0;  // ch0: offset of the group's per-channel scale / bias rows
            if constexpr (GROUPED) {
                const GroupTile gt = grouped_tile(p.grp_off, p.grp_tile, p.n_groups, TN, kGemmBlockW, u.tile);
                n0 = gt.n0; m0 = gt.m0; m_end = gt.m_end < p.M ? gt.m_end : p.M; ch0 = gt.group * p.N;
            } else {
                n0 = (w_fast ? u.tile % nb_w : u.tile / nb_m) * kGemmBlockW;
                m0 = (w_fast ? u.tile / nb_w : u.tile % nb_m) * TN;
                m_end = p.M; ch0 = 0;
            }
            const uint32_t b = tile_iter % NBUF, phb = (tile_iter / NBUF) & 1;
            const int n = n0 + q * 32 + lane;
            const bool n_ok = n < p.N;
            const int c0 = half * COLS;
            float sw_n = 0.f, bias_n = 0.f;
            if (MODE != kEpiI32) {
                if (n_ok) {
                    sw_n = p.s_w[ch0 + n];
                    if (p.bias) {
                        bias_n = MODE == kEpiF16 ? __half2float(reinterpret_cast<const __half*>(p.bias)[ch0 + n])
                                                 : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[ch0 + n]);
                    }
                }
            }
            // Per-token scales of this quarter's columns go through shared memory: one broadcast LDS.128 then serves
            // four columns (a SHFL per column costs as much shared-memory bandwidth as an MMA operand row, and the
            // tensor pipe needs that bandwidth).  Staged while the MMAs of this tile still run.
            const bool fold = p.acc_shift == 4 && p.K <= 16384;
            if (MODE != kEpiI32 && c0 < TN) {
                named_bar_sync(1 + half, 128);   // the quarter's four warps have finished reading the previous tile's scales
                if (q == 0) {
#pragma unroll
                    for (int i = lane; i < COLS; i += 32) {
                        const int m = m0 + c0 + i;
                        const float v = m < m_end ? p.s_x[m] : 0.f;
                        sx_s[c0 + i] = fold ? v * 0.0625f : v;
                    }
                }
                named_bar_sync(1 + half, 128);   // staged
            }
            FLATQ_TWAIT(0, mbar_wait(&full_acc[b], phb));
            tc_fence_after();
            const long long t_epi0 = PROF ? clock64() : 0;
            const uint32_t t_acc = tmem + ((uint32_t)(q * 32) << 16) + b * 128 + c0;
            const bool live = c0 < TN && !(PROF && (p.dbg & 4));
            if (MODE == kEpiI32) {
                if (live) {
#pragma unroll 1
                    for (int c = 0; c < COLS / 16; ++c) {
                        uint32_t r[16];
                        tmem_ld_x16(t_acc + c * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int m = m0 + c0 + c * 16 + j;
                            if (m < m_end && n_ok) {
                                int32_t* dst = reinterpret_cast<int32_t*>(p.out) + (size_t)m * p.N + n;
                                const int32_t val = static_cast<int32_t>(r[j]) >> p.acc_shift;  // exact per partial sum
                                if (splits > 1) atomicAdd(dst, val); else *dst = val;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_acc[b]);
            } else {
                // Copy this warp's COLS raw accumulator columns into registers with one batch of tcgen05.ld, then
                // release the TMEM buffer: the next tile's MMAs overlap all the arithmetic and the stores below.
                uint32_t r[COLS];
                if (live) {
                    if constexpr (COLS == 16) {
                        tmem_ld_x16(t_acc, r);
                    } else {
#pragma unroll
                        for (int c = 0; c < COLS / 32; ++c) tmem_ld_x32(t_acc + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
                    }
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_acc[b]);
                if (u.kind == 1) {
                    // stream-K continued part: publish this warp's slice of the raw partial ([column][row], 128-B rows)
                    if (live) {
                        int32_t* part = p.sk_part + (size_t)blockIdx.x * (kGemmBlockW * TN) + c0 * kGemmBlockW + q * 32 + lane;
#ifndef FLATQ_SK_NO_WRITE
#pragma unroll
                        for (int j = 0; j < COLS; ++j) part[j * kGemmBlockW] = static_cast<int32_t>(r[j]);
                        __threadfence();
#endif
                        __syncwarp();
                        if (lane == 0) st_release_u64(&p.sk_flags[blockIdx.x * 16 + ew], p.epoch);
                    }
                    continue;
                }
                if (u.kind == 2 && live) {
                    // stream-K first part: the following CTAs continue this tile; add their partials (published early
                    // in their own schedules) before dequantising
                    const int tile_end = (u.tile - sc.F + 1) * sc.KB;
                    for (int c2 = sc.cta + 1; sc.lo(c2) < tile_end; ++c2) {
                        if (lane == 0) {
                            // peers publish in the first unit of their schedule; a wait of ~1 s means a broken launch
                            // (e.g. a workspace shared by concurrent launches): fail loudly rather than hang
                            uint32_t spins = 0;
                            while (ld_acquire_u64(&p.sk_flags[c2 * 16 + ew]) != p.epoch) {
                                __nanosleep(64);
                                if (++spins > (1u << 24)) asm volatile("trap;");
                            }
                        }
                        __syncwarp();
                        const int32_t* part = p.sk_part + (size_t)c2 * (kGemmBlockW * TN) + c0 * kGemmBlockW + q * 32 + lane;
#ifndef FLATQ_SK_NO_READ
#pragma unroll
                        for (int j = 0; j < COLS; ++j) r[j] += static_cast<uint32_t>(__ldcg(part + j * kGemmBlockW));
#else
                        (void)part;
#endif
                    }
                }
                if (live) {
                    // fast path (fold): the raw accumulator (16x the true one) is exactly representable in fp32
                    // (|raw| <= 2^10 K), so the power-of-two factor sits in the staged per-token scale instead of a shift
                    const bool has_bias = p.bias != nullptr;
                    // 16-bit stores: lane = output feature n, so a warp store covers 64 contiguous bytes of one token row.
                    // (M + 256) * N < 2^32 is checked on the host.
                    uint16_t* const out16 = reinterpret_cast<uint16_t*>(p.out);
                    uint32_t idx = (uint32_t)(m0 + c0) * (uint32_t)p.N + (uint32_t)n;
                    const int j_end = n_ok ? m_end - (m0 + c0) : 0;  // store only tokens m0 + c0 + j < M (< the group's end)
                    auto finish = [&](auto fold_c) {
                        constexpr bool FOLD = decltype(fold_c)::value;
#pragma unroll
                        for (int j = 0; j < COLS; j += 4) {
                            const float4 s4 = *reinterpret_cast<const float4*>(&sx_s[c0 + j]);
                            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
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
            }
            if (PROF) tacc[1] += clock64() - t_epi0;
            __syncwarp();
        }
    }

    if (PROF && lane == 0 && p.prof) {
        long long* o = p.prof + ((size_t)blockIdx.x * kGemmWarps + warp) * 5;
        o[0] = tacc[0]; o[1] = tacc[1]; o[2] = tacc[2]; o[3] = tacc[3]; o[4] = clock64() - t_begin;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// Tile prefix of a grouped launch, built on the device from the router's row offsets (no host read of the routing):
// the parallel form of grouped_build_prefix (csrc/grouped_sched.h) -- one block, one thread per group, a shared-memory
// inclusive scan per chunk of 1024 groups (a serial walk over 256 experts would cost ~500 dependent loads).
constexpr int kPrefixThreads = 1024;
__global__ void __launch_bounds__(kPrefixThreads)
grouped_prefix_kernel(const int* __restrict__ off, int n_groups, int tn, int nb_w, int* __restrict__ tile_prefix) {
    __shared__ int s[kPrefixThreads];
    __shared__ int base;
    const int t = threadIdx.x;
    if (t == 0) { base = 0; tile_prefix[0] = 0; }
    __syncthreads();
    for (int c0 = 0; c0 < n_groups; c0 += kPrefixThreads) {
        const int e = c0 + t;
        int v = 0;
        if (e < n_groups) {
            int rows = off[e + 1] - off[e];
            if (rows < 0) rows = 0;
            v = nb_w * ((rows + tn - 1) / tn);
        }
        s[t] = v;
        __syncthreads();
        for (int d = 1; d < kPrefixThreads; d <<= 1) {   // Hillis-Steele inclusive scan
            const int add = t >= d ? s[t - d] : 0;
            __syncthreads();
            s[t] += add;
            __syncthreads();
        }
        if (e < n_groups) tile_prefix[e + 1] = base + s[t];
        __syncthreads();
        if (t == kPrefixThreads - 1) base += s[t];
        __syncthreads();
    }
}

}  // namespace flatq
