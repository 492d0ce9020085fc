// One-launch FlatLinear for decode (M <= 2 tokens): online Kronecker transform -> per-token INT4 quantise -> weight
// stream GEMV -> dequant (+bias) [-> SiLU(gate) * up] in ONE kernel.
//
// Replaces, for one or two tokens, the pair (kron_quant_ts_kernel, gemv_w4a8_pre_kernel) -- i.e. the reference's
// OnlineTrans -> Quantizer -> Linear4bit chain (deploy/nn/online_trans.py:55-67, quantization.py:13-36, linear.py:40-54;
// flatquant/flat_linear.py:75-80).  In a decode step those two launches are a dependent pair of microsecond kernels:
// the transform runs on ONE SM (3-4 us of TMA -> tcgen05 -> TMEM round trips for 0.5 MFLOP) while 147 SMs wait, then the
// GEMV starts.  Here EVERY CTA of the GEMV recomputes the transform of the token(s) itself, on mma.sync, while its
// first weight rows -- requested before anything else, they are static -- are in flight:
//   prologue (before griddepcontrol.wait): weight rows -> registers; L^T and R -> shared memory
//   after the wait: x -> shared memory; the eight warps share the token's [n1, n2] matrix as (16-row slab, column part):
//     Y = L^T X   (m16n8k16, fp32 accumulate, ONE fp16 rounding -- the rounding points of kron_quant_ts.cuh: left factor
//                  first), through shared memory to
//     Z = Y R     (fp32 accumulate, one fp16 rounding), kept in registers
//     max / min of Z -> shared; scale = max(|min c_min|, max c_max) / 7; codes = rint(Z / scale) clamped to [-8, 7] with
//     the exact-division sequence of kron_quant_ts.cuh; written as int8 in the GEMM operand's K order [0,2,4,6,1,3,5,7]
//   then the GEMV loop of gemv_w4a8.cuh over this CTA's rows, with s_x from shared memory.
// The products are the same fp16 -> fp32 sums as the tcgen05 kernel's, in a different order: Z can differ in the last
// fp16 bit on a few elements, so an activation code differs from the two-launch path about once in 10^3-10^4 (the same
// class of difference as between that kernel and the oracle); given its codes the result is exact.
// Redundant work: 2 n1 n2 (n1 + n2) FLOP per CTA (0.5 M at 64 x 64) and one L2 read of the factors; instantiated for
// K = 2048 / 4096 / 8192 -- at 112 x 128 (K = 14336: 6.9 MFLOP per CTA) the two-launch path measured faster and is kept.
#pragma once
#include "gemv_w4a8.cuh"
#include "kron_quant_tc.cuh"

namespace flatq {

struct FlatDecodeParams {
    GemvParams g;            // w, s_w, bias, out, M, N, K, bf16_out (x / s_x unused: produced in shared memory)
    const __half* x;         // [M, K] fp16 activations
    const __half* left;      // [n1, n1] row-major or null (identity)
    const __half* right;     // [n2, n2] row-major or null (identity)
    int n1, n2;
    float c_max, c_min;
    int8_t* dbg_codes;       // tests: CTA 0 dumps its codes [M, K] and scales [M] (null in production)
    float* dbg_sx;
};

__host__ __device__ inline size_t flat_decode_smem_bytes(int n1, int n2, int mt) {
    return ((size_t)n1 * (n1 + 8) + (size_t)n2 * (n2 + 8) + 2 * (size_t)n1 * (n2 + 8)) * 2 + (size_t)mt * n1 * n2 + 64 * 4;
}

template <int MT, int N2, int CPL, int R, int MINB, bool SILU>
__global__ void __launch_bounds__(kGemvWarps * 32, MINB) flatlinear_decode_kernel(const FlatDecodeParams q) {
    static_assert(!SILU || R == 2, "SiLU pairs: two rows per warp and pass");
    static_assert(N2 == 64 || N2 == 128, "right factor size");
    extern __shared__ __align__(16) uint8_t fd_smem[];
    const GemvParams& p = q.g;
    const int n1 = q.n1, ldl = n1 + 8;
    constexpr int ldr = N2 + 8;
    __half* LTs = reinterpret_cast<__half*>(fd_smem);            // [n1][ldl]  L^T (A of stage 1), row-major
    __half* Rs = LTs + (size_t)n1 * ldl;                         // [N2][ldr]  R   (B of stage 2), row-major [k][n]
    __half* Xs = Rs + (size_t)N2 * ldr;                          // [n1][ldr]  the token as [n1, N2] (B of stage 1)
    __half* Ys = Xs + (size_t)n1 * ldr;                          // [n1][ldr]  Y = L^T X (A of stage 2)
    uint8_t* xs = reinterpret_cast<uint8_t*>(Ys + (size_t)n1 * ldr);   // [MT][K] int8 codes, GEMM operand order
    float* red = reinterpret_cast<float*>(xs + (size_t)MT * p.K);      // [8 warps][2] max / min, then [32..33] scales
    grid_dep_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kGemvWarps + warp, nw = gridDim.x * kGemvWarps;
    const size_t row_bytes = (size_t)p.K / 2;
    const int units = SILU ? p.N / 2 : p.N;
    auto row_of = [&](int u0, int r) { return SILU ? 2 * u0 + r : u0 + r * nw; };
    uint4 wb[R][CPL];
    auto load_rows = [&](int u0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = row_of(u0, r);
            if (n < p.N) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) wb[r][j] = ldg_nc_v4(p.w + (size_t)n * row_bytes + (size_t)(lane + 32 * j) * 16);
            }
        }
    };
    if (gw < units) load_rows(gw);

    // ---- static factors -> shared memory (before the dependency wait)
    if (q.left) {   // transposed: LTs[i'][i] = L[i][i'] (16-byte reads of L's rows, 2-byte scattered stores)
        for (int e = threadIdx.x; e < n1 * (n1 / 8); e += blockDim.x) {
            const int i = e / (n1 / 8), ip0 = (e - i * (n1 / 8)) * 8;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(q.left + (size_t)i * n1 + ip0));
            const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
            for (int u = 0; u < 8; ++u) LTs[(size_t)(ip0 + u) * ldl + i] = h[u];
        }
    }
    if (q.right) {
        for (int e = threadIdx.x; e < N2 * (N2 / 8); e += blockDim.x) {
            const int r = e / (N2 / 8), c = (e - r * (N2 / 8)) * 8;
            *reinterpret_cast<uint4*>(Rs + (size_t)r * ldr + c) = __ldg(reinterpret_cast<const uint4*>(q.right + (size_t)r * N2 + c));
        }
    }
    grid_dep_wait();

    const int g = lane >> 2, tig = lane & 3, mi = lane >> 3, r8 = lane & 7;
    // Work items of the transform: (16-row slab, column part) -- all eight warps take part in BOTH products (a 64 x 64
    // token has four slabs: without the column split half of the CTA would idle through the latency-bound MMA chains).
    constexpr int N1 = CPL * 1024 / N2, SLABS = N1 / 16, NT = N2 / 16;
    constexpr int NSPLIT = SLABS >= 8 ? 1 : (SLABS >= 4 ? 2 : 4), NTP = NT / NSPLIT;
    constexpr int ITEMS = SLABS * NSPLIT, ROUNDS = (ITEMS + kGemvWarps - 1) / kGemvWarps;
    static_assert(N1 % 16 == 0 && NT % NSPLIT == 0, "factor sizes");
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
        if (m >= p.M) {   // the second token slot of a one-token call: zero codes
            for (int i = threadIdx.x; i < p.K / 16; i += blockDim.x) reinterpret_cast<uint4*>(xs + (size_t)m * p.K)[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        __syncthreads();   // Xs / Ys / red of the previous token are no longer read
        for (int e = threadIdx.x; e < N1 * (N2 / 8); e += blockDim.x) {
            const int r = e / (N2 / 8), c = (e - r * (N2 / 8)) * 8;
            *reinterpret_cast<uint4*>(Xs + (size_t)r * ldr + c) = *reinterpret_cast<const uint4*>(q.x + (size_t)m * p.K + (size_t)r * N2 + c);
        }
        __syncthreads();
        // ---- stage 1: Y = L^T X, one fp16 rounding, into Ys (identity left factor: Y is X itself)
        if (q.left) {
#pragma unroll
            for (int rd = 0; rd < ROUNDS; ++rd) {
                const int item = warp + rd * kGemvWarps;
                if (item < ITEMS) {
                    const int i0 = (item / NSPLIT) * 16, nt0 = (item % NSPLIT) * NTP;
                    uint32_t a[SLABS][4];
#pragma unroll
                    for (int kt = 0; kt < SLABS; ++kt)
                        ldsm_x4(smem_u32(LTs + (size_t)(i0 + (mi & 1) * 8 + r8) * ldl + kt * 16 + (mi >> 1) * 8), a[kt]);
#pragma unroll
                    for (int t = 0; t < NTP; ++t) {
                        const int nt = nt0 + t;
                        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int kt = 0; kt < SLABS; ++kt) {
                            uint32_t b[4];
                            ldsm_x4_t(smem_u32(Xs + (size_t)(kt * 16 + (mi & 1) * 8 + r8) * ldr + nt * 16 + (mi >> 1) * 8), b);
                            hmma16816(c0, a[kt], b[0], b[1]);
                            hmma16816(c1, a[kt], b[2], b[3]);
                        }
                        *reinterpret_cast<uint32_t*>(Ys + (size_t)(i0 + g) * ldr + nt * 16 + 2 * tig) = pack_h2(c0[0], c0[1]);
                        *reinterpret_cast<uint32_t*>(Ys + (size_t)(i0 + g + 8) * ldr + nt * 16 + 2 * tig) = pack_h2(c0[2], c0[3]);
                        *reinterpret_cast<uint32_t*>(Ys + (size_t)(i0 + g) * ldr + nt * 16 + 8 + 2 * tig) = pack_h2(c1[0], c1[1]);
                        *reinterpret_cast<uint32_t*>(Ys + (size_t)(i0 + g + 8) * ldr + nt * 16 + 8 + 2 * tig) = pack_h2(c1[2], c1[3]);
                    }
                }
            }
            __syncthreads();
        }
        const __half* Yp = q.left ? Ys : Xs;
        // ---- stage 2: Z = Y R, one fp16 rounding, kept in registers: zf[rd][t][0] row g cols 16 nt + 2 tig, +1; [1] row g + 8;
        //      [2], [3]: cols + 8 (identity right factor: Z is Y itself)
        uint32_t zf[ROUNDS][NTP][4];
        __half2 hmax = __float2half2_rn(0.f), hmin = hmax;   // hi = max(max Z, 0), lo = min(min Z, 0) (quant_utils.py:96-107)
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int item = warp + rd * kGemvWarps;
            if (item < ITEMS) {
                const int i0 = (item / NSPLIT) * 16, nt0 = (item % NSPLIT) * NTP;
                if (q.right) {
                    uint32_t a[NT][4];
#pragma unroll
                    for (int kt = 0; kt < NT; ++kt)
                        ldsm_x4(smem_u32(Yp + (size_t)(i0 + (mi & 1) * 8 + r8) * ldr + kt * 16 + (mi >> 1) * 8), a[kt]);
#pragma unroll
                    for (int t = 0; t < NTP; ++t) {
                        const int nt = nt0 + t;
                        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int kt = 0; kt < NT; ++kt) {
                            uint32_t b[4];
                            ldsm_x4_t(smem_u32(Rs + (size_t)(kt * 16 + (mi & 1) * 8 + r8) * ldr + nt * 16 + (mi >> 1) * 8), b);
                            hmma16816(c0, a[kt], b[0], b[1]);
                            hmma16816(c1, a[kt], b[2], b[3]);
                        }
                        zf[rd][t][0] = pack_h2(c0[0], c0[1]); zf[rd][t][1] = pack_h2(c0[2], c0[3]);
                        zf[rd][t][2] = pack_h2(c1[0], c1[1]); zf[rd][t][3] = pack_h2(c1[2], c1[3]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < NTP; ++t) {
                        const int nt = nt0 + t;
                        zf[rd][t][0] = *reinterpret_cast<const uint32_t*>(Yp + (size_t)(i0 + g) * ldr + nt * 16 + 2 * tig);
                        zf[rd][t][1] = *reinterpret_cast<const uint32_t*>(Yp + (size_t)(i0 + g + 8) * ldr + nt * 16 + 2 * tig);
                        zf[rd][t][2] = *reinterpret_cast<const uint32_t*>(Yp + (size_t)(i0 + g) * ldr + nt * 16 + 8 + 2 * tig);
                        zf[rd][t][3] = *reinterpret_cast<const uint32_t*>(Yp + (size_t)(i0 + g + 8) * ldr + nt * 16 + 8 + 2 * tig);
                    }
                }
#pragma unroll
                for (int t = 0; t < NTP; ++t) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 zz = *reinterpret_cast<const __half2*>(&zf[rd][t][j]);
                        hmax = __hmax2(hmax, zz);
                        hmin = __hmin2(hmin, zz);
                    }
                }
            }
        }
        float vmax = fmaxf(__low2float(hmax), __high2float(hmax));
        float vmin = fminf(__low2float(hmin), __high2float(hmin));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        if (lane == 0) { red[warp * 2] = vmax; red[warp * 2 + 1] = vmin; }
        __syncthreads();
        vmax = red[0]; vmin = red[1];
#pragma unroll
        for (int w = 1; w < kGemvWarps; ++w) { vmax = fmaxf(vmax, red[w * 2]); vmin = fminf(vmin, red[w * 2 + 1]); }
        const float amax = fmaxf(fabsf(vmin * q.c_min), vmax * q.c_max);
        const float scale = amax == 0.f ? 1.f : __fdiv_rn(amax, 7.0f);
        const float inv_s = __frcp_rn(scale);
        if (threadIdx.x == 0) red[32 + m] = scale;
        const uint64_t inv2 = pack_f32x2(inv_s, inv_s), nscale2 = pack_f32x2(-scale, -scale);
        const __half2 hi2 = __float2half2_rn(7.25f * scale), lo2 = __float2half2_rn(-8.25f * scale);   // see kron_quant_ts.cuh
        uint8_t* dst = xs + (size_t)m * p.K;
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int item = warp + rd * kGemvWarps;
            if (item < ITEMS) {
                const int i0 = (item / NSPLIT) * 16, nt0 = (item % NSPLIT) * NTP;
#pragma unroll
                for (int t = 0; t < NTP; ++t) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 zc = __hmax2(__hmin2(*reinterpret_cast<const __half2*>(&zf[rd][t][j]), hi2), lo2);
                        const float2 f = __half22float2(zc);
                        const uint64_t z2 = pack_f32x2(f.x, f.y);
                        uint64_t q0, rr, q1, mm;
                        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q0) : "l"(z2), "l"(inv2));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(nscale2), "l"(q0), "l"(z2));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q1) : "l"(rr), "l"(inv2), "l"(q0));   // correctly rounded z / s
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(mm) : "l"(q1), "l"(pack_f32x2(12582912.f, 12582912.f)));
                        uint32_t even, odd;
                        asm("mov.b64 {%0, %1}, %2;" : "=r"(even), "=r"(odd) : "l"(mm));
                        // element (row, col): row = i0 + g (+8 for j odd), col = 16 nt + (j >= 2 ? 8 : 0) + 2 tig (+1 for `odd`);
                        // inside its group of 8 the operand order is [0,2,4,6,1,3,5,7]: logical 2 tig -> byte tig, 2 tig + 1 -> byte 4 + tig
                        const int row = i0 + g + (j & 1) * 8, grp = row * N2 + (nt0 + t) * 16 + (j >> 1) * 8;
                        dst[grp + tig] = (uint8_t)even;
                        dst[grp + 4 + tig] = (uint8_t)odd;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (q.dbg_codes && blockIdx.x == 0) {
        for (int i = threadIdx.x; i < p.M * p.K; i += blockDim.x) q.dbg_codes[i] = (int8_t)xs[i];
        if ((int)threadIdx.x < p.M) q.dbg_sx[threadIdx.x] = red[32 + threadIdx.x];
    }

    // ---- the GEMV over this CTA's rows (gemv_w4a8_pre_kernel's loop; s_x from shared memory)
    const int stride = (SILU ? 1 : R) * nw;
    for (int u0 = gw; u0 < units; u0 += stride) {
        int32_t acc[R][MT];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int m = 0; m < MT; ++m) acc[r][m] = 0;
            if (row_of(u0, r) >= p.N) break;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const uint32_t pw[4] = {wb[r][j].x, wb[r][j].y, wb[r][j].z, wb[r][j].w};
                const int c = lane + 32 * j;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const uint4 x0 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32);
                    const uint4 x1 = *reinterpret_cast<const uint4*>(xs + (size_t)m * p.K + (size_t)c * 32 + 16);
                    const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    int32_t a = acc[r][m];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        a = __dp4a((int)((pw[k4] << 4) & 0xF0F0F0F0u), (int)xw[2 * k4], a);
                        a = __dp4a((int)(pw[k4] & 0xF0F0F0F0u), (int)xw[2 * k4 + 1], a);
                    }
                    acc[r][m] = a;
                }
            }
        }
        if (u0 + stride < units) load_rows(u0 + stride);
        float yv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = row_of(u0, r);
            yv[r] = 0.f;
            if (n >= p.N) break;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][m] += __shfl_xor_sync(0xffffffffu, acc[r][m], o);
            }
            if (lane < MT && lane < p.M) {
                int32_t s = acc[r][0];
#pragma unroll
                for (int m = 1; m < MT; ++m) if (lane == m) s = acc[r][m];
                float y = __fmul_rn(__fmul_rn((float)(s >> 4), red[32 + lane]), p.s_w[n]);
                if (SILU) {
                    yv[r] = y;
                } else {
                    if (p.bias) {
                        y = __fadd_rn(y, p.bf16_out ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                                    : __half2float(reinterpret_cast<const __half*>(p.bias)[n]));
                    }
                    if (p.bf16_out) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)lane * p.N + n] = __float2bfloat16_rn(y);
                    else reinterpret_cast<__half*>(p.out)[(size_t)lane * p.N + n] = __float2half_rn(y);
                }
            }
        }
        if (SILU && lane < MT && lane < p.M) {
            const float gf = __half2float(__float2half_rn(yv[0]));
            const __half act = __float2half_rn(gf / (1.f + expf(-gf)));
            reinterpret_cast<__half*>(p.out)[(size_t)lane * (p.N / 2) + u0] = __hmul(act, __float2half_rn(yv[R - 1]));
        }
    }
}

}  // namespace flatq
