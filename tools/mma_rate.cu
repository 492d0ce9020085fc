// Microbenchmark: issue rate of tcgen05.mma per kind / shape / operand source on this GPU, and an
// exactness probe of integer-valued FP8 accumulation.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_rate tools/mma_rate.cu && /tmp/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../flatquant_b200/csrc/ptx.cuh"
using namespace flatq;

__device__ __forceinline__ void mma_f8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 i8 TS, 1 i8 SS, 2 f8 SS, 3 f8 TS, 4 f16 SS
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n, int reps, long long* out, int nacc = 1) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint64_t adesc = make_smem_desc_kmajor(smem_u32(smem), 1024, kLayoutSw128);
        const uint64_t bdesc = make_smem_desc_kmajor(smem_u32(smem + 16384), 1024, kLayoutSw128);
        uint32_t idesc;
        if (mode <= 1) idesc = make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, 128, n);
        else if (mode <= 3) idesc = make_idesc(kDF32, 0 /*e4m3*/, 0, 0, 0, 128, n);
        else idesc = make_idesc(kDF32, kFmtF16, kFmtF16, 0, 0, 128, n);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t koff = (r & 3) * 2;  // advance K inside the atom like a real loop
            const uint32_t d = tmem + (uint32_t)(r % nacc) * (uint32_t)n;   // nacc > 1: independent accumulators, round robin
            const uint32_t a_t = tmem + (nacc > 1 ? 448 : 256) + (r & 7) * 8;
            if (mode == 0) mma_i8_ts(d, a_t, bdesc + koff, idesc, r >= nacc);
            else if (mode == 1) mma_i8_ss(d, adesc + koff, bdesc + koff, idesc, r >= nacc);
            else if (mode == 2) mma_f8_ss(d, adesc + koff, bdesc + koff, idesc, r >= nacc);
            else if (mode == 3) mma_f8_ts(d, a_t, bdesc + koff, idesc, r >= nacc);
            else mma_f16_ss(d, adesc + koff, bdesc + koff, idesc, r >= nacc);
        }
        const long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}


// ---- issue-pattern probe: i8 TS M128 N256, `per` MMAs then commit (+ optional wait/fence), with `stw` warps
// hammering tcgen05.st into the A region like the nibble expanders do.
__global__ void __launch_bounds__(32 * 9, 1) pattern_kernel(int per, int do_wait, int stw, int reps, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[8];
    __shared__ uint64_t done;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (49152) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); stop = 0; fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const uint64_t bdesc = make_smem_desc_kmajor(smem_u32(smem + 16384), 1024, kLayoutSw128);
        const uint32_t idesc = make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, 128, 256);
        const long long t0 = clock64();
        int issued = 0;
        for (int it = 0; issued < reps; ++it) {
            const int s = it & 7;
            if (do_wait && it >= 8) { mbar_wait(&bar[s], ((it >> 3) - 1) & 1); tc_fence_after(); }
            for (int k = 0; k < per; ++k, ++issued)
                mma_i8_ts(tmem, tmem + 256 + s * 32 + (k & 3) * 8, bdesc + (k & 3) * 2, idesc, issued != 0);
            tc_commit(&bar[s]);
        }
        tc_commit(&done);
        mbar_wait(&done, 0);
        const long long t2 = clock64();
        stop = 1;
        if (blockIdx.x == 0) { out[0] = t2 - t0; out[1] = issued; }
    } else if (warp >= 1 && warp <= stw) {
        uint32_t a[16];
        for (int i = 0; i < 16; ++i) a[i] = i * 0x01010101u;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + 160;  // columns outside the MMA's A stages
        while (!stop) {
            tmem_st_x16(taddr, a);
            tmem_st_x16(taddr + 16, a);
            tmem_st_wait();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// ---- exactness: D = sum over `reps` k-blocks of A(128x128 e4m3 ints) . B(128x128)^T, fp32 accumulate
__host__ __device__ inline uint8_t e4m3_of_int(int v) {  // |v| <= 8 exactly representable
    if (v == 0) return 0;
    const uint8_t s = v < 0 ? 0x80 : 0;
    int a = v < 0 ? -v : v, e = 0;
    while ((a >> (e + 1)) != 0) ++e;               // floor(log2 a)
    const int mant = ((a << 3) >> e) & 7;          // 3 mantissa bits (exact for a <= 15, and 16)
    return s | (uint8_t)(((e + 7) << 3) | mant);
}

__global__ void __launch_bounds__(128, 1) exact_kernel(const uint8_t* a_g, const uint8_t* b_g, int reps, int use_i8, float* d_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    // operands: row-major [128][128] bytes -> SW128 K-major tile
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(a_g)[i];
        *reinterpret_cast<uint4*>(smem + 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(b_g)[i];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint64_t adesc = make_smem_desc_kmajor(smem_u32(smem), 1024, kLayoutSw128);
        const uint64_t bdesc = make_smem_desc_kmajor(smem_u32(smem + 16384), 1024, kLayoutSw128);
        const uint32_t idesc = use_i8 ? make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, 128, 128) : make_idesc(kDF32, 0, 0, 0, 0, 128, 128);
        for (int r = 0; r < reps; ++r)
            for (int kk = 0; kk < 4; ++kk) {
                if (use_i8) mma_i8_ss(tmem, adesc + kk * 2, bdesc + kk * 2, idesc, (r | kk) != 0);
                else mma_f8_ss(tmem, adesc + kk * 2, bdesc + kk * 2, idesc, (r | kk) != 0);
            }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j)
            d_out[(warp * 32 + lane) * 128 + c + j] = use_i8 ? (float)(int)r[j] : __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    const char* names[] = {"i8  TS", "i8  SS", "f8  SS", "f8  TS", "f16 SS"};
    const int kper[] = {32, 32, 32, 32, 16};
    const int reps = 4096;
    for (int mode = 0; mode < 5; ++mode)
        for (int n : {64, 128, 256}) {
            for (int grid : {1, 148}) {
                rate_kernel<<<grid, 128, 60 * 1024>>>(mode, n, reps, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                const double cyc = (double)h[1] / reps;
                printf("%s M128 N%3d K%d grid %3d: issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/cyc/SM %s\n", names[mode], n,
                       kper[mode], grid, (double)h[0] / reps, cyc, 128.0 * n * kper[mode] / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    // round 2: is the ~128-cycle cost of a small-N MMA an issue floor, or the latency of accumulating into the SAME
    // accumulator?  Same loop, `nacc` independent accumulators (n columns each) round robin.
    for (int mode : {0, 1, 4})
        for (int n : {16, 32, 64, 128})
            for (int nacc : {1, 2, 4}) {
                if (n * nacc > 448) continue;
                rate_kernel<<<148, 128, 60 * 1024>>>(mode, n, reps, out, nacc);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                const double cyc = (double)h[1] / reps;
                printf("%s M128 N%3d K%d nacc %d: issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/cyc/SM %s\n", names[mode], n,
                       kper[mode], nacc, (double)h[0] / reps, cyc, 128.0 * n * kper[mode] / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    for (int per : {4, 8, 16, 4096})
        for (int do_wait : {0, 1})
            for (int stw : {0, 4, 8}) {
                pattern_kernel<<<148, 32 * 9, 60 * 1024>>>(per, do_wait, stw, 4096, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("pattern i8 TS N256: %4d MMAs/commit wait=%d st_warps=%d: %.1f cyc/mma %s\n", per, do_wait, stw,
                       (double)h[0] / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    // exactness of integer-valued fp8 accumulation
    uint8_t *ha = (uint8_t*)malloc(16384), *hb = (uint8_t*)malloc(16384), *ha8 = (uint8_t*)malloc(16384), *hb8 = (uint8_t*)malloc(16384);
    int* ia = (int*)malloc(16384 * 4); int* ib = (int*)malloc(16384 * 4);
    srand(1);
    for (int i = 0; i < 16384; ++i) {
        ia[i] = rand() % 16 - 8; ib[i] = rand() % 16 - 8;
        ha[i] = e4m3_of_int(ia[i]); hb[i] = e4m3_of_int(ib[i]);
        ha8[i] = (uint8_t)(int8_t)ia[i]; hb8[i] = (uint8_t)(int8_t)ib[i];
    }
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 16384); cudaMalloc(&dd, 128 * 128 * 4);
    float* hd = (float*)malloc(128 * 128 * 4);
    for (int use_i8 = 0; use_i8 < 2; ++use_i8)
        for (int repsk : {1, 32, 112, 512}) {
            cudaMemcpy(da, use_i8 ? ha8 : ha, 16384, cudaMemcpyHostToDevice);
            cudaMemcpy(db, use_i8 ? hb8 : hb, 16384, cudaMemcpyHostToDevice);
            exact_kernel<<<1, 128, 60 * 1024>>>(da, db, repsk, use_i8, dd);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hd, dd, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            long long bad = 0; double maxabs = 0, maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 128; ++n) {
                    long long acc = 0;
                    for (int k = 0; k < 128; ++k) acc += (long long)ia[m * 128 + k] * ib[n * 128 + k];
                    const double want = (double)acc * repsk, got = hd[m * 128 + n];
                    if (want != got) ++bad;
                    if (fabs(want) > maxabs) maxabs = fabs(want);
                    if (fabs(want - got) > maxerr) maxerr = fabs(want - got);
                }
            printf("exactness %s K=%6d: mismatches %lld / 16384, max|acc| %.0f, max err %.1f %s\n", use_i8 ? "i8" : "e4m3", repsk * 128, bad,
                   maxabs, maxerr, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
