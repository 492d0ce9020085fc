// Round-2 microbenchmark: cost of one tcgen05.mma as a function of N with a LEAN issue loop (straight-line, unrolled 16x,
// compile-time shapes).  tools/mma_rate.cu (round 1) selected the instruction through a run-time if-chain inside the
// issuing thread's loop and reported ~127 cycles per MMA whatever N; this probe separates the issue-loop overhead from
// the tensor pipe.  NACC independent accumulators round robin (N columns each).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_rate2 tools/mma_rate2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../flatquant_b200/csrc/ptx.cuh"
using namespace flatq;

// MODE 0: i8 SS, 1: i8 TS (A from TMEM), 2: f16 SS, 3: f16 TS
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate2_kernel(int reps16, long long* out) {
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
        constexpr uint32_t idesc = MODE < 2 ? make_idesc(kDS32, kFmtS8, kFmtS8, 0, 0, 128, N)
                                            : make_idesc(kDF32, kFmtF16, kFmtF16, 0, 0, 128, N);
        const uint32_t a_t = tmem + 448;
        const long long t0 = clock64();
        for (int r = 0; r < reps16; ++r) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint32_t d = tmem + (u % NACC) * N;
                const uint32_t acc = (r | (u >= NACC)) ? 1u : 0u;
                if (MODE == 0) mma_i8_ss(d, adesc + (u & 3) * 2, bdesc + (u & 3) * 2, idesc, acc);
                else if (MODE == 1) mma_i8_ts(d, a_t + (u & 7) * 8, bdesc + (u & 3) * 2, idesc, acc);
                else if (MODE == 2) mma_f16_ss(d, adesc + (u & 3) * 2, bdesc + (u & 3) * 2, idesc, acc);
                else mma_f16_ts(d, a_t + (u & 7) * 8, bdesc + (u & 3) * 2, idesc, acc);
            }
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

template <int MODE, int N, int NACC>
void run(long long* out) {
    static const char* names[] = {"i8  SS", "i8  TS", "f16 SS", "f16 TS"};
    const int kper = MODE < 2 ? 32 : 16;
    const int reps16 = 512;
    cudaFuncSetAttribute(rate2_kernel<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    rate2_kernel<MODE, N, NACC><<<148, 128, 60 * 1024>>>(reps16, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double n = reps16 * 16.0;
    printf("%s M128 N%3d K%d nacc %d: issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/cyc/SM %s\n", names[MODE], N, kper, NACC,
           h[0] / n, h[1] / n, 128.0 * N * kper / (h[1] / n), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int MODE>
void sweep(long long* out) {
    run<MODE, 16, 1>(out); run<MODE, 16, 4>(out);
    run<MODE, 32, 1>(out); run<MODE, 32, 4>(out);
    run<MODE, 64, 1>(out); run<MODE, 64, 2>(out); run<MODE, 64, 4>(out);
    run<MODE, 128, 1>(out); run<MODE, 128, 2>(out);
    run<MODE, 256, 1>(out);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    sweep<0>(out); sweep<1>(out); sweep<2>(out); sweep<3>(out);
    return 0;
}
