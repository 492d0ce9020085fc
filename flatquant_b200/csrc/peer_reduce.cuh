// The reduction behind a row-parallel FlatLinear over NVLink 5 / NVSwitch peer memory (the reference
// issues dist.all_reduce after F.linear: flatquant/model_tools/deepseekv3_utils.py:98-105, vllm_custom/.../
// llama_flatquant.py:117-126,261-289).  No NCCL on the data path: every rank maps every peer's buffers (symmetric
// memory) plus one NVSwitch MULTICAST mapping of them, and the kernels below move the bytes themselves.
//
// Prefill (M % (256 tp) == 0) -- reduce-scatter FUSED INTO THE GEMM, then one broadcast kernel:
//   * rank r owns token rows [r M/tp, (r+1) M/tp).  gemm_w4a8_2sm_kernel's epilogue stores every output tile straight
//     into the OWNER's staging buffer, slot `rank` (st.global through the mapped peer pointer, 64 contiguous bytes per
//     warp store): the reduce-scatter traffic leaves the SM tile by tile while the tensor pipe works on the next tile.
//   * peer_reduce_bcast_kernel: barrier; the owner adds the tp slots of its rows in fp32 (local HBM reads), adds the
//     bias once, rounds to fp16 and stores the result into EVERY rank's output through the peer pointers; barrier.
//     NVLink bytes per rank and linear: (tp-1)/tp S out during the GEMM (hidden under the MMAs), (tp-1)/tp S out and
//     in during the broadcast (S = M N 2 bytes) -- half of the exposed traffic of an all-reduce after the GEMM.
// Any other shape (decode sizes, ragged M) -- peer_allreduce_nvls_kernel on the GEMM's local output:
//   barrier; rank r pulls ITS 1/tp of the elements already summed by the switch (multimem.ld_reduce.add, fp32
//   accumulation), adds the bias, multimem.st to every rank; barrier.
//
// Barriers: epoch flags in each rank's signal pad, posted with st.release.sys and awaited on local memory (peer_barrier).
#pragma once
#include "ptx.cuh"

namespace flatq {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerThreads = 512;

struct PeerCtx {
    int rank, world;
    int pad_ctas;                           // CTA slots of the signal pad: the layout below does NOT depend on a launch's grid
    uint32_t* signal_pads[kPeerMaxWorld];   // each rank's signal pad (zero-initialised symmetric memory), >= 2 * pad_ctas * (world + 1) words
};

// Barrier over the ranks, one per CTA index: flags are EPOCH counters, never reset.  Every CTA keeps the number of
// barriers it has passed in its own (local) counter word; a barrier bumps it to e, posts e into slot [blk][rank] of
// every peer's pad with one st.release.sys (a posted write: no NVLink round trip), and spins on its OWN pad (local L2)
// until every peer's slot has reached e.  All ranks issue the same sequence of launches, so the counters stay in step;
// nothing is baked into launch arguments, so a CUDA graph can replay the launch.  Pad layout per rank (u32 words):
//   [2 phases][pad_ctas][world] flags, then [2 phases][pad_ctas] local counters -- indexed by CTA number with the pad's own
//   stride, so launches of different grid sizes (a prefill's broadcast, a decode step's all-reduce: one reducer serves
//   both) share the pad without aliasing each other's slots; every launch uses grid <= pad_ctas.
__device__ __forceinline__ void peer_barrier(const PeerCtx& c, int phase) {
    __shared__ uint32_t epoch_s;
    __syncthreads();
    uint32_t* pad = c.signal_pads[c.rank];
    const size_t slot = ((size_t)phase * c.pad_ctas + blockIdx.x) * c.world;
    if (threadIdx.x == 0) {
        uint32_t* ctr = pad + (size_t)2 * c.pad_ctas * c.world + (size_t)phase * c.pad_ctas + blockIdx.x;
        epoch_s = *ctr + 1;
        *ctr = epoch_s;
    }
    __syncthreads();
    const uint32_t e = epoch_s;
    if ((int)threadIdx.x < c.world) {
        uint32_t* remote = c.signal_pads[threadIdx.x] + slot + c.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
        const uint32_t* mine = pad + slot + threadIdx.x;
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        } while ((int32_t)(v - e) < 0);
    }
    __syncthreads();
}

__device__ __forceinline__ uint4 multimem_ld_reduce_f16x8(const void* mc_addr) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc_addr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

__device__ __forceinline__ uint4 add_bias_f16x8(uint4 v, const __half* bias, size_t col) {
    const uint4 b = *reinterpret_cast<const uint4*>(bias + col);
    __half2* vh = reinterpret_cast<__half2*>(&v);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) vh[j] = __hadd2(vh[j], bh[j]);
    return v;
}

// In-switch all-reduce of a symmetric fp16 buffer [rows, N] (N % 8 == 0): src_mc / dst_mc are the MULTICAST addresses
// of the partial and of the result buffer (they may be the same buffer).
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_nvls_kernel(const PeerCtx c, const __half* src_mc, __half* dst_mc, const __half* bias, long long rows, int N) {
    peer_barrier(c, 0);                                   // every rank's partial is complete and visible
    const long long vec_total = rows * N / 8;
    const long long per = (vec_total + c.world - 1) / c.world;
    const long long v0 = per * c.rank, v1 = v0 + per < vec_total ? v0 + per : vec_total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += 4 * stride) {
        uint4 v[4];   // four 16-byte reductions in flight per thread: the NVLink round trip is microseconds
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < v1) v[u] = multimem_ld_reduce_f16x8(src_mc + (i + u * stride) * 8);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u * stride < v1) {
                if (bias) v[u] = add_bias_f16x8(v[u], bias, (size_t)(((i + u * stride) * 8) % N));
                multimem_st_v4(dst_mc + (i + u * stride) * 8, v[u]);
            }
        }
    }
    __threadfence_system();
    peer_barrier(c, 1);                                   // every rank's slice has landed everywhere
}

// Second half of the fused prefill path: staging [world slots][rows_per_owner][N] fp16 in this rank's memory (filled by
// all ranks' GEMM epilogues) -> sum in fp32 (+ bias) -> rows [rank rows_per_owner, ...) of every rank's output.
// dst_mc: multicast address of the [M, N] result; dst_peers: the same buffer through plain peer pointers (used when
// the box has no multicast support: dst_mc == null).
struct PeerBcastDst { __half* peers[kPeerMaxWorld]; };

__global__ void __launch_bounds__(kPeerThreads)
peer_reduce_bcast_kernel(const PeerCtx c, const __half* staging, __half* dst_mc, const PeerBcastDst dst, const __half* bias,
                         int rows_per_owner, int N) {
    peer_barrier(c, 0);                                   // every rank's GEMM has stored its tiles into my staging slots
    const long long vec_rows = (long long)rows_per_owner * N / 8;
    const size_t slot_elems = (size_t)rows_per_owner * N;
    const size_t out0 = (size_t)c.rank * slot_elems;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < vec_rows; i0 += 2 * stride) {
        // two 16-byte columns per thread, all `world` slot loads of both in flight before the first add
        uint4 v[2][kPeerMaxWorld];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long i = i0 + u * stride;
#pragma unroll
            for (int s = 0; s < kPeerMaxWorld; ++s)
                if (s < c.world && i < vec_rows) v[u][s] = __ldcg(reinterpret_cast<const uint4*>(staging + s * slot_elems + i * 8));   // written by peers: L2, never a stale L1 line
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long i = i0 + u * stride;
            if (i >= vec_rows) break;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < kPeerMaxWorld; ++s) {
                if (s >= c.world) break;
                const __half2* h = reinterpret_cast<const __half2*>(&v[u][s]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
            if (bias) {
                const uint4 bv = *reinterpret_cast<const uint4*>(bias + (size_t)((i * 8) % N));
                const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(bh[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
            // Unicast peer stores (own copy: a local store).  A multimem.st would cut the bytes LEAVING this GPU to S/world,
            // but the switch also returns this rank's own copy, so every rank RECEIVES S bytes instead of (world-1)/world S
            // -- and the receive side is what bounds an all-gather (measured at world 2: 71 us multicast for 16 MiB).
            if (dst.peers[0]) {
#pragma unroll
                for (int r = 0; r < kPeerMaxWorld; ++r)
                    if (r < c.world) *reinterpret_cast<uint4*>(dst.peers[r] + out0 + i * 8) = o;
            } else {
                multimem_st_v4(dst_mc + out0 + i * 8, o);
            }
        }
    }
    __threadfence_system();
    peer_barrier(c, 1);
}

}  // namespace flatq
