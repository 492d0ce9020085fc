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
//     bias once, rounds to fp16 and writes the result to EVERY rank with one multimem.st per 16 bytes (the switch
//     replicates); barrier.  NVLink bytes per rank and linear: (tp-1)/tp S out during the GEMM, S/tp out +
//     (tp-1)/tp S in during the broadcast (S = M N 2 bytes), against 2 (tp-1)/tp S each way for a ring all-reduce.
// Any other shape (decode sizes, ragged M) -- peer_allreduce_nvls_kernel on the GEMM's local output:
//   barrier; rank r pulls ITS 1/tp of the elements already summed by the switch (multimem.ld_reduce.add, fp32
//   accumulation), adds the bias, multimem.st to every rank; barrier.
//
// Barriers: one 32-bit flag per (CTA, peer) in each rank's signal pad; a CTA of rank r raises flag [blk][r] on every
// peer (release, system scope) and waits for [blk][p] of all peers p on its own pad (acquire), then clears them.
#pragma once
#include "ptx.cuh"

namespace flatq {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerThreads = 512;

struct PeerCtx {
    int rank, world;
    uint32_t* signal_pads[kPeerMaxWorld];   // each rank's signal pad (symmetric memory), >= 2 * grid * world words
};

__device__ __forceinline__ void peer_put_flag(uint32_t* addr) {
    uint32_t old;
    do {
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 0);
}
__device__ __forceinline__ void peer_wait_flag(uint32_t* addr) {
    uint32_t old;
    do {
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 1);
}
// all CTAs with the same blockIdx on every rank meet here; `phase` selects one of two flag sets so that back-to-back
// barriers of one kernel cannot overtake each other
__device__ __forceinline__ void peer_barrier(const PeerCtx& c, int phase) {
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
        const size_t slot = ((size_t)phase * gridDim.x + blockIdx.x) * c.world;
        peer_put_flag(c.signal_pads[threadIdx.x] + slot + c.rank);
        peer_wait_flag(c.signal_pads[c.rank] + slot + threadIdx.x);
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
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec_rows; i += (long long)gridDim.x * blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < c.world; ++s) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(staging + s * slot_elems + i * 8));   // written by peers: L2, never a stale L1 line
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
        if (bias) {
            const uint4 b = *reinterpret_cast<const uint4*>(bias + (size_t)((i * 8) % N));
            const __half2* bh = reinterpret_cast<const __half2*>(&b);
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
        if (dst_mc) {
            multimem_st_v4(dst_mc + out0 + i * 8, o);
        } else {
            for (int r = 0; r < c.world; ++r) *reinterpret_cast<uint4*>(dst.peers[r] + out0 + i * 8) = o;
        }
    }
    __threadfence_system();
    peer_barrier(c, 1);
}

}  // namespace flatq
