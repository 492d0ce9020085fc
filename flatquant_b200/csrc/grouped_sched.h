// Tile schedule of the grouped (mixture-of-experts) W4A4 GEMM: one launch covers E independent GEMMs that share
// N and K -- the routed-expert FlatLinears of the DeepSeek block, which the reference runs as a Python loop over
// experts (deepseek_v3/model.py:677-683 under flatquant/model_tools/deepseekv3_utils.py:413-470).
//
//   group e owns token rows [off[e], off[e + 1]) of the sorted activation / output matrices ([M_total, K] / [M_total, N])
//   and weight rows [e * N, (e + 1) * N) of the stacked packed weight ([E * N, K / 2]).
//
// Output tiles are numbered group by group; inside a group token blocks run fastest (the order of the plain GEMM).
// `tile_prefix[e]` = number of tiles of the groups before e, built on the device from `off` (so the host never reads
// the routing result), and looked up by binary search by every warp role that needs coordinates.
//
// Plain C++ without CUDA intrinsics: compiled into the kernels (csrc/gemm_w4a8.cuh) and, by g++, into the CPU test
// tests/test_grouped_sched.py, which checks it against a Python model on random routings.
#pragma once

#if defined(__CUDACC__)
#define FLATQ_HD __host__ __device__ __forceinline__
#else
#define FLATQ_HD inline
#endif

namespace flatq {

struct GroupTile {
    int group;    // expert index
    int m0;       // first token row of the tile (row of the sorted activation matrix)
    int m_end;    // first row that no longer belongs to the group (rows >= m_end of the tile are masked)
    int n0;       // first output feature of the tile, inside the group
};

// tile_prefix[0..E]: tile_prefix[e + 1] - tile_prefix[e] = nb_w * ceil(rows_e / tn).  Serial: E <= a few hundred.
FLATQ_HD void grouped_build_prefix(const int* off, int n_groups, int tn, int nb_w, int* tile_prefix) {
    int acc = 0;
    tile_prefix[0] = 0;
    for (int e = 0; e < n_groups; ++e) {
        int rows = off[e + 1] - off[e];
        if (rows < 0) rows = 0;  // a malformed routing must not produce a negative tile count
        acc += nb_w * ((rows + tn - 1) / tn);
        tile_prefix[e + 1] = acc;
    }
}

// largest e with tile_prefix[e] <= tile; requires 0 <= tile < tile_prefix[n_groups] (so the group found is not empty)
FLATQ_HD int grouped_find(const int* tile_prefix, int n_groups, int tile) {
    int lo = 0, hi = n_groups;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

FLATQ_HD GroupTile grouped_tile(const int* off, const int* tile_prefix, int n_groups, int tn, int block_w, int tile) {
    GroupTile t;
    t.group = grouped_find(tile_prefix, n_groups, tile);
    const int local = tile - tile_prefix[t.group];
    const int row0 = off[t.group];
    t.m_end = off[t.group + 1];
    const int nb_m = (t.m_end - row0 + tn - 1) / tn;  // >= 1: the group has tiles
    t.m0 = row0 + (local % nb_m) * tn;
    t.n0 = (local / nb_m) * block_w;
    return t;
}

}  // namespace flatq
