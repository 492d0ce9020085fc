// TEST INFRASTRUCTURE: exposes flatquant_b200/csrc/grouped_sched.h (the tile schedule of the grouped GEMM, shared with the
// CUDA kernel) to tests/test_grouped_sched.py.  Built by g++ at test time; not part of the product library.
#include "../../flatquant_b200/csrc/grouped_sched.h"

extern "C" {
void shim_build_prefix(const int* off, int n_groups, int tn, int nb_w, int* tile_prefix) {
    flatq::grouped_build_prefix(off, n_groups, tn, nb_w, tile_prefix);
}
// out[4] = {group, m0, m_end, n0}
void shim_tile(const int* off, const int* tile_prefix, int n_groups, int tn, int block_w, int tile, int* out) {
    const flatq::GroupTile t = flatq::grouped_tile(off, tile_prefix, n_groups, tn, block_w, tile);
    out[0] = t.group; out[1] = t.m0; out[2] = t.m_end; out[3] = t.n0;
}
}
