// Instantiations of the lnL-only operand-resident walk (pair_walk.cuh): eight categories.
#include "pair_walk.cuh"

namespace phb {

int launch_pair_k8(Ctx* c, int nc, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift) {
    return nc == 8 ? launch_pair_v<8, 8, 2>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift) : launch_pair_v<8, 16, 2>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
}

}  // namespace phb
