// Instantiations of the lnL-only operand-resident walk (pair_walk.cuh): four categories, look-up tables of at most 8 rows - the headline shape.
#include "pair_walk.cuh"

namespace phb {

int launch_pair_k4n8(Ctx* c, int ppt, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift) {
#ifndef PHB_PAIR_PROBE_ONLY
    if (ppt == 4) return launch_pair_v<4, 8, 4>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
#endif
    (void)ppt;
    return launch_pair_v<4, 8, 2>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
}

}  // namespace phb
