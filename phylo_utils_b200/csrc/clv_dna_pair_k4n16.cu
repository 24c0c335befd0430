// Instantiations of the lnL-only operand-resident walk (pair_walk.cuh): four categories, look-up tables of up to 16 rows (partial ambiguity codes).
#include "pair_walk.cuh"

namespace phb {

int launch_pair_k4n16(Ctx* c, int ppt, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift) {
    if (ppt == 4) return launch_pair_v<4, 16, 4>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
    return launch_pair_v<4, 16, 2>(c, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
}

}  // namespace phb
