// Pre-order ("up") partials and per-edge likelihood derivatives.  Filled in after the pruning
// path is validated on hardware; until then the entry points report PHB_ERR_UNSUPPORTED.
#include "common.cuh"

namespace phb {

int launch_up_partials(Ctx* c) { return c->fail(PHB_ERR_UNSUPPORTED, "up partials: not built yet"); }

int launch_edge_derivatives(Ctx* c, int, const int32_t*, const double*, int, double*) {
    return c->fail(PHB_ERR_UNSUPPORTED, "edge derivatives: not built yet");
}

}  // namespace phb
