// NUTS kernel instantiations, exact arithmetic (no FMA contraction; replay parity).
#include "mmc_nuts_inst.cuh"
namespace mmc {
int nuts_dispatch_exact(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {
    return nuts_dispatch<Exact>(L, p, grid, scratch, query, s);
}
}  // namespace mmc
