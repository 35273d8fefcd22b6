// Several-chains-per-warp NUTS kernel instantiations, reference arithmetic (no FMA contraction; -fmad=false).
#include "mmc_nuts_group_inst.cuh"
namespace mmc {
int nuts_group_dispatch_exact(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {
    return nuts_group_dispatch<Exact>(L, p, grid, scratch, query, s);
}
}  // namespace mmc
