// Several-chains-per-warp NUTS kernel instantiations, throughput arithmetic (FMA contraction allowed).
#include "mmc_nuts_group_inst.cuh"
namespace mmc {
int nuts_group_dispatch_fast(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {
    return nuts_group_dispatch<Fast>(L, p, grid, scratch, query, s);
}
}  // namespace mmc
