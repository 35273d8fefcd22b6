// K3 tensor-core path (tcgen05 + TMA): placeholder until the kernel lands.
#include "mmc_dense.cuh"

namespace mmc {
struct DenseState;
int dense_tc_prepare(DenseState *) {
    set_error("dense Gaussian: the tcgen05 path is not built yet");
    return MMC_ERR_UNSUPPORTED;
}
int dense_gemm_tc(DenseState *, int, int64_t, int, float, int, cudaStream_t) { return MMC_ERR_UNSUPPORTED; }
void dense_tc_destroy(DenseState *) {}
}  // namespace mmc
