// Launch dispatch for the several-chains-per-warp NUTS kernel (mmc_nuts_group.cuh), instantiated once per arithmetic
// policy (mmc_nuts_group_fast.cu / mmc_nuts_group_exact.cu).
#pragma once

#include "mmc_nuts_group.cuh"
#include "mmc_nuts_inst.cuh"

namespace mmc {

template <class Target, class A, class ST, int E, int G, bool kReplay>
int nuts_group_launch_one(const Target &tgt, NutsParams p, int sm_count, int64_t *grid_out, size_t *scratch_floats,
                          bool query_only, cudaStream_t stream) {
    using W = NutsGroup<Target, A, ST, E, G, kReplay>;
    auto kernel = nuts_group_kernel<Target, A, ST, E, G, kReplay>;
    const size_t smem = (size_t)kGrpWarps * ((size_t)W::kWarpFloats * sizeof(float) + W::kScalBytes);
    MMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGrpWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * sm_count;
    const int64_t per_cta = (int64_t)kGrpWarps * W::NG;
    const int64_t need = (p.chains + per_cta - 1) / per_cta;
    if (grid > need) grid = need;
    const int n_glob = p.max_depth > W::kL ? p.max_depth - W::kL : 0;
    *grid_out = grid;
    *scratch_floats = (size_t)grid * kGrpWarps * n_glob * 3 * W::kVec;
    if (query_only) return MMC_OK;
    kernel<<<(unsigned)grid, kGrpWarps * 32, smem, stream>>>(tgt, p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// lanes per chain the group kernel is compiled with for a target (0: not available -> warp kernel)
inline int nuts_group_lanes(const mmc_target_desc &t) {
    switch (t.kind) {
    case MMC_T_ROSENBROCK_ND:
        if (t.dim <= 4) return 4;
#ifdef MMC_NUTS_GROUP_TUNE_G4   // tuning builds: eight chains per warp at D = 100 (E = 26)
        if (t.dim > 64 && t.dim <= 104) return 4;
#endif
#ifdef MMC_NUTS_GROUP_TUNE_G16   // tuning builds: two chains per warp at D = 100 (E = 8)
        if (t.dim > 64 && t.dim <= 128) return 16;
#endif
        if (t.dim <= 104) return 8;
        if (t.dim <= 128) return 16;
        return 0;
    case MMC_T_STD_NORMAL:
        if (t.dim <= 4) return 4;
        if (t.dim <= 32) return 8;
        return 0;
    case MMC_T_ROSENBROCK_2D:
    case MMC_T_DIFF_GAUSSIAN2D:
        return 4;
    default:
        return 0;
    }
}

#define MMC_GROUP_LAUNCH(TGT, E_, G_, init) \
    return nuts_group_launch_one<TGT<A, E_, G_>, A, ST, E_, G_, kReplay>(init, p, L.sm_count, grid, scratch, query, s)

template <class A, class ST, bool kReplay>
int nuts_group_dispatch_target(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query,
                               cudaStream_t s) {
    const mmc_target_desc &t = L.target;
    switch (t.kind) {
    case MMC_T_ROSENBROCK_ND:
        if (t.dim <= 4) MMC_GROUP_LAUNCH(GRosenbrockND, 1, 4, {t.dim});
        // throughput policy above 4 dimensions: packed f32x2 kernels (E / 2 register pairs per lane)
#define MMC_GROUP_LAUNCH_PACKED(E_, G_) \
    return nuts_group_launch_one<GRosenbrockNDP<E_, G_>, A, ST, E_, G_, kReplay>({t.dim}, p, L.sm_count, grid, scratch, query, s)
        if constexpr (A::kContract) {
            if (t.dim <= 32) MMC_GROUP_LAUNCH_PACKED(4, 8);
            if (t.dim <= 64) MMC_GROUP_LAUNCH_PACKED(8, 8);
        }
        if (t.dim <= 32) MMC_GROUP_LAUNCH(GRosenbrockND, 4, 8, {t.dim});
        if (t.dim <= 64) MMC_GROUP_LAUNCH(GRosenbrockND, 8, 8, {t.dim});
#ifdef MMC_NUTS_GROUP_TUNE_G16   // tuning builds: two chains per warp at D = 100
        if constexpr (A::kContract) {
            if (t.dim <= 128) MMC_GROUP_LAUNCH_PACKED(8, 16);
        }
#endif
#ifdef MMC_NUTS_GROUP_TUNE_G4
        if constexpr (A::kContract) {
            if (t.dim <= 104) MMC_GROUP_LAUNCH_PACKED(26, 4);
        }
#endif
        if (t.dim <= 104) {
            if constexpr (A::kContract) MMC_GROUP_LAUNCH_PACKED(14, 8);
            else MMC_GROUP_LAUNCH(GRosenbrockND, 13, 8, {t.dim});
        }
        if (t.dim <= 128) {
            if constexpr (A::kContract) MMC_GROUP_LAUNCH_PACKED(8, 16);
            else MMC_GROUP_LAUNCH(GRosenbrockND, 8, 16, {t.dim});
        }
#undef MMC_GROUP_LAUNCH_PACKED
        break;
    case MMC_T_STD_NORMAL:
        if (t.dim <= 4) MMC_GROUP_LAUNCH(GStdNormal, 1, 4, {t.dim});
        if (t.dim <= 32) MMC_GROUP_LAUNCH(GStdNormal, 4, 8, {t.dim});
        break;
    case MMC_T_ROSENBROCK_2D: {
        GSmall<Rosenbrock2D<A>, 1, 4> w;
        w.t.a = (float)t.params[0];
        w.t.b = (float)t.params[1];
        return nuts_group_launch_one<GSmall<Rosenbrock2D<A>, 1, 4>, A, ST, 1, 4, kReplay>(w, p, L.sm_count, grid, scratch, query, s);
    }
    case MMC_T_DIFF_GAUSSIAN2D: {
        GSmall<DiffGaussian2D<A>, 1, 4> w;
        w.t = nuts_make_diff_gaussian<A>(t);
        return nuts_group_launch_one<GSmall<DiffGaussian2D<A>, 1, 4>, A, ST, 1, 4, kReplay>(w, p, L.sm_count, grid, scratch, query, s);
    }
    default: break;
    }
    set_error("NUTS group layout: target kind %d with dim %d is not compiled in", t.kind, t.dim);
    return MMC_ERR_UNSUPPORTED;
}
#undef MMC_GROUP_LAUNCH

template <class A>
int nuts_group_dispatch(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {
    if (L.scalar_f64)
        return L.replay ? nuts_group_dispatch_target<A, double, true>(L, p, grid, scratch, query, s)
                        : nuts_group_dispatch_target<A, double, false>(L, p, grid, scratch, query, s);
    return L.replay ? nuts_group_dispatch_target<A, float, true>(L, p, grid, scratch, query, s)
                    : nuts_group_dispatch_target<A, float, false>(L, p, grid, scratch, query, s);
}

int nuts_group_dispatch_fast(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s);
int nuts_group_dispatch_exact(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s);

}  // namespace mmc
