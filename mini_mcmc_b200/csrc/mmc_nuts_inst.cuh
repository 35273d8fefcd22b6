// Launch dispatch for the NUTS kernel, instantiated once per arithmetic policy (mmc_nuts_fast.cu /
// mmc_nuts_exact.cu) to keep compile times parallel.
#pragma once

#include <cmath>

#include "mmc_nuts.cuh"

namespace mmc {

struct NutsLaunch {
    mmc_target_desc target;
    bool scalar_f64;
    bool replay;
    int sm_count;
};

// mmc_nuts_set_layout values (include/minimcmc.h)
constexpr int kNutsLayoutAuto = 0, kNutsLayoutWarp = 32;

template <class A>
DiffGaussian2D<A> nuts_make_diff_gaussian(const mmc_target_desc &t) {
    // DiffableGaussian2D::new, src/distributions.rs:227-251 (T = f64), then cast to the backend float
    const double c00 = t.params[2], c01 = t.params[3], c10 = t.params[4], c11 = t.params[5];
    const double det = c00 * c11 - c01 * c10;
    const double inv_det = 1.0 / det;
    DiffGaussian2D<A> g;
    g.m0 = (float)t.params[0];
    g.m1 = (float)t.params[1];
    g.p00 = (float)(c11 * inv_det);
    g.p01 = (float)(-c01 * inv_det);
    g.p10 = (float)(-c10 * inv_det);
    g.p11 = (float)(c00 * inv_det);
    const double two = 2.0;
    g.norm_const = (float)(-(two * std::log(two * M_PI) + std::log(det)) / two);
    return g;
}

// grid size (persistent warps) and scratch floats needed for a configuration
template <class Target, class A, class ST, int E, bool kReplay>
int nuts_launch_one(const Target &tgt, NutsParams p, int sm_count, int64_t *grid_out, size_t *scratch_floats,
                    bool query_only, cudaStream_t stream) {
    auto kernel = nuts_run_kernel<Target, A, ST, E, kReplay>;
    constexpr int V = 32 * E;
    const size_t smem = (size_t)kNutsWarps * kNutsSmemLevels * 3 * V * sizeof(float) + (size_t)kNutsWarps * 256;
    int per_sm = 0;
    MMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kNutsWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * sm_count;
    const int64_t need = (p.chains + kNutsWarps - 1) / kNutsWarps;
    if (grid > need) grid = need;
    const int n_glob = p.max_depth > kNutsSmemLevels ? p.max_depth - kNutsSmemLevels : 0;
    *grid_out = grid;
    *scratch_floats = (size_t)grid * kNutsWarps * n_glob * 3 * V;
    if (query_only) return MMC_OK;
    kernel<<<(unsigned)grid, kNutsWarps * 32, smem, stream>>>(tgt, p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

template <class A, class ST, bool kReplay>
int nuts_dispatch_target(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query,
                         cudaStream_t s) {
    const mmc_target_desc &t = L.target;
    switch (t.kind) {
    case MMC_T_ROSENBROCK_ND:
        if (t.dim <= 32) return nuts_launch_one<WRosenbrockND<A, 1>, A, ST, 1, kReplay>({t.dim}, p, L.sm_count, grid, scratch, query, s);
        if (t.dim <= 128) return nuts_launch_one<WRosenbrockND<A, 4>, A, ST, 4, kReplay>({t.dim}, p, L.sm_count, grid, scratch, query, s);
        break;
    case MMC_T_STD_NORMAL:
        if (t.dim <= 32) return nuts_launch_one<WStdNormal<A, 1>, A, ST, 1, kReplay>({t.dim}, p, L.sm_count, grid, scratch, query, s);
        if (t.dim <= 128) return nuts_launch_one<WStdNormal<A, 4>, A, ST, 4, kReplay>({t.dim}, p, L.sm_count, grid, scratch, query, s);
        break;
    case MMC_T_ROSENBROCK_2D: {
        WSmall<Rosenbrock2D<A>, 1> w;
        w.t.a = (float)t.params[0];
        w.t.b = (float)t.params[1];
        return nuts_launch_one<WSmall<Rosenbrock2D<A>, 1>, A, ST, 1, kReplay>(w, p, L.sm_count, grid, scratch, query, s);
    }
    case MMC_T_DIFF_GAUSSIAN2D: {
        WSmall<DiffGaussian2D<A>, 1> w;
        w.t = nuts_make_diff_gaussian<A>(t);
        return nuts_launch_one<WSmall<DiffGaussian2D<A>, 1>, A, ST, 1, kReplay>(w, p, L.sm_count, grid, scratch, query, s);
    }
    default: break;
    }
    set_error("NUTS: target kind %d with dim %d is not compiled in (dim <= 128)", t.kind, t.dim);
    return MMC_ERR_UNSUPPORTED;
}

template <class A>
int nuts_dispatch(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {
    if (L.scalar_f64)
        return L.replay ? nuts_dispatch_target<A, double, true>(L, p, grid, scratch, query, s)
                        : nuts_dispatch_target<A, double, false>(L, p, grid, scratch, query, s);
    return L.replay ? nuts_dispatch_target<A, float, true>(L, p, grid, scratch, query, s)
                    : nuts_dispatch_target<A, float, false>(L, p, grid, scratch, query, s);
}

int nuts_dispatch_fast(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s);
int nuts_dispatch_exact(const NutsLaunch &L, const NutsParams &p, int64_t *grid, size_t *scratch, bool query, cudaStream_t s);

}  // namespace mmc
