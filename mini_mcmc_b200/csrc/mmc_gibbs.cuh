// Shared pieces of the Gibbs sampler (csrc/mmc_gibbs.cu) that custom conditionals instantiate
// (include/minimcmc_target.cuh, MMC_REGISTER_GIBBS_CONDITIONAL).
#pragma once

#include "mmc_common.cuh"

namespace mmc {

struct GibbsParams {
    double *state;          // [chains, D] in/out
    double *out;            // [chains, out_pitch, D]
    const double *normals;  // replay [chains, steps]   (built-in mixture only)
    const double *unifs;    // replay [chains, steps]
    double *trace;          // optional [chains, steps, 2]: the (z-score, uniform) each sweep consumed
    int64_t chains, chain_offset, step_base, n_collect, n_discard, out_pitch;
    int32_t kind, D;
    double p[8];
    uint2 key;
};

// Draws of one coordinate update: Philox counter (global chain, step, sub = coordinate + 256 * draw index), so a
// conditional may consume any number of uniforms / normals per coordinate and stays independent of the sharding.
struct GibbsRng {
    uint2 key;
    uint32_t gc_lo, gc_hi, step, coord, draw;
    __device__ __forceinline__ uint4 next() {
        const uint4 w = philox4x32_10(key, make_uint4(gc_lo, gc_hi, step, coord + 256u * draw));
        ++draw;
        return w;
    }
    __device__ __forceinline__ double uniform() {   // [0, 1), 53 bits (rand's StandardUniform for f64)
        const uint4 w = next();
        return u53_half_open(w.x, w.y);
    }
    __device__ __forceinline__ double normal() {    // standard normal z-score
        double n0, n1;
        box_muller_f64(next(), n0, n1);
        return n0;
    }
};

// GibbsMarkovChain::step, src/gibbs.rs:122-126, for a user conditional: one thread per chain, state in registers.
template <class Cond>
__global__ void __launch_bounds__(128) gibbs_generic_kernel(const Cond cond, const GibbsParams p) {
    constexpr int D = Cond::kDim;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = p.state[c * D + i];
    const uint64_t gc = (uint64_t)(c + p.chain_offset);
    GibbsRng rng{p.key, (uint32_t)gc, (uint32_t)(gc >> 32), 0u, 0u, 0u};
    const int64_t steps = p.n_collect + p.n_discard;
    for (int64_t s = 0; s < steps; ++s) {
        rng.step = (uint32_t)(p.step_base + s);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            rng.coord = (uint32_t)i;
            rng.draw = 0u;
            x[i] = cond.sample(i, x, rng);
        }
        if (s >= p.n_discard) {
            double *o = p.out + (c * p.out_pitch + (s - p.n_discard)) * D;
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = x[i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) p.state[c * D + i] = x[i];
}

template <class Cond>
int launch_gibbs_generic(const Cond &cond, const GibbsParams &p, cudaStream_t stream) {
    const unsigned grid = (unsigned)((p.chains + 127) / 128);
    gibbs_generic_kernel<Cond><<<grid, 128, 0, stream>>>(cond, p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace mmc
