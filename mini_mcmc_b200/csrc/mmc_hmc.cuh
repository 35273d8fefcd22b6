// K2: fused HMC.  One thread owns one chain for the WHOLE run: position, momentum and the cached
// half-step gradient stay in registers across all L leapfrog steps and all transitions; draws are
// streamed to HBM in the [chains, n_collect, dim] layout the reference returns.
//
// Reproduces HMC::step + leapfrog + run, src/hmc.rs:137-158,304-431:
//   g_half = grad(x) * (eps*0.5);  L x { p += g_half; x += p*eps; g_half = grad(x)*(eps*0.5); p += g_half }
//   H = -logp + 0.5 sum p^2;  accept iff H_cur - H_prop >= ln(u)   (non-strict, NaN rejects)
// The reference's extra logp evaluation at the end of the trajectory (:429) is the value the last
// gradient evaluation already produced, so it is reused rather than recomputed.
#pragma once

#include "mmc_common.cuh"

namespace mmc {

struct HmcParams {
    float *positions;        // [chains, D] in/out
    float *out;              // [chains, n_collect, D] or nullptr
    const float *momenta;    // replay [steps, chains, D] or nullptr
    const float *u;          // replay [steps, chains]
    float *trace;            // optional [steps, chains, 4]
    unsigned long long *accept_count;  // device counter (accepted transitions)
    int64_t chains;
    int64_t chain_offset;
    int64_t step_base;
    int64_t n_collect, n_discard;
    int64_t out_pitch;       // draws per chain row of `out` (>= n_collect)
    float eps;
    int n_leapfrog;
    uint2 key;
};

template <class Target, class A, bool kReplay>
__global__ void __launch_bounds__(128) hmc_run_kernel(const Target tgt, const HmcParams p) {
    constexpr int D = Target::kDim;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int n_acc = 0;
    if (c < p.chains) {
        float x[D], pos[D], mom[D], g[D], gh[D];
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = p.positions[c * D + i];
        const float eps = p.eps;
        const float eps_half = A::mul(eps, 0.5f);
        const int64_t steps = p.n_collect + p.n_discard;
        const uint64_t gchain = (uint64_t)(c + p.chain_offset);
        for (int64_t s = 0; s < steps; ++s) {
            float u;
            if (kReplay) {
#pragma unroll
                for (int i = 0; i < D; ++i) mom[i] = __ldg(p.momenta + (s * p.chains + c) * D + i);
                u = __ldg(p.u + s * p.chains + c);
            } else {
                const uint32_t gstep = (uint32_t)(p.step_base + s);
                philox_normals_f32<D>(p.key, gchain, gstep, mom);
                u = u24_half_open(philox_scalar_words(p.key, gchain, gstep).x);
            }
            const float logp_cur = tgt.logp_grad(x, g);
            float ke = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                gh[i] = A::mul(g[i], eps_half);
                pos[i] = x[i];
                ke = A::mad(mom[i], mom[i], ke);
            }
            const float h_cur = A::mad(ke, 0.5f, -logp_cur);
            float logp_prop = logp_cur;
            for (int l = 0; l < p.n_leapfrog; ++l) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    mom[i] = A::add(mom[i], gh[i]);
                    pos[i] = A::mad(mom[i], eps, pos[i]);
                }
                logp_prop = tgt.logp_grad(pos, g);
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    gh[i] = A::mul(g[i], eps_half);
                    mom[i] = A::add(mom[i], gh[i]);
                }
            }
            float ke2 = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) ke2 = A::mad(mom[i], mom[i], ke2);
            const float h_prop = A::mad(ke2, 0.5f, -logp_prop);
            const float accept_logp = A::sub(h_cur, h_prop);
            const bool acc = accept_logp >= logf(u);
            if (acc) {
#pragma unroll
                for (int i = 0; i < D; ++i) x[i] = pos[i];
                ++n_acc;
            }
            if (p.trace) {
                float4 t = make_float4(logp_cur, logp_prop, accept_logp, acc ? 1.0f : 0.0f);
                reinterpret_cast<float4 *>(p.trace)[s * p.chains + c] = t;
            }
            if (s >= p.n_discard && p.out) {
                float *o = p.out + (c * p.out_pitch + (s - p.n_discard)) * D;
#pragma unroll
                for (int i = 0; i < D; ++i) o[i] = x[i];
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) p.positions[c * D + i] = x[i];
    }
    // one atomic per warp
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(p.accept_count, (unsigned long long)n_acc);
}

// native-mode draws for (chain, step) exactly as hmc_run_kernel consumes them
template <int D>
__global__ void hmc_export_tape_kernel(uint2 key, int64_t chains, int64_t chain_offset, int64_t step_base,
                                       int64_t steps, float *momenta, float *u) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chains * steps) return;
    const int64_t s = idx / chains, c = idx % chains;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    const uint32_t gstep = (uint32_t)(step_base + s);
    float m[D];
    philox_normals_f32<D>(key, gchain, gstep, m);
#pragma unroll
    for (int i = 0; i < D; ++i) momenta[idx * D + i] = m[i];
    u[idx] = u24_half_open(philox_scalar_words(key, gchain, gstep).x);
}

template <class Target, class A>
int launch_hmc(const Target &tgt, const HmcParams &p, bool replay, cudaStream_t stream) {
    const int block = 128;
    const int64_t grid = (p.chains + block - 1) / block;
    if (replay)
        hmc_run_kernel<Target, A, true><<<(unsigned)grid, block, 0, stream>>>(tgt, p);
    else
        hmc_run_kernel<Target, A, false><<<(unsigned)grid, block, 0, stream>>>(tgt, p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace mmc
