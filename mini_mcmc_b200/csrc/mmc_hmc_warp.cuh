// K2w: HMC with one chain per warp, for dimensions the register-resident thread kernel is not instantiated for
// (D <= 512).  Same transition as hmc_run_kernel (src/hmc.rs:304-431); the D-vector is blocked over the lanes
// (lane l owns elements l*E .. l*E+E-1), dot products are warp butterflies, draws are stored as coalesced rows.
#pragma once

#include "mmc_hmc.cuh"
#include "mmc_nuts.cuh"   // warp-form targets (WRosenbrockND, WStdNormal, WSmall) and warp_sum helpers

namespace mmc {

template <class Target, class A, int E, bool kReplay>
__global__ void __launch_bounds__(128) hmc_warp_kernel(const Target tgt, const HmcParams p, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= p.chains) return;
    float x[E], pos[E], mom[E], g[E], gh[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const int i = lane * E + k;
        x[k] = i < D ? p.positions[c * D + i] : 0.0f;
    }
    const float eps = p.eps, eps_half = A::mul(eps, 0.5f);
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(c + p.chain_offset);
    unsigned int n_acc = 0;
    for (int64_t s = 0; s < steps; ++s) {
        float u;
        if (kReplay) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = lane * E + k;
                mom[k] = i < D ? __ldg(p.momenta + (s * p.chains + c) * D + i) : 0.0f;
            }
            u = __ldg(p.u + s * p.chains + c);
        } else {
            const uint32_t gstep = (uint32_t)(p.step_base + s);
#pragma unroll
            for (int k = 0; k < E; ++k) {   // sub j -> normals 4j..4j+3 (same contract as the thread kernel)
                const int i = lane * E + k;
                if ((k & 3) == 0 || E < 4) {
                    const uint4 w = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)(i >> 2)));
                    float n[4];
                    box_muller_f32(w.x, w.y, n[0], n[1]);
                    box_muller_f32(w.z, w.w, n[2], n[3]);
                    if (E >= 4) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (k + e < E) mom[k + e] = (i + e < D) ? n[e] : 0.0f;
                    } else {
                        mom[k] = i < D ? n[i & 3] : 0.0f;
                    }
                }
            }
            u = u24_half_open(philox_scalar_words(p.key, gchain, gstep).x);
        }
        float lp = tgt.logp_grad(x, g, lane);
        float ke = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            gh[k] = A::mul(g[k], eps_half);
            pos[k] = x[k];
            ke = A::mad(mom[k], mom[k], ke);
        }
        if (Target::kPartial) warp_sum2<A>(lp, ke); else ke = warp_sum<A>(ke);
        const float logp_cur = lp;
        const float h_cur = A::mad(ke, 0.5f, -logp_cur);
        float logp_prop_partial = 0.0f;
        for (int l = 0; l < p.n_leapfrog; ++l) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                mom[k] = A::add(mom[k], gh[k]);
                pos[k] = A::mad(mom[k], eps, pos[k]);
            }
            logp_prop_partial = tgt.logp_grad(pos, g, lane);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                gh[k] = A::mul(g[k], eps_half);
                mom[k] = A::add(mom[k], gh[k]);
            }
        }
        float ke2 = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) ke2 = A::mad(mom[k], mom[k], ke2);
        float logp_prop = logp_prop_partial;
        if (p.n_leapfrog == 0) { logp_prop = logp_cur; ke2 = warp_sum<A>(ke2); }
        else if (Target::kPartial) warp_sum2<A>(logp_prop, ke2);
        else ke2 = warp_sum<A>(ke2);
        const float h_prop = A::mad(ke2, 0.5f, -logp_prop);
        const float accept_logp = A::sub(h_cur, h_prop);
        const bool acc = accept_logp >= logf(u);
        if (acc) {
#pragma unroll
            for (int k = 0; k < E; ++k) x[k] = pos[k];
            ++n_acc;
        }
        if (p.trace && lane == 0)
            reinterpret_cast<float4 *>(p.trace)[s * p.chains + c] = make_float4(logp_cur, logp_prop, accept_logp, acc ? 1.0f : 0.0f);
        if (s >= p.n_discard && p.out) {
            float *o = p.out + (c * p.out_pitch + (s - p.n_discard)) * D;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = lane * E + k;
                if (i < D) o[i] = x[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const int i = lane * E + k;
        if (i < D) p.positions[c * D + i] = x[k];
    }
    if (lane == 0 && n_acc) atomicAdd(p.accept_count, (unsigned long long)n_acc);
}

template <class Target, class A, int E>
int launch_hmc_warp(const Target &tgt, const HmcParams &p, int D, bool replay, cudaStream_t stream) {
    const int block = 128;
    const int64_t grid = (p.chains * 32 + block - 1) / block;
    if (replay) hmc_warp_kernel<Target, A, E, true><<<(unsigned)grid, block, 0, stream>>>(tgt, p, D);
    else hmc_warp_kernel<Target, A, E, false><<<(unsigned)grid, block, 0, stream>>>(tgt, p, D);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// native-mode draws for any D: momenta [steps, chains, D], u [steps, chains]
__global__ void hmc_export_tape_any_kernel(uint2 key, int64_t chains, int D, int64_t chain_offset, int64_t step_base,
                                           int64_t steps, float *momenta, float *u) {
    const int64_t groups = (D + 3) / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= steps * chains * groups) return;
    const int64_t j = idx % groups, sc = idx / groups, c = sc % chains, s = sc / chains;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    const uint32_t gstep = (uint32_t)(step_base + s);
    const uint4 w = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)j));
    float n[4];
    box_muller_f32(w.x, w.y, n[0], n[1]);
    box_muller_f32(w.z, w.w, n[2], n[3]);
    for (int e = 0; e < 4; ++e)
        if (4 * j + e < D) momenta[sc * D + 4 * j + e] = n[e];
    if (j == 0) u[sc] = u24_half_open(philox_scalar_words(key, gchain, gstep).x);
}

}  // namespace mmc
