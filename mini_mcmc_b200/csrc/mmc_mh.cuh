// K1: batched Metropolis-Hastings, one chain per thread for the whole run.
// Reproduces MHMarkovChain::step (src/metropolis_hastings.rs:303-315) inside run_chain
// (src/core.rs:55-73): x' = proposal.sample(x); r = (lp' + q_b) - (lp + q_f); accept iff r > ln(u).
//
// This translation unit is compiled with -fmad=false: the f64 expressions below must round exactly
// like the CPU arithmetic (Rust does not contract a*b+c).
#pragma once

#include "mmc_common.cuh"

namespace mmc {

// ------------------------------------------------------------------ continuous state (f64)
struct MhContParams {
    double *state;         // [chains, D] in/out
    double *out;           // [chains, n_collect, D]
    const double *noise;   // replay [chains, steps, D]
    const double *u;       // replay [chains, steps]
    double *trace;         // optional [chains, steps, 4]
    int64_t chains, chain_offset, step_base, n_collect, n_discard;
    uint2 key;
    int32_t target_kind;   // MMC_T_GAUSSIAN2D | MMC_T_ISO_GAUSSIAN
    double tp[6];          // Gaussian2D: mean0, mean1, a, b, c, d ; Iso: std
    double prop_std;
    double prop_norm_term; // -D * 0.5 * ln(var * pi * std * std), evaluated by the host libm
};

// Gaussian2D::unnorm_logp, src/distributions.rs:193-205 (inverse re-derived per call in the reference;
// the quotients are loop-invariant so they are hoisted, the values are identical).
struct Gauss2DInv {
    double m0, m1, i00, i01, i10, i11;
    __device__ __forceinline__ explicit Gauss2DInv(const double *tp) {
        const double a = tp[2], b = tp[3], c = tp[4], d = tp[5];
        const double det = a * d - b * c;
        m0 = tp[0]; m1 = tp[1];
        i00 = d / det; i01 = -b / det; i10 = -c / det; i11 = a / det;
    }
    __device__ __forceinline__ double logp(const double *x) const {
        const double d0 = x[0] - m0, d1 = x[1] - m1;
        const double r0 = d0 * i00 + d1 * i10;
        const double r1 = d0 * i01 + d1 * i11;
        return -0.5 * (r0 * d0 + r1 * d1);
    }
};

template <int D>
__device__ __forceinline__ double iso_target_logp(double std, const double (&x)[D]) {
    // IsotropicGaussian as Target, src/distributions.rs:394-402
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) sum = sum + x[i] * x[i];
    return -0.5 * sum / (std * std);
}

template <int D>
__device__ __forceinline__ double iso_proposal_logp(double std, double norm_term, const double (&from)[D],
                                                    const double (&to)[D]) {
    // IsotropicGaussian::logp, src/distributions.rs:374-386
    double lp = 0.0;
    const double var = std * std;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const double diff = to[i] - from[i];
        lp += -(diff * diff) / (2.0 * var);
    }
    lp += norm_term;
    return lp;
}

template <int D, bool kReplay>
__global__ void __launch_bounds__(128) mh_cont_kernel(const MhContParams p) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    double x[D], prop[D], z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = p.state[c * D + i];
    const Gauss2DInv g2(p.tp);
    const bool is_g2 = p.target_kind == MMC_T_GAUSSIAN2D;
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(c + p.chain_offset);
    for (int64_t s = 0; s < steps; ++s) {
        double u;
        if (kReplay) {
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] = p.noise[(c * steps + s) * D + i];
            u = p.u[c * steps + s];
        } else {
            const uint32_t gstep = (uint32_t)(p.step_base + s);
            const uint4 w0 = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, 0u));
            u = u53_half_open(w0.x, w0.y);
#pragma unroll
            for (int j = 0; j < (D + 1) / 2; ++j) {
                const uint4 w = philox4x32_10(
                    p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)(1 + j)));
                double n0, n1;
                box_muller_f64(w, n0, n1);
                z[2 * j] = n0;
                if (2 * j + 1 < D) z[2 * j + 1] = n1;
            }
        }
        // Normal(0, std).sample = 0 + std*z, then `x + *eps` (src/distributions.rs:364-372)
#pragma unroll
        for (int i = 0; i < D; ++i) prop[i] = (0.0 + p.prop_std * z[i]) + x[i];
        double cur_lp, prop_lp;
        if (is_g2) {
            cur_lp = g2.logp(x);
            prop_lp = g2.logp(prop);
        } else {
            cur_lp = iso_target_logp<D>(p.tp[0], x);
            prop_lp = iso_target_logp<D>(p.tp[0], prop);
        }
        const double qf = iso_proposal_logp<D>(p.prop_std, p.prop_norm_term, x, prop);
        const double qb = iso_proposal_logp<D>(p.prop_std, p.prop_norm_term, prop, x);
        const double r = (prop_lp + qb) - (cur_lp + qf);
        const bool acc = r > log(u);
        if (acc) {
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] = prop[i];
        }
        if (p.trace) {
            double *t = p.trace + (c * steps + s) * 4;
            t[0] = cur_lp; t[1] = prop_lp; t[2] = r; t[3] = acc ? 1.0 : 0.0;
        }
        if (s >= p.n_discard && p.out) {
            double *o = p.out + (c * p.n_collect + (s - p.n_discard)) * D;
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = x[i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) p.state[c * D + i] = x[i];
}

// native-mode draws exactly as mh_cont_kernel consumes them: noise [chains, steps, D], u [chains, steps]
template <int D>
__global__ void mh_cont_export_tape_kernel(uint2 key, int64_t chains, int64_t chain_offset, int64_t step_base,
                                           int64_t steps, double *noise, double *u) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chains * steps) return;
    const int64_t c = idx / steps, s = idx % steps;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    const uint32_t gstep = (uint32_t)(step_base + s);
    const uint4 w0 = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, 0u));
    u[idx] = u53_half_open(w0.x, w0.y);
    for (int j = 0; j < (D + 1) / 2; ++j) {
        const uint4 w = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)(1 + j)));
        double n0, n1;
        box_muller_f64(w, n0, n1);
        noise[idx * D + 2 * j] = n0;
        if (2 * j + 1 < D) noise[idx * D + 2 * j + 1] = n1;
    }
}

// ------------------------------------------------------------------ Poisson / integer state (config C2)
// PoissonTarget + NonnegativeProposal, examples/poisson_mh.rs:10-89.
// Draw write-out is the HBM-bound part (8 B per collected transition): each warp stages T steps of its
// 32 chains in shared memory and emits them as contiguous 256 B row segments (full 32 B sectors)
// instead of 32 strided 8 B stores per step.
struct MhPoissonParams {
    uint64_t *state;        // [chains] in/out
    uint64_t *out;          // [chains, n_collect]
    const uint8_t *flip;    // replay [chains, steps]
    const double *u;        // replay [chains, steps]
    const double *lnfact;   // [table_len]  sum_{i<=k} ln i, built by the host libm in the reference's order
    const uint64_t *thr_up; // [table_len]  accept k -> k+1 iff u53 < thr_up[k]
    const uint64_t *thr_dn; // [table_len]  accept k -> k-1 iff u53 < thr_dn[k]
    int32_t table_len;
    double lambda, ln_lambda, ln_half;
    int64_t chains, chain_offset, step_base, n_collect, n_discard;
    uint2 key;
    int32_t *error_flag;    // set to 1 when a chain leaves the table range
};

constexpr int kPoisTile = 64;              // steps staged per write-out
constexpr int kPoisPitch = kPoisTile + 2;  // halfwords; 33 words -> conflict-free rows
constexpr int kPoisWarps = 8;

template <bool kReplay, bool kThreshold>
__global__ void __launch_bounds__(kPoisWarps * 32) mh_poisson_kernel(const MhPoissonParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [table_len] u64 x2 (threshold) or [table_len] f64 (log mode), then the warp tiles
    uint64_t *s_up = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_dn = s_up + p.table_len;
    const double *s_lnf = reinterpret_cast<const double *>(smem_raw);
    uint16_t *tiles = reinterpret_cast<uint16_t *>(smem_raw + (size_t)p.table_len * 16);
    for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) {
        if (kThreshold) {
            s_up[i] = p.thr_up[i];
            s_dn[i] = p.thr_dn[i];
        } else {
            reinterpret_cast<double *>(smem_raw)[i] = p.lnfact[i];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t *tile = tiles + warp * 32 * kPoisPitch;
    const int64_t chain0 = ((int64_t)blockIdx.x * kPoisWarps + warp) * 32;
    if (chain0 >= p.chains) return;
    const int64_t c = chain0 + lane;
    const bool active = c < p.chains;
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(c + p.chain_offset);
    const uint32_t kmax = (uint32_t)p.table_len - 1;
    uint32_t x = 0;
    bool overflow = false;
    if (active) {
        const uint64_t x0 = p.state[c];
        if (x0 >= kmax) { overflow = true; x = 0; } else x = (uint32_t)x0;
    }
    uint4 w = make_uint4(0, 0, 0, 0);

    auto transition = [&](int64_t s) {
        uint32_t flip;
        uint64_t u53 = 0;
        double u = 0.0;
        if (kReplay) {
            flip = active ? p.flip[c * steps + s] : 0;
            u = active ? p.u[c * steps + s] : 0.5;
            if (kThreshold) u53 = (uint64_t)(u * 9007199254740992.0);
        } else {
            const uint64_t gs = (uint64_t)(p.step_base + s);
            if ((gs & 1) == 0 || s == 0)
                w = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), (uint32_t)(gs >> 1), 0u));
            const uint32_t lo = (gs & 1) ? w.z : w.x, hi = (gs & 1) ? w.w : w.y;
            const uint64_t bits = (uint64_t)lo | ((uint64_t)hi << 32);
            flip = (uint32_t)(bits & 1);
            u53 = bits >> 11;
            if (!kThreshold) u = (double)u53 * (1.0 / 9007199254740992.0);
        }
        // NonnegativeProposal::sample, examples/poisson_mh.rs:34-47
        const uint32_t y = (x == 0) ? 1u : (flip ? x + 1 : x - 1);
        if (y >= kmax) { overflow = true; return; }
        bool acc;
        if (kThreshold) {
            const uint64_t thr = (y > x) ? s_up[x] : s_dn[x];
            acc = u53 < thr;
        } else {
            // PoissonTarget::unnorm_logp, examples/poisson_mh.rs:19-25: -lambda + k ln(lambda) - ln k!
            const double cur_lp = (-p.lambda + (double)x * p.ln_lambda) - s_lnf[x];
            const double prop_lp = (-p.lambda + (double)y * p.ln_lambda) - s_lnf[y];
            // NonnegativeProposal::logp, examples/poisson_mh.rs:53-71 (y is always x +- 1 here)
            const double qf = (x == 0) ? 0.0 : p.ln_half;
            const double qb = (y == 0) ? ((x == 1) ? 0.0 : -INFINITY) : p.ln_half;
            const double r = (prop_lp + qb) - (cur_lp + qf);
            acc = r > log(u);
        }
        if (acc) x = y;
    };

    for (int64_t s = 0; s < p.n_discard; ++s) transition(s);

    for (int64_t t0 = 0; t0 < p.n_collect; t0 += kPoisTile) {
        const int nt = (int)((p.n_collect - t0 < kPoisTile) ? (p.n_collect - t0) : kPoisTile);
        for (int t = 0; t < nt; ++t) {
            transition(p.n_discard + t0 + t);
            tile[lane * kPoisPitch + t] = (uint16_t)x;
        }
        __syncwarp();
        const int nrows = (int)((p.chains - chain0 < 32) ? (p.chains - chain0) : 32);
        for (int r = 0; r < nrows; ++r) {
            uint64_t *row = p.out + (chain0 + r) * p.n_collect + t0;
            const uint16_t *trow = tile + r * kPoisPitch;
            if (lane < nt) __stcs(reinterpret_cast<unsigned long long *>(row + lane), (unsigned long long)trow[lane]);
            if (lane + 32 < nt)
                __stcs(reinterpret_cast<unsigned long long *>(row + lane + 32), (unsigned long long)trow[lane + 32]);
        }
        __syncwarp();
    }
    if (active) p.state[c] = x;
    if (overflow) *p.error_flag = 1;
}

}  // namespace mmc
