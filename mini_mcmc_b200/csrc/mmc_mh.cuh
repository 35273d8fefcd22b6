// K1: batched Metropolis-Hastings, one chain per thread for the whole run.
// Reproduces MHMarkovChain::step (src/metropolis_hastings.rs:303-315) inside run_chain
// (src/core.rs:55-73): x' = proposal.sample(x); r = (lp' + q_b) - (lp + q_f); accept iff r > ln(u).
//
// This translation unit is compiled with -fmad=false: the f64 expressions below must round exactly
// like the CPU arithmetic (Rust does not contract a*b+c).
#pragma once

#include <type_traits>

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
//
// Draw write-out is the HBM-bound part (8 B per collected transition): each warp stages T steps of its
// 32 chains in shared memory and emits them as contiguous 256 B row segments (full 32 B sectors)
// instead of 32 strided 8 B stores per step.
//
// A transition consumes one 64-bit word `bits`: flip = bits & 1, u = (bits >> 11) * 2^-53.
//   threshold mode: accept  <=>  u53 < thr[x][dir]  <=>  bits <= lim[x][dir] = (thr << 11) - 1, where thr is
//     the host-built count of 53-bit uniforms satisfying (lp'+q_b)-(lp+q_f) > ln(u) (mmc_mh.cu).
//   log mode: evaluates that predicate in f64 on the device.
// One Philox4x32-10 call feeds two consecutive steps (global step parity picks the word pair).
struct MhPoissonParams {
    uint64_t *state;        // [chains] in/out
    uint64_t *out;          // [chains, n_collect]
    const uint8_t *flip;    // replay [chains, steps]
    const double *u;        // replay [chains, steps]
    const double *lnfact;   // [table_len]  sum_{i<=k} ln i, built by the host libm in the reference's order
    const uint2 *lim;       // [table_len][2] (lo, hi) of lim[k][0] = down, lim[k][1] = up
    int32_t table_len;
    double lambda, ln_lambda, ln_half;
    int64_t chains, chain_offset, step_base, n_collect, n_discard;
    uint32_t rk[20];        // Philox round keys (key + r * Weyl), host-expanded: they depend on the seed only
    int32_t *error_flag;    // set to 1 when a chain reaches the end of the table
};

constexpr int kPoisTile = 64;              // steps staged per write-out
constexpr int kPoisPitch = kPoisTile + 2;  // halfwords; 33 words -> conflict-free rows
constexpr int kPoisWarps = 8;

// Philox4x32-10 with the round keys read straight from the kernel-parameter constant bank.
__device__ __forceinline__ uint4 philox_rk(const uint32_t (&rk)[20], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

template <bool kReplay, bool kThreshold>
__global__ void __launch_bounds__(kPoisWarps * 32) mh_poisson_kernel(const __grid_constant__ MhPoissonParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [table_len][2] uint2 (threshold) or [table_len] f64 (log mode), then the warp tiles
    uint2 *s_lim = reinterpret_cast<uint2 *>(smem_raw);
    double *s_lnf = reinterpret_cast<double *>(smem_raw);
    uint16_t *tiles = reinterpret_cast<uint16_t *>(smem_raw + (size_t)p.table_len * 16);
    if (kThreshold) {
        for (int i = threadIdx.x; i < 2 * p.table_len; i += blockDim.x) s_lim[i] = p.lim[i];
    } else {
        for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) s_lnf[i] = p.lnfact[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t *tile = tiles + warp * 32 * kPoisPitch;
    uint16_t *my_row = tile + lane * kPoisPitch;
    const int64_t chain0 = ((int64_t)blockIdx.x * kPoisWarps + warp) * 32;
    if (chain0 >= p.chains) return;
    const int64_t c = chain0 + lane;
    const bool active = c < p.chains;
    const int64_t cc = active ? c : p.chains - 1;  // inactive lanes shadow the last chain (never stored)
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(cc + p.chain_offset);
    const uint32_t gc_lo = (uint32_t)gchain, gc_hi = (uint32_t)(gchain >> 32);
    const uint32_t kmax = (uint32_t)p.table_len - 1;
    uint32_t x, xmax = 0;
    {
        const uint64_t x0 = p.state[cc];
        x = x0 >= kmax ? kmax : (uint32_t)x0;
    }

    auto transition = [&](uint32_t lo, uint32_t hi) {
        // NonnegativeProposal::sample, examples/poisson_mh.rs:34-47: 0 -> 1, else +-1 by the flip
        const uint32_t up = (lo & 1u) | (uint32_t)(x == 0);
        const uint32_t y = min(x + 2u * up - 1u, kmax);
        bool acc;
        if (kThreshold) {
            const uint2 lim = s_lim[2u * x + up];
            acc = (hi < lim.y) || (hi == lim.y && lo <= lim.x);
        } else {
            const double u = (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
            // PoissonTarget::unnorm_logp, examples/poisson_mh.rs:19-25: -lambda + k ln(lambda) - ln k!
            const double cur_lp = (-p.lambda + (double)x * p.ln_lambda) - s_lnf[x];
            const double prop_lp = (-p.lambda + (double)y * p.ln_lambda) - s_lnf[y];
            // NonnegativeProposal::logp, examples/poisson_mh.rs:53-71 (y is always x +- 1 here)
            const double qf = (x == 0) ? 0.0 : p.ln_half;
            const double qb = (y == 0) ? ((x == 1) ? 0.0 : -INFINITY) : p.ln_half;
            const double r = (prop_lp + qb) - (cur_lp + qf);
            acc = (y != x) && (r > log(u));
        }
        x = acc ? y : x;
        xmax = max(xmax, x);
    };

    // ---- staging tile: column j holds global step G + j, G even, so that a Philox pair is one 32-bit store
    const uint64_t g_first = (uint64_t)(p.step_base + p.n_discard);  // global index of the first collected step
    int col_lo = (int)(g_first & 1);   // first valid column of the current tile (1 only for an odd start)
    int tpos = col_lo;                 // next column to fill
    int64_t t_base = -(int64_t)col_lo; // collected index of column 0
    // 16-byte stores need (chain * n_collect + t_base) even for every chain of the warp
    const bool vec_ok = (p.n_collect % 2 == 0) && col_lo == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    auto flush = [&]() {
        __syncwarp();
        const int nrows = (int)((p.chains - chain0 < 32) ? (p.chains - chain0) : 32);
        uint64_t *row = p.out + chain0 * p.n_collect + t_base;
        const uint16_t *trow = tile;
        if (vec_ok && (tpos % 2 == 0)) {
            for (int r = 0; r < nrows; ++r) {
                if (2 * lane < tpos) {
                    const uint32_t v = *reinterpret_cast<const uint32_t *>(trow + 2 * lane);
                    const ulonglong2 o = make_ulonglong2((unsigned long long)(v & 0xffffu), (unsigned long long)(v >> 16));
                    __stcs(reinterpret_cast<ulonglong2 *>(row + 2 * lane), o);
                }
                row += p.n_collect;
                trow += kPoisPitch;
            }
        } else {
            for (int r = 0; r < nrows; ++r) {
                if (lane >= col_lo && lane < tpos)
                    __stcs(reinterpret_cast<unsigned long long *>(row + lane), (unsigned long long)trow[lane]);
                if (lane + 32 < tpos)
                    __stcs(reinterpret_cast<unsigned long long *>(row + lane + 32), (unsigned long long)trow[lane + 32]);
                row += p.n_collect;
                trow += kPoisPitch;
            }
        }
        __syncwarp();
        t_base += tpos;
        tpos = 0;
        col_lo = 0;
    };
    auto emit1 = [&]() {
        my_row[tpos] = (uint16_t)x;
        if (++tpos == kPoisTile) flush();
    };

    auto replay_bits = [&](int64_t s, uint32_t &lo, uint32_t &hi) {
        const double u = p.u[cc * steps + s];
        const uint64_t bits = ((uint64_t)(u * 9007199254740992.0) << 11) | (uint64_t)(p.flip[cc * steps + s] & 1);
        lo = (uint32_t)bits;
        hi = (uint32_t)(bits >> 32);
    };

    // steps [s0, s1) of this run; kEmit selects the collect phase
    auto run_steps = [&](int64_t s0, int64_t s1, auto emit_tag) {
        constexpr bool kEmit = decltype(emit_tag)::value;
        int64_t s = s0;
        if (kReplay) {
            for (; s < s1; ++s) {
                uint32_t lo, hi;
                replay_bits(s, lo, hi);
                transition(lo, hi);
                if (kEmit) emit1();
            }
            return;
        }
        const uint64_t g0 = (uint64_t)p.step_base;
        if (s < s1 && ((g0 + s) & 1)) {  // odd first step: second half of its Philox pair
            const uint4 w = philox_rk(p.rk, gc_lo, gc_hi, (uint32_t)((g0 + s) >> 1), 0u);
            transition(w.z, w.w);
            if (kEmit) emit1();
            ++s;
        }
        uint32_t pair = (uint32_t)((g0 + s) >> 1);
        for (; s + 1 < s1; s += 2, ++pair) {
            const uint4 w = philox_rk(p.rk, gc_lo, gc_hi, pair, 0u);
            transition(w.x, w.y);
            const uint32_t xa = x;
            transition(w.z, w.w);
            if (kEmit) {  // tpos is even here: one 32-bit shared store for both steps
                *reinterpret_cast<uint32_t *>(my_row + tpos) = xa | (x << 16);
                tpos += 2;
                if (tpos == kPoisTile) flush();
            }
        }
        if (s < s1) {
            const uint4 w = philox_rk(p.rk, gc_lo, gc_hi, pair, 0u);
            transition(w.x, w.y);
            if (kEmit) emit1();
        }
    };
    run_steps(0, p.n_discard, std::false_type{});
    run_steps(p.n_discard, steps, std::true_type{});
    if (tpos > col_lo) flush();
    if (active) p.state[c] = x;
    if (xmax >= kmax) *p.error_flag = 1;
}

}  // namespace mmc
