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
    int64_t out_pitch;     // draws per chain row of `out` (>= n_collect; the caller may fill a window of a longer tensor)
    uint2 key;
    int32_t target_kind;   // MMC_T_GAUSSIAN2D | MMC_T_ISO_GAUSSIAN
    double tp[6];          // Gaussian2D: mean0, mean1, a, b, c, d ; Iso: std
    double prop_std;
    double prop_norm_term; // -D * 0.5 * ln(var * pi * std * std), evaluated by the host libm
    int32_t dim;           // state dimension (the any-dimension kernel reads it; the register kernels are templated on it)
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
            double *o = p.out + (c * p.out_pitch + (s - p.n_discard)) * D;
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

// ------------------------------------------------------------------ functor form (any Target / Proposal, f64 or f32 state)
// MetropolisHastings<S, T, D, Q> is generic over the state / float type and over the Target and Proposal traits
// (src/metropolis_hastings.rs:87,149-159; src/distributions.rs:92-108).  On the device a target is a functor with
//     static constexpr int kDim;  __device__ T unnorm_logp(const T (&x)[kDim]) const;
// and a proposal a functor with
//     template <int D> __device__ void sample(const T (&cur)[D], const T (&z)[D], T (&out)[D]) const;   // z: D standard normals
//     template <int D> __device__ T logp(const T (&from)[D], const T (&to)[D]) const;
// The step is MHMarkovChain::step (:303-315) verbatim; the built-in f64 pairs keep their dedicated kernel above, this one
// runs f32 state (every operation in f32 like MetropolisHastings<f32, f32, ..>) and the registered custom pairs
// (include/minimcmc_target.cuh).  Replay tapes are f64 arrays holding T-typed values; traces are written as f64.
template <class T> struct MhReal;
template <> struct MhReal<double> {
    static __device__ __forceinline__ double ln(double u) { return log(u); }
    static __device__ __forceinline__ double uniform(const uint4 &w) { return u53_half_open(w.x, w.y); }
    static constexpr int kPerBlock = 2;
    static __device__ __forceinline__ void normals(const uint4 &w, double (&n)[4]) { box_muller_f64(w, n[0], n[1]); n[2] = n[3] = 0.0; }
};
template <> struct MhReal<float> {
    static __device__ __forceinline__ float ln(float u) { return logf(u); }
    static __device__ __forceinline__ float uniform(const uint4 &w) { return u24_half_open(w.x); }
    static constexpr int kPerBlock = 4;
    static __device__ __forceinline__ void normals(const uint4 &w, float (&n)[4]) {
        box_muller_f32(w.x, w.y, n[0], n[1]);
        box_muller_f32(w.z, w.w, n[2], n[3]);
    }
};

// Gaussian2D<T>::unnorm_logp, src/distributions.rs:193-205, in T arithmetic
template <class T>
struct Gauss2DTargetF {
    static constexpr int kDim = 2;
    T m0, m1, i00, i01, i10, i11;
    __host__ explicit Gauss2DTargetF(const double *tp) {
        const T a = (T)tp[2], b = (T)tp[3], c = (T)tp[4], d = (T)tp[5];
        const T det = a * d - b * c;
        m0 = (T)tp[0]; m1 = (T)tp[1];
        i00 = d / det; i01 = -b / det; i10 = -c / det; i11 = a / det;
    }
    __device__ __forceinline__ T unnorm_logp(const T (&x)[2]) const {
        const T d0 = x[0] - m0, d1 = x[1] - m1;
        const T r0 = d0 * i00 + d1 * i10;
        const T r1 = d0 * i01 + d1 * i11;
        return (T)-0.5 * (r0 * d0 + r1 * d1);
    }
};
// IsotropicGaussian<T> as Target, src/distributions.rs:394-402
template <class T, int D>
struct IsoTargetF {
    static constexpr int kDim = D;
    T std;
    __host__ explicit IsoTargetF(const double *tp) : std((T)tp[0]) {}
    __device__ __forceinline__ T unnorm_logp(const T (&x)[D]) const {
        T sum = (T)0;
#pragma unroll
        for (int i = 0; i < D; ++i) sum = sum + x[i] * x[i];
        return (T)-0.5 * sum / (std * std);
    }
};
// IsotropicGaussian<T> as Proposal, src/distributions.rs:364-386 (the ln(var pi std std) normaliser kept verbatim; the
// logarithm is evaluated by the host libm)
template <class T>
struct IsoProposalF {
    T std, ln_term;
    __host__ explicit IsoProposalF(double s) : std((T)s) {
        const T var = std * std;
        ln_term = sizeof(T) == 8 ? (T)::log((double)(var * (T)M_PI * std * std)) : (T)::logf((float)(var * (T)M_PI * std * std));
    }
    template <int D>
    __device__ __forceinline__ void sample(const T (&cur)[D], const T (&z)[D], T (&out)[D]) const {
#pragma unroll
        for (int i = 0; i < D; ++i) out[i] = ((T)0 + std * z[i]) + cur[i];
    }
    template <int D>
    __device__ __forceinline__ T logp(const T (&from)[D], const T (&to)[D]) const {
        T lp = (T)0;
        const T var = std * std;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const T diff = to[i] - from[i];
            lp += -(diff * diff) / ((T)2 * var);
        }
        lp += -(T)D * (T)0.5 * ln_term;
        return lp;
    }
};

template <class TGT, class PROP, class T, int D, bool kReplay>
__global__ void __launch_bounds__(128) mh_functor_kernel(const TGT tgt, const PROP prop, const MhContParams p) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    T *state = reinterpret_cast<T *>(p.state), *out = reinterpret_cast<T *>(p.out);
    T x[D], y[D], z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = state[c * D + i];
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(c + p.chain_offset);
    for (int64_t s = 0; s < steps; ++s) {
        T u;
        if (kReplay) {
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] = (T)p.noise[(c * steps + s) * D + i];
            u = (T)p.u[c * steps + s];
        } else {
            const uint32_t gstep = (uint32_t)(p.step_base + s);
            u = MhReal<T>::uniform(philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, 0u)));
            constexpr int K = MhReal<T>::kPerBlock;
#pragma unroll
            for (int j = 0; j < (D + K - 1) / K; ++j) {
                T n[4];
                MhReal<T>::normals(philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)(1 + j))), n);
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (K * j + k < D) z[K * j + k] = n[k];
            }
        }
        prop.sample(x, z, y);
        const T cur_lp = tgt.unnorm_logp(x);
        const T prop_lp = tgt.unnorm_logp(y);
        const T qf = prop.logp(x, y);
        const T qb = prop.logp(y, x);
        const T r = (prop_lp + qb) - (cur_lp + qf);
        const bool acc = r > MhReal<T>::ln(u);
        if (acc) {
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] = y[i];
        }
        if (p.trace) {
            double *t = p.trace + (c * steps + s) * 4;
            t[0] = (double)cur_lp; t[1] = (double)prop_lp; t[2] = (double)r; t[3] = acc ? 1.0 : 0.0;
        }
        if (s >= p.n_discard && out) {
            T *o = out + (c * p.out_pitch + (s - p.n_discard)) * D;
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = x[i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) state[c * D + i] = x[i];
}

template <class TGT, class PROP, class T, int D>
int launch_mh_functor(const TGT &tgt, const PROP &prop, const MhContParams &p, bool replay, cudaStream_t stream) {
    const int block = 128;
    const unsigned grid = (unsigned)((p.chains + block - 1) / block);
    if (replay) mh_functor_kernel<TGT, PROP, T, D, true><<<grid, block, 0, stream>>>(tgt, prop, p);
    else mh_functor_kernel<TGT, PROP, T, D, false><<<grid, block, 0, stream>>>(tgt, prop, p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// IsotropicGaussian target + proposal for ANY dimension (<= kMhDynMax): the vectors live in local memory and the loops
// run to p.dim in the reference's (sequential) order, so the sums round exactly like the CPU code.
constexpr int kMhDynMax = 256;
template <class T, bool kReplay>
__global__ void __launch_bounds__(128) mh_iso_dyn_kernel(const MhContParams p, const T tstd, const T pstd, const T ln_term) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    const int D = p.dim;
    T *state = reinterpret_cast<T *>(p.state), *out = reinterpret_cast<T *>(p.out);
    T x[kMhDynMax], y[kMhDynMax];
    for (int i = 0; i < D; ++i) x[i] = state[c * D + i];
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gchain = (uint64_t)(c + p.chain_offset);
    auto tlogp = [&](const T *v) {
        T sum = (T)0;
        for (int i = 0; i < D; ++i) sum = sum + v[i] * v[i];
        return (T)-0.5 * sum / (tstd * tstd);
    };
    auto qlogp = [&](const T *from, const T *to) {
        T lp = (T)0;
        const T var = pstd * pstd;
        for (int i = 0; i < D; ++i) {
            const T diff = to[i] - from[i];
            lp += -(diff * diff) / ((T)2 * var);
        }
        lp += -(T)D * (T)0.5 * ln_term;
        return lp;
    };
    for (int64_t s = 0; s < steps; ++s) {
        T u;
        if (kReplay) {
            for (int i = 0; i < D; ++i) y[i] = ((T)0 + pstd * (T)p.noise[(c * steps + s) * D + i]) + x[i];
            u = (T)p.u[c * steps + s];
        } else {
            const uint32_t gstep = (uint32_t)(p.step_base + s);
            u = MhReal<T>::uniform(philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, 0u)));
            constexpr int K = MhReal<T>::kPerBlock;
            for (int j = 0; j < (D + K - 1) / K; ++j) {
                T n[4];
                MhReal<T>::normals(philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)(1 + j))), n);
                for (int k = 0; k < K; ++k)
                    if (K * j + k < D) y[K * j + k] = ((T)0 + pstd * n[k]) + x[K * j + k];
            }
        }
        const T cur_lp = tlogp(x), prop_lp = tlogp(y);
        const T qf = qlogp(x, y), qb = qlogp(y, x);
        const T r = (prop_lp + qb) - (cur_lp + qf);
        const bool acc = r > MhReal<T>::ln(u);
        if (acc)
            for (int i = 0; i < D; ++i) x[i] = y[i];
        if (p.trace) {
            double *t = p.trace + (c * steps + s) * 4;
            t[0] = (double)cur_lp; t[1] = (double)prop_lp; t[2] = (double)r; t[3] = acc ? 1.0 : 0.0;
        }
        if (s >= p.n_discard && out) {
            T *o = out + (c * p.out_pitch + (s - p.n_discard)) * D;
            for (int i = 0; i < D; ++i) o[i] = x[i];
        }
    }
    for (int i = 0; i < D; ++i) state[c * D + i] = x[i];
}

// ------------------------------------------------------------------ Poisson / integer state (config C2)
// PoissonTarget + NonnegativeProposal, examples/poisson_mh.rs:10-89.
//
// Draw write-out is the HBM-bound part (8 B per collected transition): each warp stages T steps of its
// 32 chains in shared memory and emits them as contiguous row segments (full 32 B sectors) instead of 32
// strided 8 B stores per step.
//
// RNG contract (minimcmc.h): global step s of a chain uses 16-bit field i = s & 7 of W = philox(key, (chain, s >> 3,
// sub 0)), i.e. h = (W[i >> 1] >> 16 (i & 1)) & 0xffff:   flip = h >> 15,  u15 = h & 0x7fff  (top 15 bits of u).
// The low 38 bits of the 53-bit uniform come from V = philox(key, (chain, s >> 3, sub 1 + (i >> 1))):
//   low38 = ((i & 1) ? V.w:V.z : V.y:V.x) >> 26,   u53 = u15 << 38 | low38,   u = u53 * 2^-53.
// The accept test u53 < thr is decided by the top 15 bits except on a tie (probability 2^-15 per step), so V is
// evaluated lazily: ONE Philox call feeds EIGHT transitions and the result is still the exact 53-bit test
// (an octet in which any lane-step tied is replayed through the exact path).
//   threshold mode: thr[x][dir] is the host-built count of 53-bit uniforms satisfying
//     (lp'+q_b)-(lp+q_f) > ln(u) (mmc_mh.cu), split as thr_hi = thr >> 38, thr_lo = thr & (2^38 - 1).
//   log mode: evaluates that predicate in f64 on the device (needs V every step).
struct MhPoissonParams {
    uint64_t *state;        // [chains] in/out
    void *out;              // [chains, n_collect] of OutT (u64 = reference layout, or the compact u8/u16 stream)
    const uint8_t *flip;    // replay [chains, steps]
    const double *u;        // replay [chains, steps]
    const double *lnfact;   // [table_len]  sum_{i<=k} ln i, built by the host libm in the reference's order
    const uint4 *lim;       // [table_len][2] (thr >> 38, low 32 of thr_lo, high 6 of thr_lo, 0) of [k][0] = down, [k][1] = up
    int32_t table_len;
    double lambda, ln_lambda, ln_half;
    int64_t chains, chain_offset, step_base, n_collect, n_discard;
    int64_t out_pitch;     // draws per chain row of `out` (>= n_collect; the caller may fill a window of a longer tensor)
    uint32_t rk[20];        // Philox round keys (key + r * Weyl), host-expanded: they depend on the seed only
    int32_t *error_flag;    // set to 1 when a chain reaches the end of the table
    uint32_t force_up_at;   // 0: NonnegativeProposal (a chain at 0 always proposes 1); 0xffffffff: reflecting +-1 walk
};

constexpr int kPoisWarps = 8;
// staging tile of one warp: 32 rows x T columns of Elem, row pitch padded by 4 bytes (odd number of words ->
// conflict-free row-per-lane stores and column-per-lane loads)
template <int T, class Elem>
struct PoisTile {
    static constexpr int kPitch = T + 4 / (int)sizeof(Elem);          // elements
    static constexpr size_t kWarpBytes = (size_t)32 * kPitch * sizeof(Elem);
};

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
__device__ __forceinline__ uint32_t pick(const uint4 &w, uint32_t i) {
    return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}
// 16-bit field i (0..7) of a Philox result
__device__ __forceinline__ uint32_t pick16(const uint4 &w, uint32_t i) { return (pick(w, i >> 1) >> (16u * (i & 1u))) & 0xffffu; }
// low 38 bits of the uniform of field i from the lazily evaluated call V (sub 1 + (i >> 1))
__device__ __forceinline__ uint64_t low38_of(const uint4 &v, uint32_t i) {
    const uint64_t bits = (i & 1u) ? (((uint64_t)v.w << 32) | v.z) : (((uint64_t)v.y << 32) | v.x);
    return bits >> 26;
}

template <bool kReplay, bool kThreshold, int T = 64, class Elem = uint16_t, class OutT = uint64_t>
__global__ void __launch_bounds__(kPoisWarps * 32, 6) mh_poisson_kernel(const __grid_constant__ MhPoissonParams p) {
    constexpr bool kWide = sizeof(OutT) == 8;   // false: OutT == Elem, draws leave the GPU in their compact form
    constexpr int kPoisTile = T;
    constexpr int kPoisPitch = PoisTile<T, Elem>::kPitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [table_len][2] uint2 (threshold) or [table_len] f64 (log mode), then the warp tiles
    uint32_t *s_hi = reinterpret_cast<uint32_t *>(smem_raw);  // [2 table_len] thr >> 38: one word per entry, conflict free
    double *s_lnf = reinterpret_cast<double *>(smem_raw);     // (the low 38 bits stay in global memory: tie path only)
    Elem *tiles = reinterpret_cast<Elem *>(smem_raw + (size_t)p.table_len * 8);
    if (kThreshold) {
        for (int i = threadIdx.x; i < 2 * p.table_len; i += blockDim.x) s_hi[i] = p.lim[i].x;
    } else {
        for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) s_lnf[i] = p.lnfact[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Elem *tile = tiles + warp * 32 * kPoisPitch;
    Elem *my_row = tile + lane * kPoisPitch;
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

    // One transition from the 16-bit field h; `low38()` yields the lazily evaluated low bits of u53.
    auto transition = [&](uint32_t h, auto low38) {
        // NonnegativeProposal::sample, examples/poisson_mh.rs:34-47: 0 -> 1, else +-1 by the flip
        const uint32_t up = (h >> 15) | (uint32_t)(x == p.force_up_at);
        const uint32_t y = min(x + 2u * up - 1u, kmax);
        const uint32_t u15 = h & 0x7fffu;
        bool acc;
        if (kThreshold) {
            const uint32_t hi = s_hi[2u * x + up];
            acc = u15 < hi;
            if (u15 == hi) {  // tie in the top 15 bits: decide on the remaining 38
                const uint4 t = __ldg(p.lim + 2u * x + up);
                acc = low38() < (((uint64_t)t.z << 32) | t.y);
            }
        } else {
            const uint64_t u53 = ((uint64_t)u15 << 38) | low38();
            const double u = (double)u53 * (1.0 / 9007199254740992.0);
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
    // Branch-free fast transition (threshold mode): decides on the top 15 bits and records whether a tie
    // occurred; the caller replays the whole octet through `transition` in that case.
    auto fast = [&](uint32_t h, bool &tie) {
        const uint32_t up = (h >> 15) | (uint32_t)(x == p.force_up_at);
        const uint32_t y = min(x + 2u * up - 1u, kmax);
        const uint32_t u15 = h & 0x7fffu;
        const uint32_t hi = s_hi[2u * x + up];
        tie = tie || (u15 == hi);
        x = (u15 < hi) ? y : x;
    };

    // ---- staging tile: column j holds global step G + j with G a multiple of 8, so the eight steps of one
    // Philox call land in aligned shared-memory stores
    const uint64_t g_first = (uint64_t)(p.step_base + p.n_discard);  // global index of the first collected step
    int col_lo = (int)(g_first & 7);   // first valid column of the current tile (non-zero only for an unaligned start)
    int tpos = col_lo;                 // next column to fill
    int64_t t_base = -(int64_t)col_lo; // collected index of column 0
    // vector stores need aligned row segments for every chain of the warp
    const bool vec_ok = kWide ? ((p.out_pitch % 2 == 0) && col_lo == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0))
                              : (((p.out_pitch * sizeof(OutT)) % 4 == 0) && col_lo == 0 &&
                                 ((reinterpret_cast<uintptr_t>(p.out) & 3) == 0));
    auto flush = [&]() {
        __syncwarp();
        const int nrows = (int)((p.chains - chain0 < 32) ? (p.chains - chain0) : 32);
        OutT *row = reinterpret_cast<OutT *>(p.out) + chain0 * p.out_pitch + t_base;
        const Elem *trow = tile;
        if (!kWide) {
            constexpr int kPer = 4 / (int)sizeof(Elem);   // columns per 32-bit word
            if (vec_ok && (tpos % kPer == 0)) {
                for (int r = 0; r < nrows; ++r) {
                    for (int col = kPer * lane; col < tpos; col += 32 * kPer)
                        __stcs(reinterpret_cast<unsigned int *>(row + col), *reinterpret_cast<const unsigned int *>(trow + col));
                    row += p.out_pitch;
                    trow += kPoisPitch;
                }
            } else {
                for (int r = 0; r < nrows; ++r) {
                    for (int col = lane; col < tpos; col += 32)
                        if (col >= col_lo) row[col] = (OutT)trow[col];
                    row += p.out_pitch;
                    trow += kPoisPitch;
                }
            }
        } else if (vec_ok && (tpos % 4 == 0) && (p.out_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0)) {
            // every warp-wide store instruction covers one contiguous 1 KB run of a row (full 32 B sectors): lane l
            // widens columns 4l .. 4l+3 of each 128-column group into one 256-bit streaming store
#pragma unroll 4
            for (int r = 0; r < nrows; ++r) {
#pragma unroll
                for (int g = 0; g < (kPoisTile + 127) / 128; ++g) {
                    const int col = 128 * g + 4 * lane;
                    if (col < tpos) {
                        unsigned long long a, b, c2, d;
                        if (sizeof(Elem) == 1) {
                            const uint32_t v = *reinterpret_cast<const uint32_t *>(trow + col);
                            a = v & 0xffu; b = (v >> 8) & 0xffu; c2 = (v >> 16) & 0xffu; d = v >> 24;
                        } else {
                            const uint32_t v0 = *reinterpret_cast<const uint32_t *>(trow + col);
                            const uint32_t v1 = *reinterpret_cast<const uint32_t *>(trow + col + 2);
                            a = v0 & 0xffffu; b = v0 >> 16; c2 = v1 & 0xffffu; d = v1 >> 16;
                        }
                        asm volatile("st.global.cs.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(row + col), "l"(a), "l"(b), "l"(c2), "l"(d)
                                     : "memory");
                    }
                }
                row += p.out_pitch;
                trow += kPoisPitch;
            }
        } else if (vec_ok && (tpos % 2 == 0)) {
            // every warp-wide store instruction covers one contiguous 512 B run of a row (full 32 B sectors):
            // lane l widens columns (2l, 2l+1) of each 64-column group into one 16-byte streaming store
#pragma unroll 4
            for (int r = 0; r < nrows; ++r) {
#pragma unroll
                for (int g = 0; g < kPoisTile / 64; ++g) {
                    const int col = 64 * g + 2 * lane;
                    if (col < tpos) {
                        uint32_t a, b;
                        if (sizeof(Elem) == 1) {
                            const uint32_t v = *reinterpret_cast<const uint16_t *>(trow + col);
                            a = v & 0xffu; b = v >> 8;
                        } else {
                            const uint32_t v = *reinterpret_cast<const uint32_t *>(trow + col);
                            a = v & 0xffffu; b = v >> 16;
                        }
                        __stcs(reinterpret_cast<ulonglong2 *>(row + col),
                               make_ulonglong2((unsigned long long)a, (unsigned long long)b));
                    }
                }
                row += p.out_pitch;
                trow += kPoisPitch;
            }
        } else {
            for (int r = 0; r < nrows; ++r) {
                for (int col = lane; col < tpos; col += 32)
                    if (col >= col_lo) row[col] = (OutT)trow[col];
                row += p.out_pitch;
                trow += kPoisPitch;
            }
        }
        __syncwarp();
        t_base += tpos;
        tpos = 0;
        col_lo = 0;
    };
    auto emit1 = [&]() {
        my_row[tpos] = (Elem)x;
        if (++tpos == kPoisTile) flush();
    };

    // steps [s0, s1) of this run; kEmit selects the collect phase
    auto run_steps = [&](int64_t s0, int64_t s1, auto emit_tag) {
        constexpr bool kEmit = decltype(emit_tag)::value;
        int64_t s = s0;
        if (kReplay) {
            for (; s < s1; ++s) {
                const double u = p.u[cc * steps + s];
                const uint64_t u53 = (uint64_t)(u * 9007199254740992.0);
                const uint32_t h = ((uint32_t)(p.flip[cc * steps + s] & 1) << 15) | (uint32_t)(u53 >> 38);
                transition(h, [&]() { return u53 & 0x3fffffffffULL; });
                if (kEmit) emit1();
            }
            return;
        }
        const uint64_t g0 = (uint64_t)p.step_base;
        auto single = [&]() {  // unaligned head / tail steps
            const uint64_t gs = g0 + s;
            const uint32_t oct = (uint32_t)(gs >> 3), i = (uint32_t)(gs & 7);
            const uint4 W = philox_rk(p.rk, gc_lo, gc_hi, oct, 0u);
            transition(pick16(W, i), [&]() { return low38_of(philox_rk(p.rk, gc_lo, gc_hi, oct, 1u + (i >> 1)), i); });
            if (kEmit) emit1();
            ++s;
        };
        while (s < s1 && ((g0 + s) & 7)) single();
        uint32_t oct = (uint32_t)((g0 + s) >> 3);
        for (; s + 7 < s1; s += 8, ++oct) {
            const uint4 W = philox_rk(p.rk, gc_lo, gc_hi, oct, 0u);
            uint32_t xs[8];
            auto exact_octet = [&]() {
#pragma unroll
                for (uint32_t i = 0; i < 8; ++i) {
                    transition(pick16(W, i), [&]() { return low38_of(philox_rk(p.rk, gc_lo, gc_hi, oct, 1u + (i >> 1)), i); });
                    xs[i] = x;
                }
            };
            if (kThreshold) {
                const uint32_t x_in = x;
                bool tie = false;
                fast(W.x & 0xffffu, tie); xs[0] = x;
                fast(W.x >> 16, tie); xs[1] = x;
                fast(W.y & 0xffffu, tie); xs[2] = x;
                fast(W.y >> 16, tie); xs[3] = x;
                fast(W.z & 0xffffu, tie); xs[4] = x;
                fast(W.z >> 16, tie); xs[5] = x;
                fast(W.w & 0xffffu, tie); xs[6] = x;
                fast(W.w >> 16, tie); xs[7] = x;
                if (tie) {  // exact 53-bit replay of this octet (2^-12 per lane-octet)
                    x = x_in;
                    exact_octet();
                }
                xmax = max(xmax, x + 7);  // x moves by at most 1 per step: bounds the excursion inside the octet
            } else {
                exact_octet();
            }
            if (kEmit) {  // tpos is a multiple of 8 here: aligned packed stores for the eight steps
                uint32_t *dst = reinterpret_cast<uint32_t *>(my_row + tpos);  // 4-byte aligned (row pitch is a whole number of words)
                if (sizeof(Elem) == 1) {
                    dst[0] = xs[0] | (xs[1] << 8) | (xs[2] << 16) | (xs[3] << 24);
                    dst[1] = xs[4] | (xs[5] << 8) | (xs[6] << 16) | (xs[7] << 24);
                } else {
                    dst[0] = xs[0] | (xs[1] << 16);
                    dst[1] = xs[2] | (xs[3] << 16);
                    dst[2] = xs[4] | (xs[5] << 16);
                    dst[3] = xs[6] | (xs[7] << 16);
                }
                tpos += 8;
                if (tpos == kPoisTile) flush();
            }
        }
        while (s < s1) single();
    };
    run_steps(0, p.n_discard, std::false_type{});
    run_steps(p.n_discard, steps, std::true_type{});
    if (tpos > col_lo) flush();
    if (active) p.state[c] = x;
    if (xmax >= kmax) *p.error_flag = 1;  // (conservative by 7 inside aligned octets)
}

}  // namespace mmc
