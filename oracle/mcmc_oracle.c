/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the mini-mcmc (v0.8.3) sampler hot path: MH / HMC / NUTS transitions and the
 * split-Rhat / ESS diagnostics.  The reference is Rust and cannot be compiled here (no rustc/cargo,
 * no vendored crates, no network), so this restatement is the parity oracle AND the reported CPU
 * baseline (bench.py cpu_baseline / --impl reference, "kind": "port").  It is pinned against every
 * golden vector the reference's own tests hold for this path (tests/test_oracle_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 * The product (mini_mcmc_b200 / libminimcmc.so) never links, imports or calls it.
 *
 * Reference lines followed (relative to /root/reference):
 *   MH step            src/metropolis_hastings.rs:303-315     run_chain   src/core.rs:55-73
 *   init / init_det    src/core.rs:394-435
 *   HMC step/leapfrog  src/hmc.rs:304-431                     HMC::run    src/hmc.rs:137-158
 *   NUTS               src/nuts.rs:410-996  (nuts_impl.inc)
 *   stats              src/stats.rs:310-336, 396-654
 *   targets/proposals  src/distributions.rs, examples/poisson_mh.rs (orc_targets.h)
 *
 * Build: gcc -O3 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_rng.h"
#include "orc_targets.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ RNG exports */

ORC_API void orc_init_tables(void) { orc_zig_init(); }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU-baseline legs ask for all host cores explicitly. */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

ORC_API void orc_smallrng_seed_state(uint64_t seed, uint64_t *state4) {
    orc_smallrng r;
    orc_smallrng_seed(&r, seed);
    memcpy(state4, r.s, 32);
}
/* kind: 0 u64, 1 f64 uniform, 2 f32 uniform (as double), 3 normal f64, 4 exp1 f64, 5 bool(0.5), 6 open01 */
ORC_API void orc_smallrng_fill(uint64_t *state4, int kind, double *out, uint64_t *out_u64, int64_t n) {
    orc_zig_init();
    orc_smallrng r;
    memcpy(r.s, state4, 32);
    for (int64_t i = 0; i < n; ++i) {
        switch (kind) {
        case 0: out_u64[i] = orc_next_u64(&r); break;
        case 1: out[i] = orc_next_f64(&r); break;
        case 2: out[i] = (double)orc_next_f32(&r); break;
        case 3: out[i] = orc_next_normal(&r); break;
        case 4: out[i] = orc_next_exp1(&r); break;
        case 5: out[i] = (double)orc_next_bool_half(&r); break;
        case 6: out[i] = orc_next_open01(&r); break;
        }
    }
    memcpy(state4, r.s, 32);
}

ORC_API void orc_philox(const uint32_t *key2, const uint32_t *ctr4, uint32_t *out4) {
    orc_philox4x32_10(key2, ctr4, out4);
}

/* _init, src/core.rs:421-435: n*d StandardNormal f64 draws from one SmallRng, row-major. */
ORC_API void orc_init_positions(double *out, int64_t n, int64_t d, uint64_t seed) {
    orc_zig_init();
    orc_smallrng r;
    orc_smallrng_seed(&r, seed);
    for (int64_t i = 0; i < n * d; ++i) out[i] = orc_next_normal(&r);
}

/* ------------------------------------------------------------------ target exports (KATs) */

ORC_API double orc_gaussian2d_logp_kat(const double *p6, const double *x2, int normalized) {
    return normalized ? orc_gaussian2d_logp(p6, x2) : orc_gaussian2d_unnorm_logp(p6, x2);
}
ORC_API double orc_iso_unnorm_logp_kat(double std, const double *x, int D) { return orc_iso_unnorm_logp(std, x, D); }
ORC_API double orc_iso_proposal_logp_kat(double std, const double *from, const double *to, int D) {
    return orc_iso_proposal_logp(std, from, to, D);
}
ORC_API double orc_poisson_logp_kat(double lambda, uint64_t k) { return orc_poisson_logp(lambda, k); }
ORC_API double orc_ln_factorial_kat(uint64_t k) { return orc_ln_factorial(k); }
ORC_API double orc_nonneg_logq_kat(uint64_t x, uint64_t y) { return orc_nonneg_logq(x, y); }

static void orc_make_target(orc_target *t, int kind, int dim, const double *params, int n_params, const float *vec,
                            const float *mat) {
    memset(t, 0, sizeof(*t));
    t->kind = kind;
    t->dim = dim;
    for (int i = 0; i < n_params && i < 8; ++i) t->p[i] = params[i];
    t->vec = vec;
    t->mat = mat;
    orc_target_prepare(t);
}

ORC_API float orc_target_logp_grad(int kind, int dim, const double *params, int n_params, const float *vec,
                                   const float *mat, const float *x, float *g) {
    orc_target t;
    orc_make_target(&t, kind, dim, params, n_params, vec, mat);
    return orc_logp_grad_f32(&t, x, g);
}

/* ------------------------------------------------------------------ Metropolis-Hastings */

/* One MH transition, src/metropolis_hastings.rs:303-315, continuous f64 state.
 * target kind GAUSSIAN2D (p6) or ISO_GAUSSIAN (p[0] = std); proposal IsotropicGaussian(prop_std).
 * noise[D] are the StandardNormal draws z_d; proposed_d = (0 + std*z_d) + x_d, the order of
 * Normal::sample (mean + std*z) followed by `x + *eps` (src/distributions.rs:364-372). */
static inline int orc_mh_cont_step(int kind, const double *tp, double prop_std, double *x, int D, const double *noise,
                                   double u, double *lp_out) {
    double prop[256];
    for (int i = 0; i < D; ++i) prop[i] = (0.0 + prop_std * noise[i]) + x[i];
    double cur_lp, prop_lp;
    if (kind == ORC_T_GAUSSIAN2D) {
        cur_lp = orc_gaussian2d_unnorm_logp(tp, x);
        prop_lp = orc_gaussian2d_unnorm_logp(tp, prop);
    } else {
        cur_lp = orc_iso_unnorm_logp(tp[0], x, D);
        prop_lp = orc_iso_unnorm_logp(tp[0], prop, D);
    }
    double qf = orc_iso_proposal_logp(prop_std, x, prop, D);
    double qb = orc_iso_proposal_logp(prop_std, prop, x, D);
    double r = (prop_lp + qb) - (cur_lp + qf);
    int acc = r > log(u);
    if (acc)
        for (int i = 0; i < D; ++i) x[i] = prop[i];
    if (lp_out) { lp_out[0] = cur_lp; lp_out[1] = prop_lp; lp_out[2] = r; }
    return acc;
}

/* run_chain (src/core.rs:55-73) over all chains with replayed noise[chains,steps,D] and u[chains,steps].
 * trace (optional) [chains,steps,4] = cur_lp, prop_lp, log_ratio, accepted.  state is updated in place. */
ORC_API int orc_mh_cont_run_replay(int kind, const double *tp, double prop_std, double *state, int64_t chains, int D,
                                   int64_t n_collect, int64_t n_discard, const double *noise, const double *u,
                                   double *out, double *trace) {
    if (D > 256) return -1;
    const int64_t steps = n_collect + n_discard;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        double *x = state + c * D;
        for (int64_t i = 0; i < steps; ++i) {
            double lp[3];
            int acc = orc_mh_cont_step(kind, tp, prop_std, x, D, noise + (c * steps + i) * D, u[c * steps + i], lp);
            if (trace) {
                double *tr = trace + (c * steps + i) * 4;
                tr[0] = lp[0]; tr[1] = lp[1]; tr[2] = lp[2]; tr[3] = (double)acc;
            }
            if (i >= n_discard) memcpy(out + (c * n_collect + (i - n_discard)) * D, x, sizeof(double) * D);
        }
    }
    return 0;
}

/* MetropolisHastings<f32, f32, ..> (the struct is generic over the state / float type, src/metropolis_hastings.rs:87):
 * the same step with every operation in f32 - Gaussian2D<f32> / IsotropicGaussian<f32> arithmetic
 * (src/distributions.rs:193-205,364-402), `u: f32 = rng.random()` and `u.ln()` in f32 (:309-311).  Tapes are f64 arrays
 * holding f32 values. */
static inline float orc_gaussian2d_unnorm_logp_f32(const float *p, const float *x) {
    float a = p[2], b = p[3], c = p[4], d = p[5];
    float det = a * d - b * c;
    float i00 = d / det, i01 = -b / det, i10 = -c / det, i11 = a / det;
    float d0 = x[0] - p[0], d1 = x[1] - p[1];
    float r0 = d0 * i00 + d1 * i10;
    float r1 = d0 * i01 + d1 * i11;
    return -0.5f * (r0 * d0 + r1 * d1);
}
static inline float orc_iso_unnorm_logp_f32(float std, const float *x, int D) {
    float sum = 0.0f;
    for (int i = 0; i < D; ++i) sum = sum + x[i] * x[i];
    return -0.5f * sum / (std * std);
}
static inline float orc_iso_proposal_logp_f32(float std, const float *from, const float *to, int D) {
    float lp = 0.0f;
    float d = (float)D;
    float two = 2.0f;
    float var = std * std;
    for (int i = 0; i < D; ++i) {
        float diff = to[i] - from[i];
        float exponent = -(diff * diff) / (two * var);
        lp += exponent;
    }
    lp += -d * 0.5f * logf(var * (float)M_PI * std * std);
    return lp;
}
ORC_API int orc_mh_cont_run_replay_f32(int kind, const double *tp, double prop_std, float *state, int64_t chains, int D,
                                       int64_t n_collect, int64_t n_discard, const double *noise, const double *u,
                                       float *out, float *trace) {
    if (D > 256) return -1;
    const int64_t steps = n_collect + n_discard;
    float tpf[6];
    for (int i = 0; i < 6; ++i) tpf[i] = (float)tp[i];
    const float std = (float)prop_std;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        float *x = state + c * D;
        float prop[256];
        for (int64_t i = 0; i < steps; ++i) {
            const double *z = noise + (c * steps + i) * D;
            for (int k = 0; k < D; ++k) prop[k] = (0.0f + std * (float)z[k]) + x[k];
            float cur_lp, prop_lp;
            if (kind == ORC_T_GAUSSIAN2D) {
                cur_lp = orc_gaussian2d_unnorm_logp_f32(tpf, x);
                prop_lp = orc_gaussian2d_unnorm_logp_f32(tpf, prop);
            } else {
                cur_lp = orc_iso_unnorm_logp_f32(tpf[0], x, D);
                prop_lp = orc_iso_unnorm_logp_f32(tpf[0], prop, D);
            }
            float qf = orc_iso_proposal_logp_f32(std, x, prop, D);
            float qb = orc_iso_proposal_logp_f32(std, prop, x, D);
            float r = (prop_lp + qb) - (cur_lp + qf);
            int acc = r > logf((float)u[c * steps + i]);
            if (acc) memcpy(x, prop, sizeof(float) * D);
            if (trace) {
                float *tr = trace + (c * steps + i) * 4;
                tr[0] = cur_lp; tr[1] = prop_lp; tr[2] = r; tr[3] = (float)acc;
            }
            if (i >= n_discard) memcpy(out + (c * n_collect + (i - n_discard)) * D, x, sizeof(float) * D);
        }
    }
    return 0;
}

/* The reference's own streams for MetropolisHastings<f64> + IsotropicGaussian:
 *  - accept uniforms: chain i uses SmallRng::seed_from_u64(1 + seed + i)  (src/metropolis_hastings.rs:187-193)
 *  - proposal noise : every chain holds a CLONE of the proposal, i.e. the same SmallRng(prop_seed) stream
 *    (:149-159); each sample() call draws D+1 normals and discards the last (zip polls the infinite
 *    sampler first, src/distributions.rs:364-372).
 * Fills noise[chains,steps,D] and u[chains,steps] so the tape can be replayed on both sides. */
ORC_API void orc_mh_cont_reference_tape(uint64_t chain_seed, uint64_t prop_seed, int64_t chains, int64_t steps, int D,
                                        double *noise, double *u) {
    orc_zig_init();
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        orc_smallrng cr, pr;
        orc_smallrng_seed(&cr, 1 + chain_seed + (uint64_t)c);
        orc_smallrng_seed(&pr, prop_seed);
        for (int64_t i = 0; i < steps; ++i) {
            for (int d = 0; d < D; ++d) noise[(c * steps + i) * D + d] = orc_next_normal(&pr);
            (void)orc_next_normal(&pr); /* the D+1-th draw, discarded */
            u[c * steps + i] = orc_next_f64(&cr);
        }
    }
}

/* Poisson MH transition (examples/poisson_mh.rs + src/metropolis_hastings.rs:303-315), u64 state. */
static inline uint64_t orc_mh_poisson_step(double lambda, uint64_t x, int flip, double u) {
    uint64_t y = (x == 0) ? 1 : (flip ? x + 1 : x - 1);
    double cur_lp = orc_poisson_logp(lambda, x);
    double prop_lp = orc_poisson_logp(lambda, y);
    double qf = orc_nonneg_logq(x, y);
    double qb = orc_nonneg_logq(y, x);
    double r = (prop_lp + qb) - (cur_lp + qf);
    return (r > log(u)) ? y : x;
}

ORC_API int orc_mh_poisson_run_replay(double lambda, uint64_t *state, int64_t chains, int64_t n_collect,
                                      int64_t n_discard, const uint8_t *flip, const double *u, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        uint64_t x = state[c];
        for (int64_t i = 0; i < steps; ++i) {
            x = orc_mh_poisson_step(lambda, x, flip[c * steps + i], u[c * steps + i]);
            if (i >= n_discard) out[c * n_collect + (i - n_discard)] = x;
        }
        state[c] = x;
    }
    return 0;
}

/* Twin of the CUDA path's native Philox keying (include/minimcmc.h "RNG contract"):
 * key = seed; global step s uses 16-bit field i = s & 7 of W = philox(ctr = (chain_lo, chain_hi, s >> 3, 0)):
 * h = (W[i >> 1] >> 16 (i & 1)) & 0xffff, flip = h >> 15, u15 = h & 0x7fff; the low 38 bits come from
 * V = philox(ctr = (.., s >> 3, 1 + (i >> 1))): low38 = ((i & 1) ? V[3]:V[2] : V[1]:V[0]) >> 26;
 * u = (u15 << 38 | low38) * 2^-53. */
ORC_API int orc_mh_poisson_run_philox(double lambda, uint64_t *state, int64_t chains, int64_t chain_offset,
                                      int64_t step_base, int64_t n_collect, int64_t n_discard, uint64_t seed,
                                      uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        uint64_t x = state[c];
        const uint64_t gc = (uint64_t)(c + chain_offset);
        for (int64_t s = 0; s < steps; ++s) {
            const uint64_t gs = (uint64_t)(step_base + s);
            const uint32_t i = (uint32_t)(gs & 7);
            uint32_t ctr[4] = {(uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)(gs >> 3), 0u};
            uint32_t w[4], v[4];
            orc_philox4x32_10(key, ctr, w);
            ctr[3] = 1u + (i >> 1);
            orc_philox4x32_10(key, ctr, v);
            const uint32_t h = (w[i >> 1] >> (16u * (i & 1u))) & 0xffffu;
            const int flip = (int)(h >> 15);
            const uint64_t bits = (i & 1u) ? (((uint64_t)v[3] << 32) | v[2]) : (((uint64_t)v[1] << 32) | v[0]);
            const uint64_t u53 = ((uint64_t)(h & 0x7fffu) << 38) | (bits >> 26);
            const double u = (double)u53 * (1.0 / 9007199254740992.0);
            x = orc_mh_poisson_step(lambda, x, flip, u);
            if (s >= n_discard) out[c * n_collect + (s - n_discard)] = x;
        }
        state[c] = x;
    }
    return 0;
}

/* The reference CPU path for Poisson MH as it runs today (rayon over chains, one SmallRng per chain for
 * the accept uniform seeded 1+seed+i; the example's flips come from ThreadRng which is not reproducible,
 * so a per-chain SmallRng(flip_seed + i) stands in for it).  Used as the CPU baseline. */
ORC_API int orc_mh_poisson_run_reference(double lambda, uint64_t *state, int64_t chains, int64_t n_collect,
                                         int64_t n_discard, uint64_t seed, uint64_t flip_seed, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < chains; ++c) {
        orc_smallrng cr, fr;
        orc_smallrng_seed(&cr, 1 + seed + (uint64_t)c);
        orc_smallrng_seed(&fr, flip_seed + (uint64_t)c);
        uint64_t x = state[c];
        for (int64_t i = 0; i < steps; ++i) {
            int flip = 0;
            if (x != 0) flip = orc_next_bool_half(&fr); /* no draw when x == 0, examples/poisson_mh.rs:37-40 */
            double u = orc_next_f64(&cr);
            x = orc_mh_poisson_step(lambda, x, flip, u);
            if (i >= n_discard) out[c * n_collect + (i - n_discard)] = x;
        }
        state[c] = x;
    }
    return 0;
}

/* ------------------------------------------------------------------ HMC (batched, f32) */

/* HMC::step + leapfrog, src/hmc.rs:304-431, for one chain (the chain axis is a pure batch axis).
 * eps is T (=f32 for HMC<f32,..>); eps_half = step_size * 0.5 computed in T before the multiply
 * (:323-324, :420).  trace[4] = logp_current, logp_proposed, accept_logp, accepted. */
static int orc_hmc_step_chain(const orc_target *t, float *x, const float *mom0, float u, float eps, int L,
                              float *scratch, float *trace) {
    const int D = t->dim;
    float *pos = scratch, *mom = scratch + D, *g = scratch + 2 * D, *lgs = scratch + 3 * D;
    const float eps_half = eps * 0.5f;
    float logp_cur = orc_logp_grad_f32(t, x, g);
    float ke = 0.0f;
    for (int i = 0; i < D; ++i) {
        lgs[i] = g[i] * eps_half;
        mom[i] = mom0[i];
        pos[i] = x[i];
        ke = ke + mom0[i] * mom0[i];
    }
    float h_cur = -logp_cur + ke * 0.5f;
    for (int l = 0; l < L; ++l) {
        for (int i = 0; i < D; ++i) {
            mom[i] = mom[i] + lgs[i];
            pos[i] = pos[i] + mom[i] * eps;
        }
        (void)orc_logp_grad_f32(t, pos, g);
        for (int i = 0; i < D; ++i) {
            lgs[i] = g[i] * eps_half;
            mom[i] = mom[i] + lgs[i];
        }
    }
    float logp_prop = orc_logp_grad_f32(t, pos, g); /* logp_final, src/hmc.rs:429 */
    float ke2 = 0.0f;
    for (int i = 0; i < D; ++i) ke2 = ke2 + mom[i] * mom[i];
    float h_prop = -logp_prop + ke2 * 0.5f;
    float accept_logp = h_cur - h_prop;
    int acc = accept_logp >= logf(u); /* non-strict, NaN -> reject (src/hmc.rs:366-367) */
    if (acc)
        for (int i = 0; i < D; ++i) x[i] = pos[i];
    if (trace) { trace[0] = logp_cur; trace[1] = logp_prop; trace[2] = accept_logp; trace[3] = (float)acc; }
    return acc;
}

/* HMC::run, src/hmc.rs:137-158, with replayed momenta[steps,chains,D] and u[steps,chains] (the
 * reference's own come from burn's backend RNG, which set_seed does not control: replay only).
 * out is [chains, n_collect, D] (the permuted view the reference returns).  positions updated in place. */
ORC_API int orc_hmc_run_replay(int kind, int D, const double *params, int n_params, const float *vec,
                               const float *mat, float *positions, int64_t chains, double step_size, int L,
                               int64_t n_collect, int64_t n_discard, const float *momenta, const float *u, float *out,
                               float *trace) {
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
    const int64_t steps = n_collect + n_discard;
    const float eps = (float)step_size;
#pragma omp parallel
    {
        float *scratch = (float *)malloc(sizeof(float) * D * 4);
#pragma omp for schedule(static)
        for (int64_t c = 0; c < chains; ++c) {
            float *x = positions + c * D;
            for (int64_t s = 0; s < steps; ++s) {
                orc_hmc_step_chain(&t, x, momenta + (s * chains + c) * D, u[s * chains + c], eps, L, scratch,
                                   trace ? trace + (s * chains + c) * 4 : NULL);
                if (s >= n_discard && out) memcpy(out + (c * n_collect + (s - n_discard)) * D, x, sizeof(float) * D);
            }
        }
        free(scratch);
    }
    return 0;
}

/* CPU baseline for HMC: same transition, momenta/uniforms from a per-chain SmallRng + ziggurat
 * (a generous stand-in for burn's backend RNG). */
ORC_API int orc_hmc_run_reference(int kind, int D, const double *params, int n_params, const float *vec,
                                  const float *mat, float *positions, int64_t chains, double step_size, int L,
                                  int64_t n_collect, int64_t n_discard, uint64_t seed, float *out) {
    orc_zig_init();
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
    const int64_t steps = n_collect + n_discard;
    const float eps = (float)step_size;
#pragma omp parallel
    {
        float *scratch = (float *)malloc(sizeof(float) * D * 5);
        float *mom0 = scratch + 4 * D;
#pragma omp for schedule(dynamic, 8)
        for (int64_t c = 0; c < chains; ++c) {
            orc_smallrng r;
            orc_smallrng_seed(&r, seed + (uint64_t)c + 1);
            float *x = positions + c * D;
            for (int64_t s = 0; s < steps; ++s) {
                for (int i = 0; i < D; ++i) mom0[i] = (float)orc_next_normal(&r);
                float u = orc_next_f32(&r);
                orc_hmc_step_chain(&t, x, mom0, u, eps, L, scratch, NULL);
                if (s >= n_discard && out) memcpy(out + (c * n_collect + (s - n_discard)) * D, x, sizeof(float) * D);
            }
        }
        free(scratch);
    }
    return 0;
}

/* ------------------------------------------------------------------ NUTS */

enum { ORC_SRC_RNG = 0, ORC_SRC_TAPE = 1 };
typedef struct {
    int mode;
    orc_smallrng rng;
    double *normals, *exps, *unifs; /* tape (read) or recording buffers (write, may be NULL) */
    int64_t n_normals, n_exps, n_unifs;
    int64_t cap_normals, cap_exps, cap_unifs;
    /* test bookkeeping (not part of the algorithm): the smallest relative distance of any accept / slice / U-turn
     * comparison of the current transition from its threshold, when track_margin is set.  Parity tests use it to
     * tell a chain that legitimately branches at an fp32 near-tie from a wrong result. */
    int track_margin;
    double min_margin;
} orc_src;
static inline void orc_margin(orc_src *s, double m) {
    if (s->track_margin && m < s->min_margin) s->min_margin = m;
}

#define ST double
#define SFX(n) n##_f64
#define ST_EXP exp
#define ST_LOG log
#define ST_SQRT sqrt
#define ST_POW pow
#define ST_EPS 2.220446049250313e-16
#define ST_IS_F32 0
#include "nuts_impl.inc"
#undef ST
#undef SFX
#undef ST_EXP
#undef ST_LOG
#undef ST_SQRT
#undef ST_POW
#undef ST_EPS
#undef ST_IS_F32

#define ST float
#define SFX(n) n##_f32
#define ST_EXP expf
#define ST_LOG logf
#define ST_SQRT sqrtf
#define ST_POW powf
#define ST_EPS 1.1920929e-07f
#define ST_IS_F32 1
#include "nuts_impl.inc"
#undef ST
#undef SFX

/* find_reasonable_epsilon KAT entry (src/nuts.rs:1049-1055). */
ORC_API double orc_nuts_find_reasonable_epsilon(int kind, int D, const double *params, int n_params, const float *vec,
                                                const float *mat, const float *x, const float *p, int scalar_f32) {
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
    return scalar_f32 ? (double)find_reasonable_epsilon_f32(&t, x, p, NULL)
                      : find_reasonable_epsilon_f64(&t, x, p, NULL);
}

/* build_tree KAT entry (src/nuts.rs:1057-1121).  vec_out[8*D] = x-, p-, g-, x+, p+, g+, x', g';
 * scal_out[5] = logp', n', s', alpha', n_alpha'.  Uniforms from SmallRng(rng_seed). */
ORC_API void orc_nuts_build_tree(int kind, int D, const double *params, int n_params, const float *vec,
                                 const float *mat, const float *x, const float *p, const float *g, double logu, int v,
                                 int j, double epsilon, double joint_0, uint64_t rng_seed, float *vec_out,
                                 double *scal_out) {
    orc_zig_init();
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
    orc_src src;
    memset(&src, 0, sizeof(src));
    src.mode = ORC_SRC_RNG;
    orc_smallrng_seed(&src.rng, rng_seed);
    tree_f64 r;
    tree_alloc_f64(&r, D);
    build_tree_f64(&t, x, p, g, logu, v, j, epsilon, joint_0, &src, &r, NULL);
    memcpy(vec_out, r.xm, sizeof(float) * D * 8);
    scal_out[0] = r.logp_prime; scal_out[1] = (double)r.n_prime; scal_out[2] = (double)r.s_prime;
    scal_out[3] = r.alpha; scal_out[4] = (double)r.n_alpha;
    tree_free_f64(&r);
}

/* build_tree for many chains with the uniforms read from per-chain tapes (both scalar types): the checker of the
 * CUDA debug entry mmc_nuts_build_tree.  x, p, g [chains, D]; scal_in [chains, 4] = logu, v, epsilon, joint_0;
 * vec_out [chains, 8, D] and scal_out [chains, 6] = logp', n', s', alpha', n_alpha', uniforms consumed. */
ORC_API void orc_nuts_build_tree_tape(int kind, int D, const double *params, int n_params, const float *vec,
                                      const float *mat, int64_t chains, const float *x, const float *p, const float *g,
                                      const double *scal_in, int j, int scalar_f32, double *unifs, int64_t cap_unifs,
                                      float *vec_out, double *scal_out, double *margin_out) {
    orc_zig_init();
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < chains; ++c) {
        orc_src src;
        memset(&src, 0, sizeof(src));
        src.mode = ORC_SRC_TAPE;
        src.unifs = unifs + c * cap_unifs;
        src.cap_unifs = cap_unifs;
        src.track_margin = margin_out != NULL;
        src.min_margin = 1e300;
        const double *si = scal_in + c * 4;
        double *so = scal_out + c * 6;
        if (scalar_f32) {
            tree_f32 r;
            tree_alloc_f32(&r, D);
            build_tree_f32(&t, x + c * D, p + c * D, g + c * D, (float)si[0], (int)si[1], j, (float)si[2], (float)si[3], &src, &r, NULL);
            memcpy(vec_out + c * 8 * D, r.xm, sizeof(float) * D * 8);
            so[0] = r.logp_prime; so[1] = (double)r.n_prime; so[2] = (double)r.s_prime; so[3] = (double)r.alpha; so[4] = (double)r.n_alpha;
            tree_free_f32(&r);
        } else {
            tree_f64 r;
            tree_alloc_f64(&r, D);
            build_tree_f64(&t, x + c * D, p + c * D, g + c * D, si[0], (int)si[1], j, si[2], si[3], &src, &r, NULL);
            memcpy(vec_out + c * 8 * D, r.xm, sizeof(float) * D * 8);
            so[0] = r.logp_prime; so[1] = (double)r.n_prime; so[2] = (double)r.s_prime; so[3] = r.alpha; so[4] = (double)r.n_alpha;
            tree_free_f64(&r);
        }
        so[5] = (double)src.n_unifs;
        if (margin_out) margin_out[c] = src.min_margin;
    }
}

/* ONE NUTSChain::step per chain (src/nuts.rs:550-691) from a given position and adaptation state, draws from per-chain
 * tapes (D normals, one Exp(1), the uniform tape): the checker of the single-transition parity tests.
 * state_io [chains, 5] as in orc_nuts_run (epsilon must be set); n_discard only enters the dual-averaging rule
 * (m <= n_discard adapts).  trace_out [chains, 8] = joint_0, logu, n, alpha, n_alpha, depth, epsilon used, uniforms.
 * margin_out [chains] (optional) = smallest relative distance of any comparison of the transition from its threshold. */
ORC_API int orc_nuts_step_trace(int kind, int D, const double *params, int n_params, const float *vec, const float *mat,
                                float *positions, int64_t chains, double target_accept, int scalar_f32, int64_t n_discard,
                                int max_depth, double *normals, int64_t cap_normals, double *exps, int64_t cap_exps,
                                double *unifs, int64_t cap_unifs, double *state_io, double *trace_out,
                                double *margin_out) {
    orc_zig_init();
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < chains; ++c) {
        orc_src src;
        memset(&src, 0, sizeof(src));
        src.mode = ORC_SRC_TAPE;
        src.normals = normals + c * cap_normals; src.cap_normals = cap_normals;
        src.exps = exps + c * cap_exps; src.cap_exps = cap_exps;
        src.unifs = unifs + c * cap_unifs; src.cap_unifs = cap_unifs;
        src.track_margin = margin_out != NULL;
        src.min_margin = 1e300;
        double *s = state_io + c * 5;
        int depth = 0;
        if (scalar_f32) {
            chain_state_f32 cs;
            chain_new_f32(&cs, target_accept);
            cs.epsilon = (float)s[0]; cs.epsilon_bar = (float)s[1]; cs.h_bar = (float)s[2]; cs.mu = (float)s[3];
            cs.m = (int64_t)s[4]; cs.n_discard = n_discard;
            chain_step_traced_f32(&t, &cs, positions + c * D, &src, max_depth, NULL, &depth, trace_out + c * 8);
            s[0] = cs.epsilon; s[1] = cs.epsilon_bar; s[2] = cs.h_bar; s[3] = cs.mu; s[4] = (double)cs.m;
        } else {
            chain_state_f64 cs;
            chain_new_f64(&cs, target_accept);
            cs.epsilon = s[0]; cs.epsilon_bar = s[1]; cs.h_bar = s[2]; cs.mu = s[3]; cs.m = (int64_t)s[4];
            cs.n_discard = n_discard;
            chain_step_traced_f64(&t, &cs, positions + c * D, &src, max_depth, NULL, &depth, trace_out + c * 8);
            s[0] = cs.epsilon; s[1] = cs.epsilon_bar; s[2] = cs.h_bar; s[3] = cs.mu; s[4] = (double)cs.m;
        }
        if (margin_out) margin_out[c] = src.min_margin;
    }
    return 0;
}

/* Full multi-chain NUTS run.
 *  mode 0: reference streams — chain i draws from SmallRng::seed_from_u64(seed + i + 1)
 *          (NUTS::set_seed, src/nuts.rs:347-353); if tape buffers are given the draws are RECORDED so the
 *          same values can be replayed into the CUDA path.
 *  mode 1: replay — draws are read from the tapes.
 * Tapes are per chain: normals[chains, cap_normals], exps[chains, cap_exps], unifs[chains, cap_unifs];
 * counts_out[chains,3] returns how many of each were consumed.
 * progress: 0 = NUTS::run semantics (n_collect+n_discard-1 steps, slot 0 = start), 1 = run_progress.
 * state_io[chains,5] (optional, in/out) = epsilon, epsilon_bar, h_bar, mu, m  (epsilon = -1: unset).
 * depths[chains, steps] optional; n_grad_out[chains] optional. */
ORC_API int orc_nuts_run(int kind, int D, const double *params, int n_params, const float *vec, const float *mat,
                         float *positions, int64_t chains, double target_accept, int scalar_f32, int mode,
                         uint64_t seed, int64_t n_collect, int64_t n_discard, int progress, int max_depth, float *out,
                         double *normals, int64_t cap_normals, double *exps, int64_t cap_exps, double *unifs,
                         int64_t cap_unifs, int64_t *counts_out, double *state_io, int32_t *depths,
                         int64_t *n_grad_out) {
    orc_zig_init();
    orc_target t;
    orc_make_target(&t, kind, D, params, n_params, vec, mat);
    const int64_t steps_total = n_collect + n_discard;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < chains; ++c) {
        orc_src src;
        memset(&src, 0, sizeof(src));
        src.mode = mode;
        if (mode == ORC_SRC_RNG) orc_smallrng_seed(&src.rng, seed + (uint64_t)c + 1);
        if (normals) { src.normals = normals + c * cap_normals; src.cap_normals = cap_normals; }
        if (exps) { src.exps = exps + c * cap_exps; src.cap_exps = cap_exps; }
        if (unifs) { src.unifs = unifs + c * cap_unifs; src.cap_unifs = cap_unifs; }
        int64_t n_grad = 0;
        float *pos = positions + c * D;
        float *o = out + c * n_collect * D;
        int32_t *dp = depths ? depths + c * steps_total : NULL;
        if (scalar_f32) {
            chain_state_f32 cs;
            chain_new_f32(&cs, target_accept);
            if (state_io) {
                double *s = state_io + c * 5;
                cs.epsilon = (float)s[0]; cs.epsilon_bar = (float)s[1]; cs.h_bar = (float)s[2]; cs.mu = (float)s[3];
                cs.m = (int64_t)s[4];
            }
            chain_run_f32(&t, &cs, pos, &src, n_collect, n_discard, progress, max_depth, o, &n_grad, dp);
            if (state_io) {
                double *s = state_io + c * 5;
                s[0] = cs.epsilon; s[1] = cs.epsilon_bar; s[2] = cs.h_bar; s[3] = cs.mu; s[4] = (double)cs.m;
            }
        } else {
            chain_state_f64 cs;
            chain_new_f64(&cs, target_accept);
            if (state_io) {
                double *s = state_io + c * 5;
                cs.epsilon = s[0]; cs.epsilon_bar = s[1]; cs.h_bar = s[2]; cs.mu = s[3]; cs.m = (int64_t)s[4];
            }
            chain_run_f64(&t, &cs, pos, &src, n_collect, n_discard, progress, max_depth, o, &n_grad, dp);
            if (state_io) {
                double *s = state_io + c * 5;
                s[0] = cs.epsilon; s[1] = cs.epsilon_bar; s[2] = cs.h_bar; s[3] = cs.mu; s[4] = (double)cs.m;
            }
        }
        if (counts_out) {
            counts_out[c * 3 + 0] = src.n_normals; counts_out[c * 3 + 1] = src.n_exps; counts_out[c * 3 + 2] = src.n_unifs;
        }
        if (n_grad_out) n_grad_out[c] = n_grad;
    }
    return 0;
}

/* ------------------------------------------------------------------ stats (src/stats.rs) */

/* autocov_bf, src/stats.rs:632-654.  data [n,d] row-major -> out [n,d]. */
ORC_API void orc_autocov_bf(const float *data, int64_t n, int64_t d, float *out) {
    float *col = (float *)malloc(sizeof(float) * n);
    for (int64_t c = 0; c < d; ++c) {
        float sum = 0.0f;
        for (int64_t t = 0; t < n; ++t) sum += data[t * d + c];
        float mean = sum / (float)n;
        for (int64_t t = 0; t < n; ++t) col[t] = data[t * d + c] - mean;
        for (int64_t lag = 0; lag < n; ++lag) {
            float s = 0.0f;
            for (int64_t t = 0; t < n - lag; ++t) s += col[t] * col[t + lag];
            out[lag * d + c] = s / (float)n;
        }
    }
    free(col);
}

/* radix-2 complex FFT in f32 (rustfft 6.4.1, Cargo.lock:4877, is not in the tree; any exact-arithmetic
 * equivalent DFT differs from it only by f32 rounding, ~1e-7 relative). */
static void orc_fft_f32(float *re, float *im, int64_t n, int inverse) {
    for (int64_t i = 1, j = 0; i < n; ++i) {
        int64_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            float tr = re[i]; re[i] = re[j]; re[j] = tr;
            float ti = im[i]; im[i] = im[j]; im[j] = ti;
        }
    }
    for (int64_t len = 2; len <= n; len <<= 1) {
        double ang = 2.0 * M_PI / (double)len * (inverse ? 1.0 : -1.0);
        for (int64_t i = 0; i < n; i += len) {
            for (int64_t k = 0; k < len / 2; ++k) {
                float wr = (float)cos(ang * (double)k), wi = (float)sin(ang * (double)k);
                int64_t a = i + k, b = i + k + len / 2;
                float xr = re[b] * wr - im[b] * wi;
                float xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] = re[a] + xr; im[a] = im[a] + xi;
            }
        }
    }
}

/* autocov_fft, src/stats.rs:576-620. */
ORC_API void orc_autocov_fft(const float *data, int64_t n, int64_t d, float *out) {
    int64_t n_padded = 1;
    while (n_padded < 2 * n - 1) n_padded <<= 1;
    float *re = (float *)malloc(sizeof(float) * n_padded * 2);
    float *im = re + n_padded;
    for (int64_t c = 0; c < d; ++c) {
        float sum = 0.0f;
        for (int64_t t = 0; t < n; ++t) sum += data[t * d + c];
        float mean = sum / (float)n;
        for (int64_t t = 0; t < n_padded; ++t) { re[t] = t < n ? data[t * d + c] - mean : 0.0f; im[t] = 0.0f; }
        orc_fft_f32(re, im, n_padded, 0);
        for (int64_t t = 0; t < n_padded; ++t) { re[t] = re[t] * re[t] + im[t] * im[t]; im[t] = 0.0f; }
        orc_fft_f32(re, im, n_padded, 1);
        for (int64_t t = 0; t < n; ++t) out[t * d + c] = re[t] / (float)n_padded / (float)n;
    }
    free(re);
}

/* split_rhat_mean_ess, src/stats.rs:396-554.  sample [c,n,p] f32 -> rhat[p], ess[p]. */
ORC_API int orc_split_rhat_mean_ess(const float *sample, int64_t c, int64_t n, int64_t p, float *rhat_out,
                                    float *ess_out) {
    const int64_t half = n / 2;
    const int64_t C = 2 * c, N = half;
    if (N < 1) return -1;
    /* splitcat :396-402: first `half` draws of every chain, then the LAST `half` draws of every chain */
    float *split = (float *)malloc(sizeof(float) * C * N * p);
    for (int64_t j = 0; j < c; ++j) {
        memcpy(split + (j * N) * p, sample + (j * n) * p, sizeof(float) * N * p);
        memcpy(split + ((c + j) * N) * p, sample + (j * n + (n - half)) * p, sizeof(float) * N * p);
    }
    float *within = (float *)malloc(sizeof(float) * p * 2);
    float *var = within + p;
    /* withinvar :429-477 */
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < p; ++q) {
        float *cm = (float *)malloc(sizeof(float) * C);
        float msum = 0.0f;
        for (int64_t j = 0; j < C; ++j) {
            float s = 0.0f;
            for (int64_t t = 0; t < N; ++t) s += split[(j * N + t) * p + q];
            cm[j] = s / (float)N;
            msum += cm[j];
        }
        float overall = msum / (float)C;
        float dsum = 0.0f;
        for (int64_t j = 0; j < C; ++j) { float dd = cm[j] - overall; dsum += dd * dd; }
        float b = dsum * ((float)N / (float)(C - 1));
        float wsum = 0.0f;
        for (int64_t j = 0; j < C; ++j) {
            float s = 0.0f;
            for (int64_t t = 0; t < N; ++t) { float v = split[(j * N + t) * p + q]; s += (v - cm[j]) * (v - cm[j]); }
            wsum += s / (float)N;
        }
        float w = wsum / (float)C;
        within[q] = w;
        var[q] = (((float)N - 1.0f) / (float)N) * w + b / (float)N;
        free(cm);
    }
    for (int64_t q = 0; q < p; ++q) rhat_out[q] = sqrtf(within[q] / var[q]); /* :425-427 (sic: W/var) */
    /* ess :496-546 */
    float *avg = (float *)calloc((size_t)(N * p), sizeof(float));
    {
        float *ac = (float *)malloc(sizeof(float) * N * p);
        for (int64_t j = 0; j < C; ++j) {
            if (N <= 100) orc_autocov_bf(split + j * N * p, N, p, ac);
            else orc_autocov_fft(split + j * N * p, N, p, ac);
            for (int64_t i = 0; i < N * p; ++i) avg[i] += ac[i];
        }
        for (int64_t i = 0; i < N * p; ++i) avg[i] = avg[i] / (float)C;
        free(ac);
    }
    for (int64_t q = 0; q < p; ++q) {
        /* rho_t = -(( -avg + within) / var) + 1 */
#define ORC_RHO(tt) (-((-avg[(tt) * p + q] + within[q]) / var[q]) + 1.0f)
        float mn = N >= 2 ? ORC_RHO(0) + ORC_RHO(1) : 0.0f;
        float o = 0.0f;
        for (int64_t t = 0; t + 1 < N; t += 2) {
            float pt = ORC_RHO(t) + ORC_RHO(t + 1);
            if (pt <= 0.0f) break;
            if (pt > mn) pt = mn;
            mn = pt;
            o += pt;
        }
#undef ORC_RHO
        float tau = -1.0f + 2.0f * o;
        ess_out[q] = (1.0f / tau) * (float)C * (float)N;
    }
    free(avg);
    free(within);
    free(split);
    return 0;
}

static int orc_cmp_desc(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (y > x) - (y < x);
}
/* basic_stats, src/stats.rs:310-336: sort descending; min = last, max = first, median = data[len/2],
 * mean, std(ddof = 1).  out5 = min, median, max, mean, std. */
ORC_API void orc_basic_stats(const float *data, int64_t n, float *out5) {
    float *d = (float *)malloc(sizeof(float) * n);
    memcpy(d, data, sizeof(float) * n);
    qsort(d, (size_t)n, sizeof(float), orc_cmp_desc);
    float sum = 0.0f;
    for (int64_t i = 0; i < n; ++i) sum += d[i];
    float mean = sum / (float)n;
    /* ndarray std(ddof): Welford in the element type */
    float m = 0.0f, s2 = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        float cnt = (float)(i + 1);
        float delta = d[i] - m;
        m += delta / cnt;
        s2 += delta * (d[i] - m);
    }
    out5[0] = d[n - 1]; out5[1] = d[n / 2]; out5[2] = d[0]; out5[3] = mean;
    out5[4] = sqrtf(s2 / ((float)n - 1.0f));
    free(d);
}

/* ------------------------------------------------------------------ Gibbs sampler (src/gibbs.rs:89-205)
 * GibbsMarkovChain::step sweeps the coordinates in order, state[i] = target.sample(i, &state) (:122-126).  The
 * conditional is the two-component Gaussian mixture of the reference's tests / examples/mixture_gibbs.rs:24-72
 * (state = [x, z]) or ConstantConditional (src/gibbs.rs:218-226).  In the reference every chain holds a CLONE of the
 * conditional including its SmallRng (GibbsSampler::new :165-176), so all chains consume the same noise stream;
 * `reference` mode reproduces that (one SmallRng(cond_seed) per chain) and can record the draws as tapes
 * normals[chains, steps] (z-scores) / unifs[chains, steps]; replay mode consumes such tapes. */
static inline double orc_mix_pdf(double x, double mu, double sigma) {
    const double var = sigma * sigma;
    const double coeff = 1.0 / sqrt(2.0 * 3.14159265358979323846 * var);
    const double d = x - mu;
    const double exp_val = exp(-(d * d) / (2.0 * var));
    return coeff * exp_val;
}

/* kind 1 = constant (p[0] = c), kind 2 = mixture (p = mu0, sigma0, mu1, sigma1, pi0) */
ORC_API int orc_gibbs_run(int kind, const double *p, double *state, int64_t chains, int D, int64_t n_collect, int64_t n_discard,
                          int reference, uint64_t cond_seed, double *normals, double *unifs, double *out) {
    const int64_t steps = n_collect + n_discard;
    if (kind == 2 && D != 2) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        orc_smallrng rng;
        orc_smallrng_seed(&rng, cond_seed);
        double *x = state + c * D;
        for (int64_t s = 0; s < steps; ++s) {
            if (kind == 1) {
                for (int i = 0; i < D; ++i) x[i] = p[0];
            } else {
                double z01, u;
                if (reference) {
                    z01 = orc_next_normal(&rng);
                    if (normals) normals[c * steps + s] = z01;
                } else {
                    z01 = normals[c * steps + s];
                }
                /* i = 0: x | z ~ Normal(mu_z, sigma_z) = mean + std * zscore (rand_distr 0.5 Normal::sample) */
                x[0] = (x[1] < 0.5) ? p[0] + p[1] * z01 : p[2] + p[3] * z01;
                if (reference) {
                    u = orc_next_f64(&rng);
                    if (unifs) unifs[c * steps + s] = u;
                } else {
                    u = unifs[c * steps + s];
                }
                /* i = 1: z | x */
                const double p0 = p[4] * orc_mix_pdf(x[0], p[0], p[1]);
                const double p1 = (1.0 - p[4]) * orc_mix_pdf(x[0], p[2], p[3]);
                const double total = p0 + p1;
                const double prob_z1 = total > 0.0 ? p1 / total : 0.5;
                x[1] = (u < prob_z1) ? 1.0 : 0.0;
            }
            if (s >= n_discard && out) {
                double *o = out + (c * n_collect + (s - n_discard)) * D;
                for (int i = 0; i < D; ++i) o[i] = x[i];
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ MH over Categorical (src/distributions.rs:422-477)
 * Target<usize, T>::unnorm_logp(position) = ln(probs[position[0]]) for position[0] < len, -inf beyond (:457-476), with
 * probs normalised by their left-fold sum in Categorical::new (:431-440); proposal = NonnegativeProposal
 * (examples/poisson_mh.rs:28-77); transition = MHMarkovChain::step (src/metropolis_hastings.rs:303-315). */
static inline double orc_cat_logp(const double *probs, double sum, int64_t K, uint64_t k) {
    return k < (uint64_t)K ? log(probs[k] / sum) : -INFINITY;
}
static inline uint64_t orc_mh_cat_step(const double *probs, double sum, int64_t K, uint64_t x, int flip, double u) {
    uint64_t y = (x == 0) ? 1 : (flip ? x + 1 : x - 1);
    double cur_lp = orc_cat_logp(probs, sum, K, x);
    double prop_lp = orc_cat_logp(probs, sum, K, y);
    double qf = orc_nonneg_logq(x, y);
    double qb = orc_nonneg_logq(y, x);
    double r = (prop_lp + qb) - (cur_lp + qf);
    return (r > log(u)) ? y : x;
}
static inline double orc_fold_sum(const double *probs, int64_t K) {
    double sum = 0.0;
    for (int64_t k = 0; k < K; ++k) sum = sum + probs[k];
    return sum;
}

ORC_API int orc_mh_categorical_run_replay(const double *probs, int64_t K, uint64_t *state, int64_t chains, int64_t n_collect,
                                          int64_t n_discard, const uint8_t *flip, const double *u, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
    const double sum = orc_fold_sum(probs, K);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        uint64_t x = state[c];
        for (int64_t i = 0; i < steps; ++i) {
            x = orc_mh_cat_step(probs, sum, K, x, flip[c * steps + i], u[c * steps + i]);
            if (i >= n_discard) out[c * n_collect + (i - n_discard)] = x;
        }
        state[c] = x;
    }
    return 0;
}

/* native Philox keying: the octet contract of the integer MH kernel (see orc_mh_poisson_run_philox) */
ORC_API int orc_mh_categorical_run_philox(const double *probs, int64_t K, uint64_t *state, int64_t chains, int64_t chain_offset,
                                          int64_t step_base, int64_t n_collect, int64_t n_discard, uint64_t seed, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const double sum = orc_fold_sum(probs, K);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        uint64_t x = state[c];
        const uint64_t gc = (uint64_t)(c + chain_offset);
        for (int64_t s = 0; s < steps; ++s) {
            const uint64_t gs = (uint64_t)(step_base + s);
            const uint32_t i = (uint32_t)(gs & 7);
            uint32_t ctr[4] = {(uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)(gs >> 3), 0u};
            uint32_t w[4], v[4];
            orc_philox4x32_10(key, ctr, w);
            ctr[3] = 1u + (i >> 1);
            orc_philox4x32_10(key, ctr, v);
            const uint32_t h = (w[i >> 1] >> (16u * (i & 1u))) & 0xffffu;
            const int flip = (int)(h >> 15);
            const uint64_t bits = (i & 1u) ? (((uint64_t)v[3] << 32) | v[2]) : (((uint64_t)v[1] << 32) | v[0]);
            const uint64_t u53 = ((uint64_t)(h & 0x7fffu) << 38) | (bits >> 26);
            const double u = (double)u53 * (1.0 / 9007199254740992.0);
            x = orc_mh_cat_step(probs, sum, K, x, flip, u);
            if (s >= n_discard) out[c * n_collect + (s - n_discard)] = x;
        }
        state[c] = x;
    }
    return 0;
}

/* ------------------------------------------------------------------ MH over a tabulated integer target
 * Any Target<i32, f64> given as logp[0..K) (-inf outside), with either NonnegativeProposal (examples/poisson_mh.rs:28-77)
 * or the symmetric +-1 walk clamped to [0, upper] whose logp is ln(0.5) both ways (PoissonRandomWalk: `new_state < 0 ->
 * 0`, BinomialRandomWalk: `.max(0).min(n)`, tests/metrohast_poisson_test.rs:63-80,193-207).  Transition =
 * MHMarkovChain::step (src/metropolis_hastings.rs:303-315).  upper < 0: no upper clamp. */
static inline double orc_tab_logp(const double *lp, int64_t K, int64_t k) { return (k >= 0 && k < K) ? lp[k] : -INFINITY; }
static inline int64_t orc_mh_tab_step(const double *lp, int64_t K, int reflect, int64_t upper, int64_t x, int flip, double u) {
    int64_t y;
    double qf, qb;
    if (reflect) {
        y = x + (flip ? 1 : -1);
        if (y < 0) y = 0;
        if (upper >= 0 && y > upper) y = upper;
        qf = qb = log(0.5);
    } else {
        y = (x == 0) ? 1 : (flip ? x + 1 : x - 1);
        qf = orc_nonneg_logq((uint64_t)x, (uint64_t)y);
        qb = orc_nonneg_logq((uint64_t)y, (uint64_t)x);
    }
    double r = (orc_tab_logp(lp, K, y) + qb) - (orc_tab_logp(lp, K, x) + qf);
    return (r > log(u)) ? y : x;
}
ORC_API int orc_mh_tabulated_run_replay(const double *lp, int64_t K, int reflect, int64_t upper, uint64_t *state, int64_t chains,
                                        int64_t n_collect, int64_t n_discard, const uint8_t *flip, const double *u, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        int64_t x = (int64_t)state[c];
        for (int64_t i = 0; i < steps; ++i) {
            x = orc_mh_tab_step(lp, K, reflect, upper, x, flip[c * steps + i], u[c * steps + i]);
            if (i >= n_discard) out[c * n_collect + (i - n_discard)] = (uint64_t)x;
        }
        state[c] = (uint64_t)x;
    }
    return 0;
}
/* native Philox keying: the octet contract of the integer MH kernel (see orc_mh_poisson_run_philox) */
ORC_API int orc_mh_tabulated_run_philox(const double *lp, int64_t K, int reflect, int64_t upper, uint64_t *state, int64_t chains,
                                        int64_t chain_offset, int64_t step_base, int64_t n_collect, int64_t n_discard,
                                        uint64_t seed, uint64_t *out) {
    const int64_t steps = n_collect + n_discard;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chains; ++c) {
        int64_t x = (int64_t)state[c];
        const uint64_t gc = (uint64_t)(c + chain_offset);
        for (int64_t s = 0; s < steps; ++s) {
            const uint64_t gs = (uint64_t)(step_base + s);
            const uint32_t i = (uint32_t)(gs & 7);
            uint32_t ctr[4] = {(uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)(gs >> 3), 0u};
            uint32_t w[4], v[4];
            orc_philox4x32_10(key, ctr, w);
            ctr[3] = 1u + (i >> 1);
            orc_philox4x32_10(key, ctr, v);
            const uint32_t h = (w[i >> 1] >> (16u * (i & 1u))) & 0xffffu;
            const int flip = (int)(h >> 15);
            const uint64_t bits = (i & 1u) ? (((uint64_t)v[3] << 32) | v[2]) : (((uint64_t)v[1] << 32) | v[0]);
            const uint64_t u53 = ((uint64_t)(h & 0x7fffu) << 38) | (bits >> 26);
            const double u = (double)u53 * (1.0 / 9007199254740992.0);
            x = orc_mh_tab_step(lp, K, reflect, upper, x, flip, u);
            if (s >= n_discard) out[c * n_collect + (s - n_discard)] = (uint64_t)x;
        }
        state[c] = (uint64_t)x;
    }
    return 0;
}
