/*
 * minimcmc.h — C ABI of the B200-native batched MCMC engine (libminimcmc.so).
 *
 * This is the drop-in boundary for the sampler hot path of mini-mcmc v0.8.3.  The reference has no FFI
 * of its own (it is one Rust crate); the entry points below are exactly what a Rust `extern "C"` shim
 * for that path binds (INTEGRATION.md shows the binding).  Each group cites the reference interface it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every call returns int: 0 = ok, < 0 = mmc_status; mmc_last_error() gives a thread-local message.
 *   - the library owns opaque handles and their device state; the caller owns every in/out buffer.
 *     Entry points without suffix take HOST pointers (copies happen inside the call, on the handle's
 *     stream, and the call returns after the result is in the caller's buffer); *_dev entry points take
 *     DEVICE pointers plus a cudaStream_t (passed as void*) and are asynchronous.
 *   - a handle is NOT thread-safe (mirrors `&mut self`); distinct handles are independent.
 *   - sampler state persists inside the handle, so calling run() again continues the chains
 *     (same as the reference structs, SURVEY.md §5 "checkpoint / resume").
 *   - outputs are C-contiguous [chains, n_collect, dim], the layout ChainRunner::run / HMC::run /
 *     NUTS::run return (src/core.rs:176-186, src/hmc.rs:157, src/nuts.rs:169).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     MMC_ERR_NO_DEVICE.
 *
 * RNG contract (native mode, i.e. no replay tape): Philox4x32-10, key = 64-bit seed,
 *   counter = (global_chain_lo, global_chain_hi, step_word, sub).  `global_chain` = local chain index +
 *   the handle's chain offset, `step` counts transitions since the handle was seeded, so results do not
 *   depend on how chains are sharded over GPUs.
 *     Poisson MH : step_word = step >> 3; W = call(sub 0); i = step & 7; h = 16-bit field i of W
 *                  (h = (W[i >> 1] >> 16 (i & 1)) & 0xffff): flip = h >> 15, u15 = h & 0x7fff;
 *                  V = call(sub 1 + (i >> 1)): low38 = ((i & 1) ? V[3]:V[2] : V[1]:V[0]) >> 26;
 *                  u = (u15 << 38 | low38) * 2^-53.  (V only matters when the top 15 bits tie with the accept
 *                  threshold, so the kernel evaluates it lazily: one Philox call per eight transitions, still the
 *                  exact 53-bit test.)
 *     MH (f64)   : step_word = step; sub 0 words (0,1) -> accept uniform (53 bit); sub 1 + j ->
 *                  Box-Muller pair of proposal normals (2j, 2j+1), words (0,1) -> u1, (2,3) -> u2.
 *     HMC (f32)  : step_word = step; sub j -> momenta 4j..4j+3 (Box-Muller on words (0,1), (2,3));
 *                  sub 0x80000000 word 0 -> accept uniform (w >> 8) * 2^-24.
 *     NUTS       : step_word = step (0 = init_chain); sub j -> normals 4j..4j+3; sub 0x80000000 word 0 ->
 *                  Exp(1) = -ln(((w >> 8) + 0.5) * 2^-24); uniform #q of the step -> sub 0x80000001 + (q >> 1),
 *                  words 2(q&1), 2(q&1)+1 (f64 uniforms use 53 bits, T = f32 uniforms the top 24 of the high word).
 *   Replay mode substitutes caller-provided tapes for all of the above (parity testing).
 */
#ifndef MINIMCMC_H
#define MINIMCMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMC_VERSION 100

typedef enum {
    MMC_OK = 0,
    MMC_ERR_INVALID = -1,      /* bad argument */
    MMC_ERR_NO_DEVICE = -2,    /* no CUDA device / driver */
    MMC_ERR_CUDA = -3,         /* CUDA runtime error (message in mmc_last_error) */
    MMC_ERR_UNSUPPORTED = -4,  /* combination not compiled in */
    MMC_ERR_OVERFLOW = -5,     /* integer state left the range the device tables cover */
    MMC_ERR_NOMEM = -6
} mmc_status;

typedef enum { MMC_F32 = 0, MMC_F64 = 1, MMC_U64 = 2 } mmc_dtype;

/* Built-in targets (src/distributions.rs, examples/poisson_mh.rs).  Custom device targets get ids
 * >= MMC_T_CUSTOM_BASE from mmc_register_target (see minimcmc_target.cuh). */
typedef enum {
    MMC_T_GAUSSIAN2D = 1,      /* Gaussian2D           src/distributions.rs:159-206  params = mean0,mean1,c00,c01,c10,c11 */
    MMC_T_ISO_GAUSSIAN = 2,    /* IsotropicGaussian    src/distributions.rs:394-402  params = std                          */
    MMC_T_POISSON = 3,         /* PoissonTarget        examples/poisson_mh.rs:10-26  params = lambda                       */
    MMC_T_ROSENBROCK_ND = 4,   /* RosenbrockND         src/distributions.rs:527-547                                        */
    MMC_T_ROSENBROCK_2D = 5,   /* Rosenbrock2D         src/distributions.rs:491-524  params = a,b                          */
    MMC_T_DIFF_GAUSSIAN2D = 6, /* DiffableGaussian2D   src/distributions.rs:213-316  params = mean0,mean1,c00,c01,c10,c11 */
    MMC_T_DENSE_GAUSSIAN = 7,  /* D-dim dense Gaussian (config C4)  params = norm_const; vec = mean[D]; mat = precision[D,D] */
    MMC_T_STD_NORMAL = 8,      /* test target          src/nuts.rs:1024-1037                                               */
    MMC_T_CATEGORICAL = 9,     /* Categorical          src/distributions.rs:422-477  (mmc_mh_create_categorical)            */
    MMC_T_TABULATED = 10,      /* any Target<i32|usize, f64> given as a table of log-probabilities (mmc_mh_create_tabulated) */
    MMC_T_CUSTOM_BASE = 1000
} mmc_target_kind;

typedef struct {
    int32_t kind;       /* mmc_target_kind or a registered custom id */
    int32_t dim;
    double params[8];
    const float *vec;   /* HOST pointers, copied to the device at create time (may be NULL) */
    const float *mat;
} mmc_target_desc;

/* Proposals for Metropolis-Hastings (trait Proposal, src/distributions.rs:92-101). */
typedef enum {
    MMC_Q_ISO_GAUSSIAN = 1, /* IsotropicGaussian::sample/logp  src/distributions.rs:360-392  param = std */
    MMC_Q_NONNEG_RW = 2,    /* NonnegativeProposal             examples/poisson_mh.rs:28-77              */
    MMC_Q_REFLECT_RW = 3    /* symmetric +-1 walk clamped to the support (PoissonRandomWalk / BinomialRandomWalk,
                               tests/metrohast_poisson_test.rs:52-84,184-214); mmc_mh_create_tabulated only */
} mmc_proposal_kind;

typedef struct {
    int32_t kind;
    double param;
} mmc_proposal_desc;

/* ------------------------------------------------------------------ library */
int mmc_version(void);
const char *mmc_last_error(void);
/* Select the CUDA device this thread's subsequent handles live on (one process per GPU). */
int mmc_init(int device);
/* SM count / name of the active device (diagnostics). */
int mmc_device_info(int *sm_count, char *name, int name_len);

/* ------------------------------------------------------------------ init   (src/core.rs:394-435)
 * init_with_seed / init_det: n*d StandardNormal f64 draws from SmallRng::seed_from_u64(seed), row-major.
 * Host routine (bit-compatible with rand 0.9 / rand_distr 0.5 so existing seeds keep their meaning). */
int mmc_init_positions(double *out_host, int64_t n, int64_t d, uint64_t seed);
/* Device-side N(0,1) starts for large batches, Philox keyed (seed, global_chain, step = 0xFFFFFFFF). */
int mmc_init_positions_dev(float *out_dev, int64_t n, int64_t d, uint64_t seed, int64_t chain_offset, void *stream);

/* ------------------------------------------------------------------ Metropolis-Hastings
 * Replaces MetropolisHastings::new / .seed / ChainRunner::run for the built-in target+proposal pairs
 * (src/metropolis_hastings.rs:149-193,303-315; src/core.rs:55-73,176-186).
 * state dtype: MMC_F64 or MMC_F32 (continuous targets; MetropolisHastings<S, T, ..> is generic over the float type,
 * src/metropolis_hastings.rs:87 - with MMC_F32 every operation runs in f32, in/out arrays are f32, replay tapes stay f64
 * arrays holding f32 values) or MMC_U64 (`usize` state of the Poisson example).  IsotropicGaussian targets run for any
 * dim <= 256; custom targets / proposals come from mmc_register_mh_target. */
typedef struct mmc_mh mmc_mh;

typedef struct {
    /* all [chains, steps(, dim)] with steps = n_collect + n_discard; host pointers for mmc_mh_run,
     * device pointers for mmc_mh_run_dev */
    const double *noise;  /* continuous: StandardNormal draws z, proposal = (0 + std*z) + x */
    const double *u;      /* accept uniforms in [0,1) */
    const uint8_t *flip;  /* Poisson: 1 -> x+1, 0 -> x-1 (ignored when x == 0) */
    double *trace;        /* optional out [chains, steps, 4]: cur_lp, prop_lp, log_ratio, accepted (continuous only) */
} mmc_replay_mh;

int mmc_mh_create(mmc_mh **h, const mmc_target_desc *target, const mmc_proposal_desc *proposal,
                  const void *init_host, int64_t chains, int32_t dim, int32_t state_dtype);
/* MetropolisHastings over Categorical<f64> (src/distributions.rs:422-477, Target<usize, T>: logp(k) = ln(probs[k] / sum)
 * for k < n, -inf beyond) with the +-1 NonnegativeProposal of examples/poisson_mh.rs:28-77; u64 state [chains], dim 1.
 * Runs through the same table-driven integer kernel as the Poisson target (threshold accept mode). */
int mmc_mh_create_categorical(mmc_mh **out, const double *probs, int32_t n_categories, const void *init_host, int64_t chains);
/* MetropolisHastings over any integer-state target tabulated on [0, n_states): logp[k] = Target::unnorm_logp(&[k]) as the
 * caller's host code evaluates it, -inf beyond the table.  proposal_kind = MMC_Q_NONNEG_RW or MMC_Q_REFLECT_RW.  This is
 * how the `i32` Poisson / Binomial variants of tests/metrohast_poisson_test.rs:18-85,157-214 run (state is carried as u64).
 * The accept thresholds are bisected on the host from (lp' + q_b) - (lp + q_f) > ln(u), like the Poisson target. */
int mmc_mh_create_tabulated(mmc_mh **out, const double *logp, int32_t n_states, int32_t proposal_kind, const void *init_host,
                            int64_t chains);
int mmc_mh_seed(mmc_mh *h, uint64_t seed);                  /* .seed(s): also resets the step counter */
int mmc_mh_set_chain_offset(mmc_mh *h, int64_t offset);     /* first global chain id held by this handle */
/* Poisson accept test: 0 = evaluate (lp'+qb)-(lp+qf) > ln(u) in f64 on the device,
 * 1 (default) = compare the 53-bit uniform against host-built integer thresholds that encode the same
 * predicate with the host libm's ln (bit-exact with the reference's CPU decisions). */
int mmc_mh_set_accept_mode(mmc_mh *h, int32_t mode);
/* *_run_dev only: the caller's tensor has `pitch_steps` draws per chain row (0 = n_collect of the call), so a run can
 * fill the window [t0, t0 + n_collect) of a [chains, pitch_steps, dim] tensor when it is handed out_dev + t0 * dim.
 * This is how run_progress samples in blocks (progress statistics between blocks) without an extra copy. */
int mmc_mh_set_out_pitch(mmc_mh *h, int64_t pitch_steps);
int mmc_mh_run(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, const mmc_replay_mh *replay);
int mmc_mh_run_dev(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_dev, const mmc_replay_mh *replay_dev,
                   void *stream);
/* bytes that cross PCIe per collected draw in mmc_mh_run (8 for the plain u64 copy; 1 or 2 when the Poisson path
 * ships compact draws and widens them to u64 on the host threads) */
/* Opt-in compact return type for the integer targets (Poisson / Categorical, table accept mode): out_host receives
 * [chains, n_collect] draws as u8 (tables of <= 256 states) or u16, *elem_bytes says which.  1-2 B per draw cross PCIe and
 * are written to host memory instead of the 8 B `usize` of the reference API (mmc_mh_run stays the drop-in). */
int mmc_mh_run_compact(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, int32_t *elem_bytes);
/* STREAM-style write bandwidth of this host (GB/s; `threads` = 0: the size of the widening pool mmc_mh_run uses): the
 * roofline of the end-to-end Poisson path, whose u64 result array the host cores have to write. */
int mmc_host_write_bandwidth(uint64_t bytes, int32_t threads, int32_t reps, double *gb_per_s, int32_t *threads_used);
int mmc_mh_d2h_bytes_per_draw(mmc_mh *h);
int mmc_mh_get_state(mmc_mh *h, void *state_host);
int mmc_mh_set_state(mmc_mh *h, const void *state_host);
void mmc_mh_destroy(mmc_mh *h);

/* ------------------------------------------------------------------ HMC
 * Replaces HMC::new / set_seed / step / run (src/hmc.rs:87-158,304-431).  Tensors are f32 (burn backend
 * element type).  exact = 1 forbids FMA contraction so replayed trajectories reproduce the CPU
 * arithmetic operation by operation; exact = 0 (default) is the throughput build. */
typedef struct mmc_hmc mmc_hmc;

typedef struct {
    const float *momenta; /* [steps, chains, dim] */
    const float *u;       /* [steps, chains] accept uniforms in [0,1) */
    float *trace;         /* optional out [steps, chains, 4]: logp_current, logp_proposed, accept_logp, accepted */
} mmc_replay_hmc;

int mmc_hmc_create(mmc_hmc **h, const mmc_target_desc *target, const float *init_host, int64_t chains, int32_t dim,
                   double step_size, int32_t n_leapfrog);
int mmc_hmc_set_seed(mmc_hmc *h, uint64_t seed);
int mmc_hmc_set_chain_offset(mmc_hmc *h, int64_t offset);
int mmc_hmc_set_exact(mmc_hmc *h, int32_t exact);
/* dense Gaussian target only: gradient GEMM on 0 = FP32 SIMT tiles, 1 = tcgen05 tensor cores (3xTF32 split, fp32
 * accumulation in TMEM, one CTA per 128 x 256 tile), 2 = the same on CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles,
 * each SM stages half of the B tile), 3 = CTA pairs with the mixed split: TF32 hi.hi plus ONE K-concatenated BF16 MMA for
 * the two cross terms hi.lo + lo.hi (two thirds of the tensor work of the 3xTF32 split, same operand bytes).  Paths 1 and 2
 * give identical results.  exact = 1 always selects the FP32 path. */
int mmc_hmc_set_gemm_path(mmc_hmc *h, int32_t path);
int mmc_hmc_set_out_pitch(mmc_hmc *h, int64_t pitch_steps); /* see mmc_mh_set_out_pitch */
int mmc_hmc_step(mmc_hmc *h);
int mmc_hmc_run(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_host, const mmc_replay_hmc *replay);
int mmc_hmc_run_dev(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_dev,
                    const mmc_replay_hmc *replay_dev, void *stream);
int mmc_hmc_get_positions(mmc_hmc *h, float *positions_host);
int mmc_hmc_positions_dev(mmc_hmc *h, float **positions_dev); /* borrowed pointer [chains, dim] */
/* accepted transitions / total transitions since creation (device counters) */
int mmc_hmc_get_accept_counts(mmc_hmc *h, int64_t *accepted, int64_t *total);
/* dump the native-mode draws the sampler WOULD consume for steps [step_base, step_base+steps) so a CPU
 * checker can replay them: momenta [steps, chains, dim], u [steps, chains] (device pointers). */
int mmc_hmc_export_tape_dev(mmc_hmc *h, int64_t step_base, int64_t steps, float *momenta_dev, float *u_dev,
                            void *stream);
void mmc_hmc_destroy(mmc_hmc *h);

/* Custom device targets (see include/minimcmc_target.cuh): a user-compiled shared library registers a launcher for
 * its functor; the returned id (>= MMC_T_CUSTOM_BASE) is used as mmc_target_desc.kind in mmc_hmc_create.
 * Replaces the role of the BatchedGradientTarget trait bound of HMC (src/distributions.rs:65-76, src/hmc.rs:36-57). */
typedef int (*mmc_hmc_launch_fn)(const void *hmc_params, int replay, int exact, const double *target_params, void *stream);
int mmc_register_hmc_target(const char *name, int32_t dim, mmc_hmc_launch_fn fn);
/* The same for NUTS (any GradientTarget in NUTS::new, src/nuts.rs:123-129, trait src/distributions.rs:78-88) and for
 * Metropolis-Hastings (any Target, optionally with its own Proposal, in MetropolisHastings::new,
 * src/metropolis_hastings.rs:149-159, traits src/distributions.rs:92-108).  A name registered for several samplers keeps
 * ONE kind id; MMC_REGISTER_NUTS_TARGET / MMC_REGISTER_MH_TARGET / MMC_REGISTER_MH_PAIR in minimcmc_target.cuh generate
 * the launchers.  query_only != 0: only report the persistent grid and the scratch floats the launch needs. */
typedef int (*mmc_nuts_launch_fn)(const void *nuts_params, int scalar_f64, int replay, int exact, const double *target_params,
                                  int sm_count, int64_t *grid, size_t *scratch_floats, int query_only, void *stream);
int mmc_register_nuts_target(const char *name, int32_t dim, mmc_nuts_launch_fn fn);
typedef int (*mmc_mh_launch_fn)(const void *mh_params, int replay, const double *target_params, double proposal_param,
                                void *stream);
int mmc_register_mh_target(const char *name, int32_t dim, mmc_mh_launch_fn fn);
int mmc_lookup_target(const char *name); /* kind id, or MMC_ERR_INVALID when unknown */

/* ------------------------------------------------------------------ NUTS
 * Replaces NUTS::new / set_seed / run / run_progress and NUTSChain (src/nuts.rs:123-170,347-353,410-691).
 * scalar_dtype = type T of epsilon / joint / logu / alpha (MMC_F64 in the reference's golden tests,
 * MMC_F32 in examples/minimal_nuts.rs).  The reference's tree has no depth cap; max_depth bounds the
 * number of doublings (default 10 when <= 0). */
typedef struct mmc_nuts mmc_nuts;

typedef struct {
    /* per chain tapes; all host (mmc_nuts_run) or all device (mmc_nuts_run_dev) */
    const double *normals; int64_t cap_normals; /* [chains, cap_normals]: D for init_chain then D per step */
    const double *exps;    int64_t cap_exps;    /* [chains, cap_exps]   : one Exp(1) per step */
    const double *unifs;   int64_t cap_unifs;   /* [chains, cap_unifs]  : sequential uniform tape */
} mmc_replay_nuts;

int mmc_nuts_create(mmc_nuts **h, const mmc_target_desc *target, const float *init_host, int64_t chains,
                    int32_t dim, double target_accept_p, int32_t scalar_dtype, int32_t max_depth);
int mmc_nuts_set_seed(mmc_nuts *h, uint64_t seed);
int mmc_nuts_set_chain_offset(mmc_nuts *h, int64_t offset);
int mmc_nuts_set_exact(mmc_nuts *h, int32_t exact);
int mmc_nuts_set_out_pitch(mmc_nuts *h, int64_t pitch_steps); /* see mmc_mh_set_out_pitch */
/* Lane layout of the tree kernel.  0 (default) = automatic: several chains per warp (G lanes per chain, 32 / G chains
 * advancing in lock step and sharing the tree bookkeeping) wherever that layout is compiled in for the target
 * (RosenbrockND D <= 128: G = 4 / 8 / 16, StdNormal D <= 32, the 2-D targets: G = 4), else one chain per warp;
 * 32 = always one chain per warp; any other value must be the G compiled in for the target.  Both layouts implement
 * the same algorithm on the same Philox counters and replay tapes; f32 sums are grouped differently, so native draws
 * agree between layouts only up to fp32 rounding (amplified by the dynamics).  mmc_nuts_get_layout reports the lanes
 * per chain of the last launch. */
int mmc_nuts_set_layout(mmc_nuts *h, int32_t lanes_per_chain);
int mmc_nuts_get_layout(mmc_nuts *h, int32_t *lanes_per_chain);
/* Load balancing of the several-chains-per-warp kernel: a native run is dispensed to the persistent warps in slices of
 * slice_steps transitions per group of chains (state handed over through global memory), which evens out the tail
 * when a shard holds only a few waves of chains.  -1 (default) = a sixteenth of the launch's iterations (at least 8), 0 = whole runs.  The draws do
 * not depend on the slicing. */
int mmc_nuts_set_slicing(mmc_nuts *h, int64_t slice_steps);
/* Lock-step efficiency of the several-chains-per-warp kernel (opt-in experiment): a warp is as slow as its deepest tree,
 * and tree depth follows the chain's adapted step size.  With mode = 1 a native run is cut into phases (iterations 32, 96,
 * 224 and the end of the burn-in) and, between them, a device-side counting sort by log2(step size) re-forms the warps
 * from chains of similar step size.  0 and -1 (default) = off: on the C5 benchmark the measured effect is within noise
 * (DESIGN.md K4b).  The draws do not depend on the grouping. */
int mmc_nuts_set_regroup(mmc_nuts *h, int32_t mode);
/* Splitting one run over several launches (run_progress in blocks): adapt_until = absolute step count m up to which
 * dual averaging adapts (-1 = the reference's rule `m <= n_discard` of each call, src/nuts.rs:681); resume = 1 makes
 * the following runs continue the chains without init_chain (src/nuts.rs:528-545), so that the blocks reproduce the
 * single-launch run exactly. */
int mmc_nuts_set_continuation(mmc_nuts *h, int64_t adapt_until, int32_t resume);
/* progress_semantics 0: NUTS::run (n_collect+n_discard-1 steps, slot 0 = starting position);
 *                    1: NUTS::run_progress (n_collect+n_discard steps). */
int mmc_nuts_run(mmc_nuts *h, int64_t n_collect, int64_t n_discard, int32_t progress_semantics, float *out_host,
                 const mmc_replay_nuts *replay);
int mmc_nuts_run_dev(mmc_nuts *h, int64_t n_collect, int64_t n_discard, int32_t progress_semantics, float *out_dev,
                     const mmc_replay_nuts *replay_dev, void *stream);
/* state_host [chains, 5] = epsilon, epsilon_bar, h_bar, mu, m (the pub fields of NUTSChain, src/nuts.rs:361-390) */
int mmc_nuts_get_state(mmc_nuts *h, double *state_host);
int mmc_nuts_set_state(mmc_nuts *h, const double *state_host);
int mmc_nuts_get_positions(mmc_nuts *h, float *positions_host);
int mmc_nuts_set_positions(mmc_nuts *h, const float *positions_host);   /* NUTSChain.position, src/nuts.rs:363 */
/* Per-transition trace of the following runs (debug / parity): trace_dev [chains, pitch_steps, 8] f64, caller-owned
 * device memory, row = iteration of the launch; columns joint_0, logu, n, alpha, n_alpha, tree depth, epsilon used,
 * uniforms consumed (src/nuts.rs:550-691).  nullptr switches it off. */
int mmc_nuts_set_trace_dev(mmc_nuts *h, double *trace_dev, int64_t pitch_steps);
/* Known-answer entry for build_tree (src/nuts.rs:764-946; the reference's test_build_tree, :1057-1121): ONE doubling of
 * depth j per chain from (the handle's positions, mom_host, grad_host) [chains, dim], scal_host [chains, 4] = logu, v
 * (+1 / -1), epsilon, joint_0, and the tree's uniforms from unifs_host [chains, cap_unifs], on the kernel, arithmetic
 * policy and lane layout the handle is set to.  out_vec_host [chains, 5, dim] = the new edge (position, momentum,
 * gradient), position' and gradient(position'); out_scal_host [chains, 6] = logp(position'), n', s', alpha', n_alpha',
 * uniforms consumed.  The opposite edge of the 13-tuple is the input, unchanged.  All chains share j. */
int mmc_nuts_build_tree(mmc_nuts *h, const float *mom_host, const float *grad_host, const double *scal_host, int32_t j,
                        const double *unifs_host, int64_t cap_unifs, float *out_vec_host, double *out_scal_host);
/* counters summed over chains since creation: grad evals, transitions, tapes consumed (uniforms) and the
 * histogram of tree depths [max_depth + 1] */
/* Debug entry: the tree-merge test "k 2^-53 < num / den" (src/nuts.rs:910-911 for a native 53-bit draw k) evaluated on
 * the device for n triples (host arrays); out_host[i] = 0 / 1. */
int mmc_debug_nuts_merge_test(const uint64_t *k53_host, const uint32_t *num_host, const uint32_t *den_host, int64_t n,
                              uint8_t *out_host);
int mmc_nuts_get_counters(mmc_nuts *h, int64_t *n_grad, int64_t *n_transitions, int64_t *depth_hist, int32_t hist_len);
void mmc_nuts_destroy(mmc_nuts *h);

/* ------------------------------------------------------------------ diagnostics
 * Replaces split_rhat_mean_ess / RunStats::from / basic_stats (src/stats.rs:310-336,396-554).
 * sample is [c, n, p] f32. */
int mmc_split_rhat_ess(const float *sample_host, int64_t c, int64_t n, int64_t p, float *rhat_host, float *ess_host);
int mmc_split_rhat_ess_dev(const float *sample_dev, int64_t c, int64_t n, int64_t p, float *rhat_host,
                           float *ess_host, void *stream);
/* Sharded form, one call (one process per GPU; src/stats.rs:416-423 over chains that live on several GPUs): `sample_dev`
 * holds this rank's c_local chains.  The per-parameter moment sums and summed autocovariances of a window of lags are
 * all-reduced with NCCL (one ncclAllReduce per window, the first fused with the chain count), the Geyer truncation is
 * checked on the device, and a wider window is computed only if some parameter has not terminated; every rank returns
 * the same rhat / ess.  The communicator is created from a 128-byte id that rank 0 obtains with mmc_comm_unique_id and
 * hands to the other ranks by any out-of-band channel (mmc_comm_create calls ncclCommInitRank on the calling thread's
 * current device), or adopted from an existing ncclComm_t (mmc_comm_wrap; the caller keeps ownership).  NCCL is
 * resolved at run time (dlopen of the libnccl.so.2 already in the process, else the system one, else MMC_NCCL_LIB). */
typedef struct mmc_comm mmc_comm;
int mmc_comm_unique_id(unsigned char *id128);
int mmc_comm_create(mmc_comm **out, const unsigned char *id128, int32_t nranks, int32_t rank);
int mmc_comm_wrap(mmc_comm **out, void *nccl_comm);
int mmc_comm_info(mmc_comm *c, int32_t *nranks, int32_t *rank, int32_t *nccl_version);
void mmc_comm_destroy(mmc_comm *c);
int mmc_split_rhat_ess_sharded(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, mmc_comm *comm, void *stream,
                               float *rhat_host, float *ess_host);
/* Building blocks of the sharded form.  `partial` is f64 [2 + n/2][p] in device memory
 * (mmc_stats_partial_len values): row 0 = sum_j m_j, row 1 = sum_j m_j^2, row 2 + t = sum_j acov_j(t) over the
 * LOCAL split chains.  Each call computes lags [lag0, lag0 + n_lags) (and rows 0-1 when lag0 == 0), zeroing
 * the rows it produces first.  The caller sums the partials across ranks (ncclAllReduce /
 * torch.distributed.all_reduce) and finalises on the host; mmc_stats_finalize returns 1 (not an error) when
 * the Geyer truncation has not terminated within `lags_available` lags, in which case further lags are
 * requested with another mmc_stats_partial_dev call. */
int64_t mmc_stats_partial_len(int64_t n, int64_t p);
int mmc_stats_partial_dev(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, int64_t lag0,
                          int64_t n_lags, double *partial_dev, void *stream);
int mmc_stats_finalize(const double *partial_host, int64_t c_total, int64_t n, int64_t p, int64_t lags_available,
                       float *rhat_host, float *ess_host);
typedef struct { float min, median, max, mean, std; } mmc_basic_stats;
typedef struct { mmc_basic_stats ess, rhat; } mmc_run_stats;
int mmc_basic_stats_of(const float *data_host, int64_t len, mmc_basic_stats *out);

/* ------------------------------------------------------------------ progress trackers (run_progress)
 * Replaces MultiChainTracker (src/stats.rs:189-307; HMC::run_progress src/hmc.rs:242-281) and ChainTracker +
 * collect_rhat (src/stats.rs:26-178; ChainRunner::run_progress src/core.rs:90-136,229-324).  The reference copies
 * every step's state to the host; here the tracker folds blocks of draws that are already in HBM:
 * mmc_tracker_steps_dev(sample [chains, n_total, dim], steps [t0, t0 + n_steps)) applies `step` once per draw, in
 * order, with the reference's f32 recurrences.  Rhat comes from f64 partial sums over the local chains
 * ([sum mean | sum mean^2 | sum sm2] per parameter, then sum of p_accept weights, then chains) that can be summed over
 * ranks before mmc_tracker_finalize.  mmc_tracker_summary = partial + finalize for one process; p_accept is
 * MultiChainTracker::p_accept (MULTI) or the mean over all chains of ChainStats::p_accept (PER_CHAIN; the
 * reference's progress bar averages the <= 5 chains it displays). */
typedef struct mmc_tracker mmc_tracker;
#define MMC_TRACK_MULTI 0
#define MMC_TRACK_PER_CHAIN 1
int mmc_tracker_create(mmc_tracker **out, int64_t chains, int32_t dim, int32_t flavor);
/* ChainTracker::new(initial_state): [chains, dim] of dtype (mmc_dtype); MULTI starts from zeros like the reference */
int mmc_tracker_set_initial_dev(mmc_tracker *t, const void *state_dev, int32_t dtype, void *stream);
int mmc_tracker_steps_dev(mmc_tracker *t, const void *sample_dev, int32_t dtype, int64_t n_total, int64_t t0,
                          int64_t n_steps, void *stream);
int64_t mmc_tracker_partial_len(int32_t dim);
int mmc_tracker_partial_dev(mmc_tracker *t, double *partial_dev, void *stream);
int mmc_tracker_finalize(const double *partial_host, int64_t chains_total, int32_t dim, uint64_t n_steps, int32_t flavor,
                         float *rhat_host, float *max_rhat);
int mmc_tracker_summary(mmc_tracker *t, float *rhat_host, float *max_rhat, float *p_accept, uint64_t *n_steps);
/* raw state for parity tests: mean / mean_sq [chains, dim]; p_accept [chains] (PER_CHAIN) or [1] (MULTI) */
int mmc_tracker_get(mmc_tracker *t, float *mean_host, float *mean_sq_host, float *p_accept_host);
void mmc_tracker_destroy(mmc_tracker *t);

/* ------------------------------------------------------------------ run_progress
 * Replaces ChainRunner::run_progress (MetropolisHastings / GibbsSampler, src/core.rs:208-360), HMC::run_progress
 * (src/hmc.rs:222-294) and NUTS::run_progress (src/nuts.rs:194-338): same sample as run(), plus RunStats, plus live
 * statistics.  The sampler runs in blocks of `block` steps (<= 0: about 1/16 of the run, multiples of 32) that write into
 * windows of one device tensor; after every block the device tracker folds the new draws and `cb(done, total, p_accept,
 * max_rhat, user)` is called (the numbers the reference's progress bars print, `p(accept)≈{:.2} max(rhat)≈{:.2}`).  MH,
 * Gibbs and NUTS track every step, burn-in included, with one ChainTracker per chain + collect_rhat (total = n_collect +
 * n_discard); HMC tracks the post-burn-in positions and the collected draws with a MultiChainTracker (total = n_collect).
 * cb and stats may be NULL.  The draws equal those of the corresponding run() call. */
typedef void (*mmc_progress_fn)(int64_t done, int64_t total, float p_accept, float max_rhat, void *user);
int mmc_mh_run_progress(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, int64_t block, mmc_progress_fn cb,
                        void *user, mmc_run_stats *stats);
int mmc_hmc_run_progress(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_host, int64_t block, mmc_progress_fn cb,
                         void *user, mmc_run_stats *stats);
int mmc_nuts_run_progress(mmc_nuts *h, int64_t n_collect, int64_t n_discard, float *out_host, int64_t block, mmc_progress_fn cb,
                          void *user, mmc_run_stats *stats);

/* ------------------------------------------------------------------ Gibbs
 * Replaces GibbsSampler::new / set_seed / run(_progress) and GibbsMarkovChain::step (src/gibbs.rs:89-205): each step
 * sweeps the coordinates in order, state[i] = conditional.sample(i, state).  `Conditional` is user code in the
 * reference (src/distributions.rs:485-487); the built-ins are the ones its tests and examples define:
 * MMC_G_CONSTANT (src/gibbs.rs:218-226, params = c) and MMC_G_MIXTURE2, the two-component Gaussian mixture with state
 * [x, z] (src/gibbs.rs:228-275, examples/mixture_gibbs.rs:24-72; params = mu0, sigma0, mu1, sigma1, pi0).
 * Native RNG: Philox counter (global chain, step, sub = coordinate): sub 0 -> Box-Muller z-score for x, sub 1 words
 * (0,1) -> 53-bit uniform for z.  Replay tapes feed the reference's own draws: normals / unifs [chains, steps];
 * trace [chains, steps, 2] receives the (z-score, uniform) every sweep consumed. */
typedef struct mmc_gibbs mmc_gibbs;
#define MMC_G_CONSTANT 1
#define MMC_G_MIXTURE2 2
#define MMC_G_CUSTOM_BASE 1000 /* kinds returned by mmc_register_gibbs_conditional */
typedef struct { int32_t kind; int32_t reserved; double params[8]; } mmc_conditional_desc;
typedef struct { const double *normals; const double *unifs; double *trace; } mmc_replay_gibbs;
int mmc_gibbs_create(mmc_gibbs **out, const mmc_conditional_desc *cond, const double *init_host, int64_t chains, int32_t dim);
int mmc_gibbs_set_seed(mmc_gibbs *h, uint64_t seed);
int mmc_gibbs_set_chain_offset(mmc_gibbs *h, int64_t offset);
int mmc_gibbs_set_out_pitch(mmc_gibbs *h, int64_t pitch_steps); /* see mmc_mh_set_out_pitch */
int mmc_gibbs_run(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_host, const mmc_replay_gibbs *replay);
int mmc_gibbs_run_dev(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_dev, const mmc_replay_gibbs *replay_dev,
                      void *stream);
int mmc_gibbs_run_progress(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_host, int64_t block,
                           mmc_progress_fn cb, void *user, mmc_run_stats *stats); /* see "run_progress" above */
int mmc_gibbs_get_state(mmc_gibbs *h, double *state_host);
/* Custom conditionals (the reference's `Conditional<S>` is user code, src/distributions.rs:485-487): a device functor
 * compiled by the user (include/minimcmc_target.cuh, MMC_REGISTER_GIBBS_CONDITIONAL) registers its launcher here and
 * gets the kind to put into mmc_conditional_desc.kind; params[8] are handed to the functor's constructor. */
typedef int (*mmc_gibbs_launch_fn)(const void *gibbs_params, const double *cond_params, void *stream);
int mmc_register_gibbs_conditional(const char *name, int32_t dim, mmc_gibbs_launch_fn fn);
void mmc_gibbs_destroy(mmc_gibbs *h);

/* ------------------------------------------------------------------ sample sinks
 * Replaces the data movement of save_arrow / save_parquet / save_parquet_tensor (src/io/arrow.rs:53-117,
 * src/io/parquet.rs:49-221) and all of save_csv / save_csv_tensor (src/io/csv.rs:47-147).
 * mmc_sink_columns_dev turns chains [c0, c0 + c_count) of a device sample [chains, n, dim] (mmc_dtype) into the f64
 * columns the Arrow schema holds: dim_cols_dev[d * rows + r] = (double) sample[c0 + r / n][r % n][d], rows = c_count * n.
 * The host wraps the copied columns as Arrow buffers (the file encoders stay library code, like the arrow / parquet
 * crates in the reference).  mmc_save_csv writes the whole file natively: header `chain,observation,dim_0..`, records
 * in chain-major order, numbers printed like Rust's Display (shortest round-trip digits, no exponent). */
int mmc_sink_columns_dev(const void *sample_dev, int32_t dtype, int64_t chains, int64_t n, int32_t dim, int64_t c0,
                         int64_t c_count, double *dim_cols_dev, void *stream);
int mmc_save_csv(const void *sample_host, int32_t dtype, int64_t chains, int64_t n, int32_t dim, const char *filename);

#ifdef __cplusplus
}
#endif
#endif /* MINIMCMC_H */
