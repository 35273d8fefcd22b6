// K3: HMC on a D-dimensional Gaussian with dense covariance (BASELINE config C4).
// The D-dim generalisation of DiffableGaussian2D::unnorm_logp_batch (src/distributions.rs:262-288) inside
// HMC::step / leapfrog (src/hmc.rs:304-431): per gradient evaluation Z = (X - mu) Sigma^-1 over all chains is one
// [chains x D] x [D x D] GEMM; logp = norm_const - 0.5 rowsum(Z o (X - mu)), grad = -Z (symmetric precision).
#pragma once

#include "mmc_common.cuh"

namespace mmc {

// device buffers of the dense-Gaussian HMC path
struct DenseState {
    int D = 0;                    // the target's dimension (positions, draws, replay tapes)
    int Dp = 0;                   // internal row pitch: D rounded up to the 256-column tile (zero padding: the extra columns
                                  // of Delta, the momenta and the precision stay exactly 0 through the trajectory)
    int64_t chains = 0;
    float norm_const = 0.f;
    float *d_mean = nullptr;      // [D]
    float *d_prec = nullptr;      // [D, D] row-major, symmetric
    float *d_delta[2] = {nullptr, nullptr};  // [chains, D] ping-pong (FP32 path)
    float *d_mom = nullptr;       // [chains, D]
    float *d_scal = nullptr;      // [6, chains]: ke_cur, quad_cur, ke_prop, quad_prop, u, (spare)
    // tensor-core path operands (hi/lo TF32 splits), see mmc_dense_tc.cu
    float *d_prec_split = nullptr;                  // [2, D, D]
    float *d_delta_split[2] = {nullptr, nullptr};   // ping-pong of [2 (hi, lo), chains, D]
    void *tc = nullptr;                             // tensor maps
    bool tc_pair = false;                           // CTA-pair (cta_group::2) kernel instead of the 1-CTA one
    // mixed split (gemm_path 3): TF32 hi.hi + one K-concatenated BF16 MMA for hi.lo + lo.hi; the full-precision Delta
    // stays in d_delta[], the TF32 operand in the hi half of d_delta_split[], the bf16 cross-term operand here
    uint32_t *d_prec_x = nullptr;                   // [D, 2 D] bf16: per 32-column k-block 32 x lo then 32 x hi
    uint32_t *d_delta_x[2] = {nullptr, nullptr};    // [chains, 2 D] bf16: per k-block 32 x hi then 32 x lo
    bool tc_mixed = false;
    bool tc_hw_trunc = false;                       // kind::tf32 reads the full-precision Delta (self-tested, mmc_dense_tc.cu)
    int tc_quad_clusters = 0;                       // resident 4-CTA clusters of the quad kernel (0: not available)
    int tc_chain_pairs = 0;                         // resident CTA pairs of the leapfrog-chain kernel (its grid must be co-resident)
    bool tc_chain_coop = true;                      // launch it cooperatively (co-residency guaranteed by the driver)
    int *d_ready = nullptr;                         // chain kernel: [L + 2][row blocks] completion counters
    size_t ready_bytes = 0;
};

enum { kModeFirst = 0, kModeMid = 1, kModeLast = 2 };

struct DenseRunArgs {
    float *positions;        // [chains, D] in/out
    float *out;              // [chains, n_collect, D] or nullptr
    const float *momenta;    // replay [steps, chains, D] or nullptr
    const float *u;          // replay [steps, chains]
    float *trace;            // optional [steps, chains, 4]
    unsigned long long *accept_count;
    int64_t chains, chain_offset, step_base, n_collect, n_discard;
    int64_t out_pitch;       // draws per chain row of `out` (>= n_collect)
    float eps;
    int n_leapfrog;
    uint64_t seed;
    int gemm_path;           // 0 = FP32 SIMT tiles, 1 / 2 = tcgen05 3xTF32 (1 CTA / CTA pairs), 3 = CTA pairs, TF32 + BF16 mixed split
};

int dense_create(DenseState **st, const mmc_target_desc *target, int64_t chains);
void dense_destroy(DenseState *st);
int dense_run(DenseState *st, const DenseRunArgs &a, cudaStream_t stream);
// native-mode draws (momenta [steps, chains, D], u [steps, chains]) as the dense path consumes them
int dense_export_tape(int64_t chains, int D, int64_t chain_offset, uint64_t seed, int64_t step_base, int64_t steps,
                      float *momenta, float *u, cudaStream_t stream);

}  // namespace mmc
