// minimcmc_target.cuh — registration hook for CUSTOM device targets.
//
// The reference lets users plug any `BatchedGradientTarget` (src/distributions.rs:65-76) into HMC through Rust
// generics + burn autodiff.  CUDA cannot call host closures, so a custom target is a device functor compiled by
// the user with nvcc into its own shared library, which registers a launcher with libminimcmc at load time:
//
//     template <class A>                       // A = mmc::Fast (FMA contraction) or mmc::Exact (bit-faithful)
//     struct MyTarget {
//         static constexpr int kDim = 3;
//         float a, b;                          // parameters, filled from mmc_target_desc.params
//         __host__ explicit MyTarget(const double *p) : a((float)p[0]), b((float)p[1]) {}
//         __device__ float logp_grad(const float (&x)[kDim], float (&g)[kDim]) const { ... return logp; }
//     };
//     MMC_REGISTER_HMC_TARGET(my_target, MyTarget)
//
//   nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a my_target.cu \
//        -I<repo>/include -L<repo>/mini_mcmc_b200 -l:libminimcmc.so -o libmy_target.so
//
// Loading the library and calling `my_target_register()` returns the target kind (>= MMC_T_CUSTOM_BASE) to put
// into mmc_target_desc.kind for mmc_hmc_create.  The fused register-resident kernel (hmc_run_kernel) is
// instantiated for the functor, so a custom target runs at the same speed as a built-in one.
#pragma once

//
// Gibbs conditionals (the reference's `Conditional<S>`, src/distributions.rs:485-487) register the same way:
//
//     struct MyConditional {
//         static constexpr int kDim = 2;
//         double rho;
//         __host__ explicit MyConditional(const double *p) : rho(p[0]) {}
//         // new value of coordinate i given the whole state; draws come from rng.normal() / rng.uniform()
//         __device__ double sample(int i, const double (&state)[kDim], mmc::GibbsRng &rng) const { ... }
//     };
//     MMC_REGISTER_GIBBS_CONDITIONAL(my_conditional, MyConditional)
//
// `my_conditional_register()` returns the kind (>= MMC_G_CUSTOM_BASE) for mmc_conditional_desc.kind.
//
// NUTS takes the same thread-form functor (the reference's `GradientTarget`, src/distributions.rs:78-88, in NUTS::new,
// src/nuts.rs:123-129): MMC_REGISTER_NUTS_TARGET(my_target, MyTarget) instantiates the one-chain-per-warp tree kernel
// (nuts_run_kernel; every lane evaluates the functor on the gathered vector, kDim <= 128) for both arithmetic policies,
// both scalar types and native / replay draws, and defines `my_target_register_nuts()`.  A name registered for several
// samplers keeps one kind id.
//
// Metropolis-Hastings (the reference's `Target<T, F>` / `Proposal<T, F>`, src/distributions.rs:92-108, in
// MetropolisHastings::new, src/metropolis_hastings.rs:149-159) takes f64 functors:
//
//     struct MyMhTarget {
//         static constexpr int kDim = 2;
//         __host__ explicit MyMhTarget(const double *p);
//         __device__ double unnorm_logp(const double (&x)[kDim]) const;
//     };
//     MMC_REGISTER_MH_TARGET(my_mh_target, MyMhTarget)       // with the built-in IsotropicGaussian proposal
//
//     struct MyProposal {                                    // optional: a custom proposal
//         __host__ explicit MyProposal(double param);
//         // z: kDim standard normals of this step (Philox / replay tape); the reference's `sample(&mut self, current)`
//         template <int D> __device__ void sample(const double (&cur)[D], const double (&z)[D], double (&out)[D]) const;
//         template <int D> __device__ double logp(const double (&from)[D], const double (&to)[D]) const;
//     };
//     MMC_REGISTER_MH_PAIR(my_pair, MyMhTarget, MyProposal)
//
// Both define `<name>_register_mh()`.
#include "../mini_mcmc_b200/csrc/mmc_gibbs.cuh"
#include "../mini_mcmc_b200/csrc/mmc_hmc.cuh"
#include "../mini_mcmc_b200/csrc/mmc_mh.cuh"
#include "../mini_mcmc_b200/csrc/mmc_nuts_inst.cuh"
#include "../mini_mcmc_b200/csrc/mmc_targets.cuh"

#define MMC_REGISTER_HMC_TARGET(NAME, FUNCTOR)                                                                      \
    static int NAME##_mmc_launch(const void *pv, int replay, int exact, const double *tp, void *stream) {           \
        const mmc::HmcParams &p = *static_cast<const mmc::HmcParams *>(pv);                                         \
        if (exact)                                                                                                   \
            return mmc::launch_hmc<FUNCTOR<mmc::Exact>, mmc::Exact>(FUNCTOR<mmc::Exact>(tp), p, replay != 0,        \
                                                                    static_cast<cudaStream_t>(stream));             \
        return mmc::launch_hmc<FUNCTOR<mmc::Fast>, mmc::Fast>(FUNCTOR<mmc::Fast>(tp), p, replay != 0,               \
                                                              static_cast<cudaStream_t>(stream));                   \
    }                                                                                                                \
    extern "C" int NAME##_register(void) {                                                                          \
        return mmc_register_hmc_target(#NAME, FUNCTOR<mmc::Fast>::kDim, NAME##_mmc_launch);                         \
    }

#define MMC_REGISTER_GIBBS_CONDITIONAL(NAME, FUNCTOR)                                                               \
    static int NAME##_mmc_gibbs_launch(const void *pv, const double *cp, void *stream) {                            \
        const mmc::GibbsParams &p = *static_cast<const mmc::GibbsParams *>(pv);                                     \
        return mmc::launch_gibbs_generic<FUNCTOR>(FUNCTOR(cp), p, static_cast<cudaStream_t>(stream));               \
    }                                                                                                                \
    extern "C" int NAME##_register(void) {                                                                          \
        return mmc_register_gibbs_conditional(#NAME, FUNCTOR::kDim, NAME##_mmc_gibbs_launch);                       \
    }

#define MMC_REGISTER_NUTS_TARGET(NAME, FUNCTOR)                                                                     \
    template <class A, class ST, bool kReplay>                                                                      \
    static int NAME##_mmc_nuts_one(const mmc::NutsParams &p, const double *tp, int sms, int64_t *grid, size_t *scratch, \
                                   bool query, cudaStream_t stream) {                                               \
        constexpr int E = (FUNCTOR<A>::kDim + 31) / 32;                                                             \
        mmc::WSmall<FUNCTOR<A>, E> w{FUNCTOR<A>(tp)};                                                               \
        return mmc::nuts_launch_one<mmc::WSmall<FUNCTOR<A>, E>, A, ST, E, kReplay>(w, p, sms, grid, scratch, query, stream); \
    }                                                                                                                \
    template <class A>                                                                                               \
    static int NAME##_mmc_nuts_policy(const mmc::NutsParams &p, int f64, int replay, const double *tp, int sms,     \
                                      int64_t *grid, size_t *scratch, bool query, cudaStream_t s) {                 \
        if (f64) return replay ? NAME##_mmc_nuts_one<A, double, true>(p, tp, sms, grid, scratch, query, s)          \
                               : NAME##_mmc_nuts_one<A, double, false>(p, tp, sms, grid, scratch, query, s);        \
        return replay ? NAME##_mmc_nuts_one<A, float, true>(p, tp, sms, grid, scratch, query, s)                    \
                      : NAME##_mmc_nuts_one<A, float, false>(p, tp, sms, grid, scratch, query, s);                  \
    }                                                                                                                \
    static int NAME##_mmc_nuts_launch(const void *pv, int f64, int replay, int exact, const double *tp, int sms,    \
                                      int64_t *grid, size_t *scratch, int query, void *stream) {                    \
        const mmc::NutsParams &p = *static_cast<const mmc::NutsParams *>(pv);                                       \
        cudaStream_t s = static_cast<cudaStream_t>(stream);                                                          \
        return exact ? NAME##_mmc_nuts_policy<mmc::Exact>(p, f64, replay, tp, sms, grid, scratch, query != 0, s)    \
                     : NAME##_mmc_nuts_policy<mmc::Fast>(p, f64, replay, tp, sms, grid, scratch, query != 0, s);    \
    }                                                                                                                \
    extern "C" int NAME##_register_nuts(void) {                                                                     \
        return mmc_register_nuts_target(#NAME, FUNCTOR<mmc::Fast>::kDim, NAME##_mmc_nuts_launch);                   \
    }

#define MMC_REGISTER_MH_PAIR(NAME, TARGET, PROPOSAL)                                                                \
    static int NAME##_mmc_mh_launch(const void *pv, int replay, const double *tp, double qparam, void *stream) {     \
        const mmc::MhContParams &p = *static_cast<const mmc::MhContParams *>(pv);                                   \
        return mmc::launch_mh_functor<TARGET, PROPOSAL, double, TARGET::kDim>(TARGET(tp), PROPOSAL(qparam), p, replay != 0, \
                                                                              static_cast<cudaStream_t>(stream));   \
    }                                                                                                                \
    extern "C" int NAME##_register_mh(void) { return mmc_register_mh_target(#NAME, TARGET::kDim, NAME##_mmc_mh_launch); }

#define MMC_REGISTER_MH_TARGET(NAME, TARGET) MMC_REGISTER_MH_PAIR(NAME, TARGET, mmc::IsoProposalF<double>)
