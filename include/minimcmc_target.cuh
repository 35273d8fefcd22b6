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
#include "../mini_mcmc_b200/csrc/mmc_gibbs.cuh"
#include "../mini_mcmc_b200/csrc/mmc_hmc.cuh"
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
