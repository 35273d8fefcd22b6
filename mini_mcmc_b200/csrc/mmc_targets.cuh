// Analytic log-density + gradient device functors for the built-in targets (thread-per-chain form:
// the whole state lives in registers).  Each functor cites the reference lines it reproduces; the
// operation ORDER matters in the Exact policy (bit-for-bit replay against the CPU arithmetic).
//
// A custom target is any struct with
//     static constexpr int kDim;
//     __device__ float logp_grad(const float (&x)[kDim], float (&g)[kDim]) const;
// registered through minimcmc_target.cuh.
#pragma once

#include "mmc_common.cuh"

namespace mmc {

// c - a*b
template <class A> __device__ __forceinline__ float cms(float c, float a, float b);
template <> __device__ __forceinline__ float cms<Fast>(float c, float a, float b) { return fmaf(-a, b, c); }
template <> __device__ __forceinline__ float cms<Exact>(float c, float a, float b) { return __fsub_rn(c, __fmul_rn(a, b)); }

// RosenbrockND::unnorm_logp_batch, src/distributions.rs:531-547 (= examples/rosenbrock3d_hmc.rs:22-42):
//   logp = -sum_i [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2];  analytic gradient.
template <class A, int D>
struct RosenbrockND {
    static constexpr int kDim = D;
    __device__ __forceinline__ float logp_grad(const float (&x)[D], float (&g)[D]) const {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) g[i] = 0.0f;
#pragma unroll
        for (int i = 0; i + 1 < D; ++i) {
            const float t = cms<A>(x[i + 1], x[i], x[i]);
            const float u = A::sub(1.0f, x[i]);
            acc = A::add(acc, A::mad(A::mul(t, t), 100.0f, A::mul(u, u)));
            g[i] = A::add(g[i], A::mad(A::mul(400.0f, x[i]), t, A::mul(2.0f, u)));
            g[i + 1] = A::mad(-200.0f, t, g[i + 1]);
        }
        return -acc;
    }
};

// Rosenbrock2D, src/distributions.rs:502-523: -((a - x)^2 + b (y - x^2)^2)
template <class A>
struct Rosenbrock2D {
    static constexpr int kDim = 2;
    float a, b;
    __device__ __forceinline__ float logp_grad(const float (&x)[2], float (&g)[2]) const {
        const float u = A::sub(a, x[0]);
        const float t = cms<A>(x[1], x[0], x[0]);
        g[0] = A::mad(A::mul(A::mul(4.0f, b), x[0]), t, A::mul(2.0f, u));
        g[1] = A::mul(A::mul(-2.0f, b), t);
        return -A::mad(A::mul(t, t), b, A::mul(u, u));
    }
};

// DiffableGaussian2D, src/distributions.rs:262-288 (batched) / :296-315 (single chain):
//   z = delta^T P, quad = z . delta, logp = norm_const - 0.5 quad;  autodiff gradient -0.5 (z + P delta).
template <class A>
struct DiffGaussian2D {
    static constexpr int kDim = 2;
    float m0, m1, p00, p01, p10, p11, norm_const;
    __device__ __forceinline__ float logp_grad(const float (&x)[2], float (&g)[2]) const {
        const float d0 = A::sub(x[0], m0), d1 = A::sub(x[1], m1);
        const float z0 = A::mad(d1, p10, A::mul(d0, p00));
        const float z1 = A::mad(d1, p11, A::mul(d0, p01));
        const float w0 = A::mad(p01, d1, A::mul(p00, d0));
        const float w1 = A::mad(p11, d1, A::mul(p10, d0));
        const float quad = A::mad(z1, d1, A::mul(z0, d0));
        g[0] = A::mul(-0.5f, A::add(z0, w0));
        g[1] = A::mul(-0.5f, A::add(z1, w1));
        return A::add(-A::mul(quad, 0.5f), norm_const);
    }
};

// test target of src/nuts.rs:1024-1037: -(sum 0.5 x^2)
template <class A, int D>
struct StdNormal {
    static constexpr int kDim = D;
    __device__ __forceinline__ float logp_grad(const float (&x)[D], float (&g)[D]) const {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            acc = A::mad(A::mul(x[i], x[i]), 0.5f, acc);
            g[i] = -x[i];
        }
        return -acc;
    }
};

}  // namespace mmc
