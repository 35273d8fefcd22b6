// Shared device/host helpers for libminimcmc (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/minimcmc.h"

namespace mmc {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MMC_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return mmc::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MMC_REQUIRE(cond, ...)                                                \
    do {                                                                      \
        if (!(cond)) {                                                        \
            mmc::set_error(__VA_ARGS__);                                      \
            return MMC_ERR_INVALID;                                           \
        }                                                                     \
    } while (0)

int ensure_device();
int sm_count();

// registry of custom device targets (include/minimcmc_target.cuh): one entry per name, one launcher per sampler
struct CustomTargetEntry {
    int dim = 0;
    mmc_hmc_launch_fn hmc = nullptr;
    mmc_nuts_launch_fn nuts = nullptr;
    mmc_mh_launch_fn mh = nullptr;
};
int custom_target_register(const char *name, int dim, mmc_hmc_launch_fn hmc, mmc_nuts_launch_fn nuts, mmc_mh_launch_fn mh);
bool custom_target_get(int kind, CustomTargetEntry *out, const char **name = nullptr);
int custom_target_lookup(const char *name);

// ---------------------------------------------------------------- arithmetic policies
// Fast : plain operators, nvcc contracts a*b+c into FFMA (throughput build).
// Exact: round-to-nearest intrinsics that are never contracted, so a replayed trajectory reproduces
//        the CPU arithmetic (Rust never fuses) operation by operation.
struct Fast {
    static constexpr bool kContract = true;
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    // a*b + c
    static __device__ __forceinline__ float mad(float a, float b, float c) { return fmaf(a, b, c); }
};
struct Exact {
    static constexpr bool kContract = false;
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mad(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
};

// ---------------------------------------------------------------- Philox4x32-10 (see minimcmc.h "RNG contract")
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

__host__ __device__ __forceinline__ uint2 seed_key(uint64_t seed) { return make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); }

constexpr uint32_t kSubScalar = 0x80000000u;  // accept uniform (HMC) / Exp(1) (NUTS)
constexpr uint32_t kSubUnif = 0x80000001u;    // NUTS sequential uniforms
constexpr uint32_t kStepInit = 0xFFFFFFFFu;   // mmc_init_positions_dev

// uniform in (0,1) from the top 24 bits (never 0, safe for log)
__device__ __forceinline__ float u24_open(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// uniform in [0,1) with 24 bits, the reference's StandardUniform<f32>
__device__ __forceinline__ float u24_half_open(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u53_half_open(uint32_t lo, uint32_t hi) {
    const uint64_t bits = (uint64_t)lo | ((uint64_t)hi << 32);
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void box_muller_f32(uint32_t w0, uint32_t w1, float &n0, float &n1) {
    const float u1 = u24_open(w0);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * ((float)w1 * (1.0f / 4294967296.0f)), &s, &c);
    n0 = r * c;
    n1 = r * s;
}

__device__ __forceinline__ void box_muller_f64(uint4 w, double &n0, double &n1) {
    const uint64_t b1 = (uint64_t)w.x | ((uint64_t)w.y << 32);
    const uint64_t b2 = (uint64_t)w.z | ((uint64_t)w.w << 32);
    const double u1 = ((double)(b1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(b2 >> 11) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// D standard normals for (chain, step): sub j -> normals 4j..4j+3.
template <int D>
__device__ __forceinline__ void philox_normals_f32(uint2 key, uint64_t chain, uint32_t step, float (&out)[D]) {
#pragma unroll
    for (int j = 0; j < (D + 3) / 4; ++j) {
        const uint4 w = philox4x32_10(key, make_uint4((uint32_t)chain, (uint32_t)(chain >> 32), step, (uint32_t)j));
        float n[4];
        box_muller_f32(w.x, w.y, n[0], n[1]);
        if (4 * j + 2 < D) box_muller_f32(w.z, w.w, n[2], n[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * j + k < D) out[4 * j + k] = n[k];
    }
}

__device__ __forceinline__ uint4 philox_scalar_words(uint2 key, uint64_t chain, uint32_t step) {
    return philox4x32_10(key, make_uint4((uint32_t)chain, (uint32_t)(chain >> 32), step, kSubScalar));
}

}  // namespace mmc
