// K2, packed variant: one thread owns TWO chains and every arithmetic instruction of the trajectory is a packed
// f32x2 operation (FFMA2 / FADD2 / FMUL2, sm_100+).  The FP32 pipe retires the same 128 FMA/clk/SM either way
// (measured: 72 vs 74 TFLOP/s for FFMA vs FFMA2 streams), but a packed instruction occupies ONE issue slot for two
// lanes of work, which leaves issue slots for the non-FMA instructions of the loop: the scalar kernel is issue bound
// (86 % of issue slots busy at 76 % FMA-pipe utilisation).  Same formulas as hmc_run_kernel<Target, Fast>; the sums
// are associated as fused chains, so results agree with the scalar kernel to rounding (not bit for bit).
#pragma once

#include "mmc_hmc.cuh"

namespace mmc {

struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_pack(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_bcast(float a) { return f2_pack(a, a); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// RosenbrockND (mmc_targets.cuh) on packed pairs with the operation count trimmed for the FMA pipe: the same
// formulas, accumulated as fused chains (acc += u^2; acc += 100 t^2;  g_i = 400 x_i t_i + (2 u_i - 200 t_{i-1})).
// Returns -logp (the caller only needs the negated value).
template <int D>
struct RosenbrockND2 {
    static constexpr int kDim = D;
    __device__ __forceinline__ F2 neg_logp_grad(const F2 (&x)[D], F2 (&g)[D]) const {
        const F2 one = f2_bcast(1.0f), c100 = f2_bcast(100.0f), c400 = f2_bcast(400.0f), c2 = f2_bcast(2.0f),
                 cm2 = f2_bcast(-2.0f), cm200 = f2_bcast(-200.0f), cm1 = f2_bcast(-1.0f);
        F2 acc, t_prev;
#pragma unroll
        for (int i = 0; i + 1 < D; ++i) {
            const F2 t = fma2(mul2(x[i], cm1), x[i], x[i + 1]);     // x_{i+1} - x_i^2
            const F2 u = sub2(one, x[i]);
            acc = (i == 0) ? mul2(u, u) : fma2(u, u, acc);
            acc = fma2(mul2(t, t), c100, acc);
            const F2 two_u = fma2(x[i], cm2, c2);
            const F2 tail = (i == 0) ? two_u : fma2(cm200, t_prev, two_u);
            g[i] = fma2(mul2(x[i], t), c400, tail);
            t_prev = t;
        }
        g[D - 1] = mul2(cm200, t_prev);
        return acc;
    }
};

// kReplay: momenta / uniforms come from the caller's tapes (layout of hmc_run_kernel) instead of Philox; p.trace, when
// set, receives (logp_cur, logp_prop, accept_logp, accepted) per (step, chain) exactly like the scalar kernel, so the
// production kernel itself is held to the single-transition parity bar (tests/test_gpu_hmc.py).
// Draw write-out: a chain's draws are consecutive in memory ([chains, n_collect, D]), so a warp stages kT collected steps
// of its 64 chains in shared memory (row = chain, kT D floats) and flushes every row as one contiguous segment: 128-bit
// stores when the segments are 16-byte aligned, consecutive lanes on consecutive floats otherwise (odd pitches, unaligned
// tensors, the ragged end of a run) - instead of 32 strided 4-byte stores per step and coordinate.
template <int D> struct PairTile {
    static constexpr int kT = D <= 5 ? 8 : 4;                 // staged steps: kT D floats = a multiple of 32 bytes (D even for kT = 4)
    static constexpr int kRow = kT * D;                       // floats per staged row
    static constexpr int kPitch = kRow + 4;                   // padded: 16-byte aligned rows, lanes spread over the banks
    static constexpr bool kOk = (kRow * 4) % 16 == 0;
    static constexpr size_t kWarpFloats = (size_t)64 * kPitch;
};

template <class Target2, bool kReplay>
__global__ void __launch_bounds__(128) hmc_run_pair_kernel(const Target2 tgt, const HmcParams p) {
    constexpr int D = Target2::kDim;
    using Tile = PairTile<D>;
    extern __shared__ __align__(16) float hmc_pair_tiles[];
    const int lane = threadIdx.x & 31;
    float *tile = hmc_pair_tiles + (threadIdx.x >> 5) * Tile::kWarpFloats;   // rows 2 lane, 2 lane + 1 are this thread's chains
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c0 = 2 * t;
    const int64_t warp_c0 = c0 - 2 * lane;                    // first chain of this warp
    // 128-bit stores need every row segment of the warp on a 16-byte boundary; other shapes flush with coalesced 32-bit stores
    const bool tiled = p.out && !kReplay;
    const bool vec_ok = Tile::kOk && ((p.out_pitch * D) % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    auto flush = [&](int64_t s_first, int n_staged) {   // staged collected steps [s_first, s_first + n_staged) of the warp's chains
        __syncwarp();
        const int64_t rows = min((int64_t)64, p.chains - warp_c0);
        if (vec_ok && n_staged == Tile::kT) {
            constexpr int kVec = Tile::kRow / 4;                 // 128-bit stores per row
            for (int i = lane; i < (int)rows * kVec; i += 32) {
                const int r = i / kVec, k = i - r * kVec;
                const float4 v = *reinterpret_cast<const float4 *>(tile + r * Tile::kPitch + 4 * k);
                __stcs(reinterpret_cast<float4 *>(p.out + ((warp_c0 + r) * p.out_pitch + s_first) * D) + k, v);
            }
        } else {
            for (int i = lane; i < (int)rows * n_staged * D; i += 32) {
                const int r = i / (n_staged * D), k = i - r * (n_staged * D);
                p.out[((warp_c0 + r) * p.out_pitch + s_first) * D + k] = tile[r * Tile::kPitch + k];
            }
        }
        __syncwarp();
    };
    unsigned int n_acc = 0;
    // (threads of a partially filled warp keep running the loop below with a shadow chain so that the warp-wide flushes
    // stay convergent; they store nothing)
    const bool live = c0 < p.chains;
    if (live || (tiled && warp_c0 < p.chains)) {
        const bool two = live && c0 + 1 < p.chains;
        const int64_t cs0 = live ? c0 : p.chains - 1;   // shadow chain of an idle thread
        const int64_t c1 = two ? c0 + 1 : cs0;          // an odd tail shadows its own chain in the upper lane (never stored)
        F2 x[D], pos[D], mom[D], g[D];
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = f2_pack(p.positions[cs0 * D + i], p.positions[c1 * D + i]);
        int n_staged = 0;
        const F2 eps2 = f2_bcast(p.eps), eps_half2 = f2_bcast(p.eps * 0.5f), half2 = f2_bcast(0.5f), zero = f2_bcast(0.0f);
        const int64_t steps = p.n_collect + p.n_discard;
        const uint64_t g0 = (uint64_t)(cs0 + p.chain_offset), g1 = (uint64_t)(c1 + p.chain_offset);
        for (int64_t s = 0; s < steps; ++s) {
            const uint32_t gstep = (uint32_t)(p.step_base + s);
            float m0[D], m1[D], u0, u1;
            if (kReplay) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    m0[i] = __ldg(p.momenta + (s * p.chains + cs0) * D + i);
                    m1[i] = __ldg(p.momenta + (s * p.chains + c1) * D + i);
                }
                u0 = __ldg(p.u + s * p.chains + cs0);
                u1 = __ldg(p.u + s * p.chains + c1);
            } else {
                philox_normals_f32<D>(p.key, g0, gstep, m0);
                philox_normals_f32<D>(p.key, g1, gstep, m1);
                u0 = u24_half_open(philox_scalar_words(p.key, g0, gstep).x);
                u1 = u24_half_open(philox_scalar_words(p.key, g1, gstep).x);
            }
            const F2 nlp_cur = tgt.neg_logp_grad(x, g);
            F2 ke = zero;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                mom[i] = f2_pack(m0[i], m1[i]);
                pos[i] = x[i];
                ke = fma2(mom[i], mom[i], ke);
            }
            const F2 h_cur = fma2(ke, half2, nlp_cur);
            F2 nlp_prop = nlp_cur;
            // leapfrog with the two half-kicks between consecutive gradient evaluations merged into one full kick
            // (p += eps g; identical in exact arithmetic, src/hmc.rs:397-431 keeps them separate)
            if (p.n_leapfrog > 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) mom[i] = fma2(g[i], eps_half2, mom[i]);
            }
            for (int l = 0; l + 1 < p.n_leapfrog; ++l) {
#pragma unroll
                for (int i = 0; i < D; ++i) pos[i] = fma2(mom[i], eps2, pos[i]);
                nlp_prop = tgt.neg_logp_grad(pos, g);
#pragma unroll
                for (int i = 0; i < D; ++i) mom[i] = fma2(g[i], eps2, mom[i]);
            }
            if (p.n_leapfrog > 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) pos[i] = fma2(mom[i], eps2, pos[i]);
                nlp_prop = tgt.neg_logp_grad(pos, g);
#pragma unroll
                for (int i = 0; i < D; ++i) mom[i] = fma2(g[i], eps_half2, mom[i]);
            }
            F2 ke2 = zero;
#pragma unroll
            for (int i = 0; i < D; ++i) ke2 = fma2(mom[i], mom[i], ke2);
            const F2 h_prop = fma2(ke2, half2, nlp_prop);
            float a0, a1;
            f2_unpack(sub2(h_cur, h_prop), a0, a1);
            const bool acc0 = a0 >= logf(u0), acc1 = a1 >= logf(u1);
            n_acc += (unsigned)(acc0 && live) + (unsigned)(acc1 && two);
            if (kReplay && p.trace) {  // traces only exist for replay runs (mmc_hmc_run_dev)
                float c_lo, c_hi, q_lo, q_hi;
                f2_unpack(nlp_cur, c_lo, c_hi);
                f2_unpack(nlp_prop, q_lo, q_hi);
                float4 *tr = reinterpret_cast<float4 *>(p.trace) + s * p.chains;
                if (live) tr[c0] = make_float4(-c_lo, -q_lo, a0, acc0 ? 1.0f : 0.0f);
                if (two) tr[c1] = make_float4(-c_hi, -q_hi, a1, acc1 ? 1.0f : 0.0f);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float xl, xh, pl, ph;
                f2_unpack(x[i], xl, xh);
                f2_unpack(pos[i], pl, ph);
                x[i] = f2_pack(acc0 ? pl : xl, acc1 ? ph : xh);
            }
            if (s >= p.n_discard && p.out) {
                if (tiled) {
                    float *r0 = tile + (2 * lane) * Tile::kPitch + n_staged * D, *r1 = r0 + Tile::kPitch;
#pragma unroll
                    for (int i = 0; i < D; ++i) f2_unpack(x[i], r0[i], r1[i]);
                    if (++n_staged == Tile::kT) {
                        flush(s - p.n_discard - (Tile::kT - 1), Tile::kT);
                        n_staged = 0;
                    }
                } else if (live) {
                    float *o0 = p.out + (c0 * p.out_pitch + (s - p.n_discard)) * D;
                    float *o1 = p.out + (c1 * p.out_pitch + (s - p.n_discard)) * D;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        float xl, xh;
                        f2_unpack(x[i], xl, xh);
                        o0[i] = xl;
                        if (two) o1[i] = xh;
                    }
                }
            }
        }
        if (tiled && n_staged > 0) flush(p.n_collect - n_staged, n_staged);
        if (live) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float xl, xh;
                f2_unpack(x[i], xl, xh);
                p.positions[c0 * D + i] = xl;
                if (two) p.positions[c1 * D + i] = xh;
            }
        }
    }
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(p.accept_count, (unsigned long long)n_acc);
}

template <class Target2>
int launch_hmc_pair(const Target2 &tgt, const HmcParams &p, bool replay, cudaStream_t stream) {
    const int block = 128;
    const int64_t threads = (p.chains + 1) / 2;
    const unsigned grid = (unsigned)((threads + block - 1) / block);
    const size_t smem = (size_t)(block / 32) * PairTile<Target2::kDim>::kWarpFloats * sizeof(float);
    if (replay) {
        hmc_run_pair_kernel<Target2, true><<<grid, block, 0, stream>>>(tgt, p);
    } else {
        auto kern = hmc_run_pair_kernel<Target2, false>;
        if (smem > 48 * 1024) MMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, block, smem, stream>>>(tgt, p);
    }
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace mmc
