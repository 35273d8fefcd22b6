// K4b: NUTS with several chains per warp ("group" layout; see include/minimcmc.h "NUTS", mmc_nuts_set_layout).
//
// Same algorithm, RNG contract and outputs as the one-chain-per-warp kernel of mmc_nuts.cuh
// (NUTSChain::{init_chain, step, run, run_progress} src/nuts.rs:457-691, build_tree :764-946, leapfrog :979-996,
// stop_criterion :963-977, find_reasonable_epsilon :695-761), but a warp carries 32 / G chains: G lanes per chain,
// E consecutive vector elements per lane (G E >= D).  The tree bookkeeping of the warp kernel (merge walk, level
// addressing, exp, uniform draws; about two thirds of its ~300 warp-instructions per leapfrog at D = 100) is issued
// once per warp and therefore shared by 32 / G chains, the reductions are log2 G butterfly steps instead of five, and
// at D = 100 a lane carries 13 (reference arithmetic) or 14 (packed f32x2 arithmetic) elements, so 100 of 104 / 112
// element slots hold data (the warp kernel: 100 of 128).
//
// The chains of a warp advance in lock step, one transition at a time.  Control flow is warp-uniform (loops run
// while ANY group of the warp still needs them, decided by votes) and every state update is predicated with the
// group's own flags, so the full-mask shuffles are always executed convergently:
//   * a group whose transition has ended (U-turn, divergence, max depth) idles until the others end theirs;
//   * inside a doubling a group that is not building (finished, or its subtree failed) integrates with step size 0,
//     which leaves (x, p) untouched, and none of its results are committed;
//   * the binary-counter merge walk visits the levels bottom-up after every pair of leaves: groups still carrying a
//     valid subtree merge at the set bits of the leaf index and park at the first clear bit; a failed subtree keeps
//     merging at all set bits and passes through the clear ones - exactly the RNG consumption and alpha / n_alpha sums
//     of the recursion.
// Registers hold only the edge being extended (x, p, grad), the subtree proposal and, during a merge walk, the first
// leaf of the subtree; the opposite edge and the current position are parked in per-lane shared-memory columns and
// swapped in when a group changes direction (the U-turn test (x+ - x-).p- >= 0 && (x+ - x-).p+ >= 0 is symmetric in
// the two momenta, and -(a - b) == b - a exactly, so it is evaluated as "extended edge vs the other state").
// Leaves are built in pairs: the level-0 merge of the binary counter happens in registers, only levels >= 1 are
// parked (levels 1..GrpTune<E>::kLevels in shared memory, deeper ones in an L2-resident scratch).
// Throughput arithmetic (policy Fast, D > 4) runs on packed f32x2 instructions over pair-interleaved lanes
// (GRosenbrockNDP).  Native runs are dispensed to the persistent warps as (group of chains, slice of the run) tickets
// with a completion flag per group, and a launch can be restricted to a window of iterations and take its groups
// from a permutation of the chains (mmc_nuts_set_slicing / mmc_nuts_set_regroup, host side in mmc_nuts.cu).
#pragma once

#include "mmc_hmc_pair.cuh"  // F2: packed f32x2 helpers
#include "mmc_nuts.cuh"

namespace mmc {

// Occupancy per elements-per-lane E (tuning builds override both): wide lanes (E > 8) keep 168 registers = 12 warps / SM
// and two tree levels in shared memory ((4 + 3 x 2) parked vectors of 32 E floats per warp: 215 KB / SM at E = 14);
// narrow lanes fit 128 registers = 16 warps / SM with three levels.
#ifdef MMC_NUTS_GROUP_MIN_BLOCKS
template <int E> struct GrpTune { static constexpr int kMinBlocks = MMC_NUTS_GROUP_MIN_BLOCKS; static constexpr int kLevels = MMC_NUTS_GROUP_SMEM_LEVELS; };
#else
template <int E> struct GrpTune { static constexpr int kMinBlocks = E > 8 ? 3 : 4; static constexpr int kLevels = E > 8 ? 2 : 3; };
#endif
constexpr int kGrpWarps = 4;
constexpr int kGrpMaxLevels = 16;

// the uniform refill (one Philox block per 2 G draws) is kept out of line: inlined at every draw site it would add
// ~75 rarely executed instructions to a hot loop whose instruction-cache footprint is what limits its issue rate
static __device__ __noinline__ uint4 philox_block_call(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return philox4x32_10(make_uint2(k0, k1), make_uint4(c0, c1, c2, c3));
}

template <class A, int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = A::add(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
template <class A, int G>
__device__ __forceinline__ void group_sum2(float &a, float &b) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const float ta = __shfl_xor_sync(kFull, a, o);
        const float tb = __shfl_xor_sync(kFull, b, o);
        a = A::add(a, ta);
        b = A::add(b, tb);
    }
}
// four independent sums in one butterfly: the shuffles of a step overlap instead of serialising
template <class A, int G>
__device__ __forceinline__ void group_sum4(float &a, float &b, float &c, float &d) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const float ta = __shfl_xor_sync(kFull, a, o);
        const float tb = __shfl_xor_sync(kFull, b, o);
        const float tc = __shfl_xor_sync(kFull, c, o);
        const float td = __shfl_xor_sync(kFull, d, o);
        a = A::add(a, ta);
        b = A::add(b, tb);
        c = A::add(c, tc);
        d = A::add(d, td);
    }
}
// four interleaved partial sums keep the dependent chain of an E-term accumulation at E / 4 operations
template <class A>
__device__ __forceinline__ float fold4(const float (&s)[4]) { return A::add(A::add(s[0], s[1]), A::add(s[2], s[3])); }

// ---------------------------------------------------------------- group-form targets
// interface: float logp_grad(const float (&x)[E], float (&g)[E], int gl) const; gl = lane inside the group, which
// owns elements gl*E .. gl*E+E-1 (zero padded).  kPartial as in the warp form.

// RosenbrockND (src/distributions.rs:531-547) for D <= G E.
template <class A, int E, int G>
struct GRosenbrockND {
    static constexpr bool kPartial = true;
    static constexpr bool kPacked = false;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        const float xn = __shfl_down_sync(kFull, x[0], 1, G);
        const int nv = D - 1 - gl * E;  // elements e < nv of this lane have a successor (i + 1 < D)
        float t[E];
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool valid = e < nv;
            const float xnext = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : xn;
            const float tt = valid ? cms<A>(xnext, x[e], x[e]) : 0.0f;
            const float u = valid ? A::sub(1.0f, x[e]) : 0.0f;
            t[e] = tt;
            acc[e & 3] = A::add(acc[e & 3], A::mad(A::mul(tt, tt), 100.0f, A::mul(u, u)));
            g[e] = A::mad(A::mul(400.0f, x[e]), tt, A::mul(2.0f, u));
        }
        float tprev = __shfl_up_sync(kFull, t[E - 1], 1, G);
        if (gl == 0) tprev = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float tp = e == 0 ? tprev : t[e == 0 ? 0 : e - 1];
            g[e] = A::add(A::mul(-200.0f, tp), g[e]);
        }
        return -fold4<A>(acc);
    }
};

// RosenbrockND on packed f32x2 pairs (throughput arithmetic only).  Lane layout "pair interleaved": array slot a of a
// lane holds element gl E + (a >> 1) + (a & 1) E / 2, so that slots (2k, 2k+1) = elements (k, k + E/2) form one f32x2
// register pair and the successor / predecessor pairs of the Rosenbrock stencil are the neighbouring pairs (one
// repacked pair each at the ends).  Same formulas as GRosenbrockND with the products regrouped for FMA chains:
// t' = mk (x_{i+1} - x_i^2), u' = mk (1 - x_i), g_i = (4 x_i)(100 t'_i) + 2 u'_i - 2 (100 t'_{i-1}); mk = 1 where the
// element has a successor, else 0.
template <int E, int G>
struct GRosenbrockNDP {
    static constexpr bool kPartial = true;
    static constexpr bool kPacked = true;
    static_assert(E % 2 == 0, "packed layout needs an even number of elements per lane");
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        constexpr int H = E / 2;
        const float nvf = (float)(D - 1 - gl * E);  // elements with offset < nv have a successor
        const F2 one = f2_bcast(1.0f), c100 = f2_bcast(100.0f), c4 = f2_bcast(4.0f), cm2 = f2_bcast(-2.0f), cm1 = f2_bcast(-1.0f);
        F2 X[H], T100[H];
#pragma unroll
        for (int k = 0; k < H; ++k) X[k] = f2_pack(x[2 * k], x[2 * k + 1]);
        const float xn_lane = __shfl_down_sync(kFull, x[0], 1, G);  // element 0 of the next lane
        F2 acc0 = f2_bcast(0.0f), acc1 = f2_bcast(0.0f);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const F2 XN = (k + 1 < H) ? X[(k + 1 < H) ? k + 1 : k] : f2_pack(x[1], xn_lane);
            const F2 mk = f2_pack((float)k < nvf ? 1.0f : 0.0f, (float)(k + H) < nvf ? 1.0f : 0.0f);
            const F2 nX = mul2(X[k], cm1);
            const F2 T = mul2(fma2(nX, X[k], XN), mk);
            const F2 U = fma2(nX, mk, mk);
            T100[k] = mul2(T, c100);
            if (k & 1) acc1 = fma2(T100[k], T, fma2(U, U, acc1));
            else acc0 = fma2(T100[k], T, fma2(U, U, acc0));
            const F2 Gk = fma2(mul2(X[k], c4), T100[k], add2(U, U));
            f2_unpack(Gk, g[2 * k], g[2 * k + 1]);
        }
        float tl_lo, tl_hi;
        f2_unpack(T100[H - 1], tl_lo, tl_hi);                // 100 t' of elements H - 1 and E - 1
        float tprev = __shfl_up_sync(kFull, tl_hi, 1, G);     // last element of the previous lane
        if (gl == 0) tprev = 0.0f;
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const F2 TP = (k > 0) ? T100[(k > 0) ? k - 1 : 0] : f2_pack(tprev, tl_lo);
            const F2 Gk = fma2(cm2, TP, f2_pack(g[2 * k], g[2 * k + 1]));
            f2_unpack(Gk, g[2 * k], g[2 * k + 1]);
        }
        float a_lo, a_hi;
        f2_unpack(add2(acc0, acc1), a_lo, a_hi);
        return -(a_lo + a_hi);
    }
};

// StdNormal (src/nuts.rs:1024-1037); padding elements are zero.
template <class A, int E, int G>
struct GStdNormal {
    static constexpr bool kPartial = true;
    static constexpr bool kPacked = false;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc[e & 3] = A::mad(A::mul(x[e], x[e]), 0.5f, acc[e & 3]);
            g[e] = -x[e];
        }
        return -fold4<A>(acc);
    }
};

// Any small thread-form target (kDim <= G E): every lane of the group gathers the full vector and evaluates it.
template <class T, int E, int G>
struct GSmall {
    static constexpr bool kPartial = false;
    static constexpr bool kPacked = false;
    T t;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        constexpr int K = T::kDim;
        static_assert(K <= E * G, "target does not fit the group");
        float xf[K], gf[K];
#pragma unroll
        for (int i = 0; i < K; ++i) xf[i] = __shfl_sync(kFull, x[i % E], i / E, G);
        const float lp = t.logp_grad(xf, gf);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            g[e] = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (gl * E + e == i) g[e] = gf[i];
        }
        return lp;
    }
};

template <class Target, class A, class ST, int E, int G, bool kReplay>
struct NutsGroup {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "lanes per chain must be a power of two");
    static constexpr int NG = 32 / G;             // chains per warp
    static constexpr int QF = E / 4, RM = E % 4;  // a parked vector: QF float4 slabs + RM float slabs, [slab][lane]
    static constexpr int kVec = E * 32;           // floats of one parked vector of the whole warp
    static constexpr int kParked = 4;             // opposite edge (x, p, grad) + current position
    static constexpr bool kFusedKick = !kReplay && A::kContract;  // native throughput runs only
    static constexpr bool kPk = Target::kPacked;  // pair-interleaved lane layout, f32x2 arithmetic (see GRosenbrockNDP)
    static_assert(!kPk || (A::kContract && E % 2 == 0), "the packed layout exists for the throughput policy only");
    // element offset inside the lane of array slot a
    static __host__ __device__ constexpr int off(int a) { return kPk ? (a >> 1) + (a & 1) * (E / 2) : a; }
    static constexpr int kL = GrpTune<E>::kLevels;  // tree levels 1..kL are parked in shared memory
    static constexpr int kWarpFloats = (kParked + 3 * kL) * kVec;
    static constexpr int kScalBytes = kGrpMaxLevels * NG * 16 + 32 * 4;  // per warp: (double alpha, int n, int n_alpha) per level and group + depth histogram

    const Target &tgt;
    const NutsParams &p;
    const int lane, gl, grp;
    float *s_park;       // shared: [kParked] vectors of this warp; every lane only touches its own column
    float *s_stack;      // shared: [kL][3] vectors (tree levels 1..kL)
    float *g_stack;      // global scratch for deeper levels, same layout
    double *s_a;         // [kGrpMaxLevels][NG]
    int *s_n, *s_na;
    // per group (identical in the G lanes of a group)
    uint64_t gchain = 0;
    int64_t chain = 0;
    uint32_t step_word = 0;
    uint32_t q = 0;                 // uniforms consumed in this step
    uint32_t q_batch = 0xffffffffu;
    uint4 ubatch;
    int64_t cur_n = 0, cur_e = 0, cur_u = 0;
    uint32_t n_grad = 0, n_unif = 0;

    __device__ NutsGroup(const Target &t, const NutsParams &pp, int ln, float *warp_smem, float *gs, void *scal)
        : tgt(t), p(pp), lane(ln), gl(ln % G), grp(ln / G), s_park(warp_smem), s_stack(warp_smem + kParked * kVec), g_stack(gs) {
        s_a = reinterpret_cast<double *>(scal);
        s_n = reinterpret_cast<int *>(s_a + kGrpMaxLevels * NG);
        s_na = s_n + kGrpMaxLevels * NG;
        ubatch = make_uint4(0, 0, 0, 0);
    }

    // ---- parked vectors: explicit shared / global paths so the compiler emits LDS/STS and LDG/STG
    template <bool kGlobal>
    __device__ __forceinline__ void load_vec(const float *base, float (&v)[E]) {
        const float4 *s4 = reinterpret_cast<const float4 *>(base) + lane;
#pragma unroll
        for (int k4 = 0; k4 < QF; ++k4) {
            const float4 t = kGlobal ? __ldcg(s4 + k4 * 32) : s4[k4 * 32];
            v[(4 * k4 + 0) % E] = t.x; v[(4 * k4 + 1) % E] = t.y; v[(4 * k4 + 2) % E] = t.z; v[(4 * k4 + 3) % E] = t.w;
        }
        const float *s1 = base + QF * 128 + lane;
#pragma unroll
        for (int r = 0; r < RM; ++r) v[(4 * QF + r) % E] = kGlobal ? __ldcg(s1 + r * 32) : s1[r * 32];
    }
    template <bool kGlobal>
    __device__ __forceinline__ void store_vec(float *base, const float (&v)[E]) {
        float4 *d4 = reinterpret_cast<float4 *>(base) + lane;
#pragma unroll
        for (int k4 = 0; k4 < QF; ++k4) {
            const float4 t = make_float4(v[(4 * k4 + 0) % E], v[(4 * k4 + 1) % E], v[(4 * k4 + 2) % E], v[(4 * k4 + 3) % E]);
            if (kGlobal) __stcg(d4 + k4 * 32, t); else d4[k4 * 32] = t;
        }
        float *d1 = base + QF * 128 + lane;
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            if (kGlobal) __stcg(d1 + r * 32, v[(4 * QF + r) % E]); else d1[r * 32] = v[(4 * QF + r) % E];
        }
    }
    // tree level lvl >= 1 (which: 0 = first-leaf x, 1 = first-leaf p, 2 = proposal)
    __device__ __forceinline__ void load_level(int lvl, int which, float (&v)[E]) {
        const int slot = lvl - 1;
        if (slot < kL) load_vec<false>(s_stack + (slot * 3 + which) * kVec, v);
        else load_vec<true>(g_stack + ((slot - kL) * 3 + which) * kVec, v);
    }
    __device__ __forceinline__ void store_level(int lvl, int which, const float (&v)[E]) {
        const int slot = lvl - 1;
        if (slot < kL) store_vec<false>(s_stack + (slot * 3 + which) * kVec, v);
        else store_vec<true>(g_stack + ((slot - kL) * 3 + which) * kVec, v);
    }
    // parked slots: 0..2 = opposite edge (x, p, grad), 3 = current position
    __device__ __forceinline__ void load_parked(int which, float (&v)[E]) { load_vec<false>(s_park + which * kVec, v); }
    __device__ __forceinline__ void store_parked(int which, const float (&v)[E]) { store_vec<false>(s_park + which * kVec, v); }
    // exchange a working vector with its parked counterpart in the groups with `doit`
    __device__ __forceinline__ void swap_parked(int which, float (&w)[E], bool doit) {
        float t[E];
        load_parked(which, t);
        if (doit) {
            store_parked(which, w);
#pragma unroll
            for (int k = 0; k < E; ++k) w[k] = t[k];
        }
    }

    // ---- random draws: the counters of the warp kernel (minimcmc.h "RNG contract"): normal i of a step comes from
    // word i & 3 of Philox block i >> 2, uniform q from words (q & 1 ? zw : xy) of block kSubUnif + (q >> 1)
    __device__ __forceinline__ void draw_normals(float (&m)[E]) {
        if (kReplay) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = gl * E + off(e);
                m[e] = i < p.D ? (float)p.normals[chain * p.cap_normals + cur_n + i] : 0.0f;
            }
            cur_n += p.D;
        } else {
            constexpr int NQ = (E % 4 == 0) ? E / 4 : (E + 6) / 4;  // Philox blocks spanned by E consecutive elements
            const int i0 = gl * E;
            const int r = (E % 4 == 0) ? 0 : (i0 & 3);
            float nb[NQ * 4];
#pragma unroll
            for (int b = 0; b < NQ; ++b) {
                const uint32_t blk = (uint32_t)(i0 >> 2) + b;
                if ((int)(blk * 4) < p.D) {
                    const uint4 w = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word, blk));
                    box_muller_f32(w.x, w.y, nb[4 * b + 0], nb[4 * b + 1]);
                    box_muller_f32(w.z, w.w, nb[4 * b + 2], nb[4 * b + 3]);
                } else {
                    nb[4 * b + 0] = nb[4 * b + 1] = nb[4 * b + 2] = nb[4 * b + 3] = 0.0f;
                }
            }
#pragma unroll
            for (int a = 0; a < E; ++a) {
                constexpr int kN = NQ * 4;
                const int e = off(a);
                float v = nb[e];
                if (E % 4 != 0) {
                    if (r == 1) v = nb[(e + 1) % kN];
                    if (r == 2) v = nb[(e + 2) % kN];
                    if (r == 3) v = nb[(e + 3) % kN];
                }
                m[a] = (i0 + e < p.D) ? v : 0.0f;
            }
        }
    }
    __device__ __forceinline__ ST draw_exp1() {
        if (kReplay) return (ST)p.exps[chain * p.cap_exps + cur_e++];
        const uint4 w = philox_scalar_words(p.key, gchain, step_word);
        return (ST)(-logf(u24_open(w.x)));
    }
    // the two words behind uniform q of this group's step; consumed (q advances) only where `take`
    __device__ __forceinline__ void next_words(bool take, uint32_t &lo, uint32_t &hw) {
        const uint32_t batch = q / (2 * G);  // G blocks = 2 G uniforms per refill, block b G + gl in lane gl
        const bool refill = take && batch != q_batch;
        if (__any_sync(kFull, refill)) {
            const uint4 nb = philox_block_call(p.key.x, p.key.y, (uint32_t)gchain, (uint32_t)(gchain >> 32), step_word,
                                               kSubUnif + batch * G + (uint32_t)gl);
            if (refill) {
                ubatch = nb;
                q_batch = batch;
            }
        }
        const int src = (q >> 1) & (G - 1);
        const bool hi = q & 1;
        lo = __shfl_sync(kFull, hi ? ubatch.z : ubatch.x, src, G);
        hw = __shfl_sync(kFull, hi ? ubatch.w : ubatch.y, src, G);
        if (take) {
            ++q;
            ++n_unif;
        }
    }
    // next uniform of the step where `take`; f64 = 53-bit (tree merges), otherwise type T
    __device__ __forceinline__ double draw_uniform(bool f64, bool take) {
        if (kReplay) {
            double u = 0.5;
            if (take) {
                u = p.unifs[chain * p.cap_unifs + cur_u++];
                ++n_unif;
            }
            return u;
        }
        uint32_t lo, hw;
        next_words(take, lo, hw);
        if (f64 || sizeof(ST) == 8) return u53_half_open(lo, hw);
        return (double)u24_half_open(hw);
    }
    // u < n'' / max(n' + n'', 1) in f64 (src/nuts.rs:910-911) for a merge of a parked half with n' = an and the later
    // half with n'' = tn; consumes one uniform where `take`
    __device__ __forceinline__ bool draw_take_later(int an, int tn, bool take) {
        if (kReplay) {
            const double u = draw_uniform(true, take);
            if (tn == 0) return false;
            if (an == 0) return u < 1.0;
            return u < ((double)tn / (double)(an + tn));
        }
        // native draws are k 2^-53 with a 53-bit integer k: k (n' + n'') < n'' 2^53 is the same test evaluated exactly
        // (the f64 quotient is rounded, which can only matter when u equals the rounded quotient itself)
        uint32_t lo, hw;
        next_words(take, lo, hw);
        const uint64_t k53 = (((uint64_t)hw << 32) | lo) >> 11;
        return tn != 0 && (an == 0 || u53_below_ratio(k53, (uint32_t)tn, (uint32_t)(an + tn)));
    }

    // ---- leapfrog, src/nuts.rs:979-996 (in place); returns logp' (partial when Target::kPartial)
    // p + (g e) 0.5 in the reference's operation order; the throughput policy folds it into one FMA with e / 2
    static __device__ __forceinline__ float half_kick(float g, float e, float he, float m) {
        if (kFusedKick) return fmaf(g, he, m);
        return A::mad(A::mul(g, e), 0.5f, m);
    }
    __device__ __forceinline__ float leapfrog(float (&x)[E], float (&m)[E], float (&g)[E], float e) {
        const float he = 0.5f * e;
        if constexpr (kPk) {
            constexpr int H = E / 2;
            const F2 e2 = f2_bcast(e), he2 = f2_bcast(he), half2 = f2_bcast(0.5f);
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const F2 Gk = f2_pack(g[2 * k], g[2 * k + 1]);
                F2 Mk = f2_pack(m[2 * k], m[2 * k + 1]);
                Mk = kFusedKick ? fma2(Gk, he2, Mk) : fma2(mul2(Gk, e2), half2, Mk);
                const F2 Xk = fma2(Mk, e2, f2_pack(x[2 * k], x[2 * k + 1]));
                f2_unpack(Mk, m[2 * k], m[2 * k + 1]);
                f2_unpack(Xk, x[2 * k], x[2 * k + 1]);
            }
            const float lp = tgt.logp_grad(x, g, gl);
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const F2 Gk = f2_pack(g[2 * k], g[2 * k + 1]);
                F2 Mk = f2_pack(m[2 * k], m[2 * k + 1]);
                Mk = kFusedKick ? fma2(Gk, he2, Mk) : fma2(mul2(Gk, e2), half2, Mk);
                f2_unpack(Mk, m[2 * k], m[2 * k + 1]);
            }
            return lp;
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                m[k] = half_kick(g[k], e, he, m[k]);
                x[k] = A::mad(m[k], e, x[k]);
            }
            const float lp = tgt.logp_grad(x, g, gl);
#pragma unroll
            for (int k = 0; k < E; ++k) m[k] = half_kick(g[k], e, he, m[k]);
            return lp;
        }
    }
    __device__ __forceinline__ float local_sumsq(const float (&m)[E]) {
        if constexpr (kPk) {
            F2 s0 = f2_bcast(0.0f), s1 = f2_bcast(0.0f);
#pragma unroll
            for (int k = 0; k < E / 2; ++k) {
                const F2 Mk = f2_pack(m[2 * k], m[2 * k + 1]);
                if (k & 1) s1 = fma2(Mk, Mk, s1); else s0 = fma2(Mk, Mk, s0);
            }
            float lo, hi;
            f2_unpack(add2(s0, s1), lo, hi);
            return lo + hi;
        } else {
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < E; ++k) s[k & 3] = A::mad(m[k], m[k], s[k & 3]);
            return fold4<A>(s);
        }
    }
    __device__ __forceinline__ float sumsq(const float (&m)[E]) { return group_sum<A, G>(local_sumsq(m)); }
    // completes a target evaluation: full logp (lp_io) and sum m^2, one fused butterfly when logp is partial
    __device__ __forceinline__ float finish(float &lp_io, const float (&m)[E]) {
        float s = local_sumsq(m);
        if (Target::kPartial) group_sum2<A, G>(lp_io, s);
        else s = group_sum<A, G>(s);
        return s;
    }
    __device__ __forceinline__ float full_logp(float lp) { return Target::kPartial ? group_sum<A, G>(lp) : lp; }
    // stop_criterion, src/nuts.rs:963-977 (true = keep going) between the edge being extended (xw, pw) and another
    // state (xo, po); plus: the extended edge is the plus side
    __device__ __forceinline__ bool keep_going(const float (&xw)[E], const float (&xo)[E], const float (&pw)[E],
                                               const float (&po)[E], bool plus) {
        float a, b;
        local_turn(xw, xo, pw, po, a, b);
        group_sum2<A, G>(a, b);
        return turn_ok(a, b, plus);
    }
    // the sign of (x+ - x-) is applied to the two sums instead of every term (negation commutes with rounding)
    static __device__ __forceinline__ bool turn_ok(float a, float b, bool plus) {
        return plus ? (a >= 0.0f && b >= 0.0f) : (-a >= 0.0f && -b >= 0.0f);
    }
    // this lane's part of (xw - xo).pw and (xw - xo).po
    __device__ __forceinline__ void local_turn(const float (&xw)[E], const float (&xo)[E], const float (&pw)[E],
                                               const float (&po)[E], float &a, float &b) {
        if constexpr (kPk) {
            F2 w0 = f2_bcast(0.0f), w1 = f2_bcast(0.0f), t0 = f2_bcast(0.0f), t1 = f2_bcast(0.0f);
#pragma unroll
            for (int k = 0; k < E / 2; ++k) {
                const F2 d = sub2(f2_pack(xw[2 * k], xw[2 * k + 1]), f2_pack(xo[2 * k], xo[2 * k + 1]));
                const F2 PW = f2_pack(pw[2 * k], pw[2 * k + 1]), PO = f2_pack(po[2 * k], po[2 * k + 1]);
                if (k & 1) { w1 = fma2(d, PW, w1); t1 = fma2(d, PO, t1); }
                else { w0 = fma2(d, PW, w0); t0 = fma2(d, PO, t0); }
            }
            float a_lo, a_hi, b_lo, b_hi;
            f2_unpack(add2(w0, w1), a_lo, a_hi);
            f2_unpack(add2(t0, t1), b_lo, b_hi);
            a = a_lo + a_hi;
            b = b_lo + b_hi;
        } else {
            float dw[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const float diff = A::sub(xw[k], xo[k]);
                dw[k & 3] = A::mad(diff, pw[k], dw[k & 3]);
                dt[k & 3] = A::mad(diff, po[k], dt[k & 3]);
            }
            a = fold4<A>(dw);
            b = fold4<A>(dt);
        }
    }
    __device__ __forceinline__ bool group_all(bool ok) {
        const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp * G);
        return (__ballot_sync(kFull, ok) & gmask) == gmask;
    }
    __device__ __forceinline__ bool all_finite(const float (&v)[E]) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < E; ++k) ok = ok && isfinite(v[k]);
        return group_all(ok);
    }

    // find_reasonable_epsilon, src/nuts.rs:695-761, for the groups with `need`; the trial leapfrogs are recomputed by
    // every group with its own (unchanged) step size, which reproduces the values it already has
    __device__ ST find_reasonable_epsilon(const float (&x0)[E], const float (&m0)[E], bool need) {
        float g0[E], x[E], m[E], g[E];
        const ST half = (ST)0.5;
        ST epsilon = (ST)1.0;
        const float ulogp = full_logp(tgt.logp_grad(x0, g0, gl));
        if (need) ++n_grad;
        auto leap = [&](ST e) {
#pragma unroll
            for (int k = 0; k < E; ++k) { x[k] = x0[k]; m[k] = m0[k]; g[k] = g0[k]; }
            return full_logp(leapfrog(x, m, g, (float)e));
        };
        float ulogp_prime = leap(epsilon);
        if (need) ++n_grad;
        ST k = (ST)1.0;
        while (true) {
            const bool fin = all_finite(g);
            const bool again = need && !isfinite(ulogp_prime) && !fin;
            if (!__any_sync(kFull, again)) break;
            if (again) { k = k * half; ++n_grad; }
            ulogp_prime = leap(epsilon * k);
        }
        epsilon = half * k * epsilon;
        const float pp0 = sumsq(m0);
        float lap_f = A::sub(A::sub(ulogp_prime, ulogp), A::mul(A::sub(sumsq(m), pp0), 0.5f));
        ST lap = (ST)(double)lap_f;
        const ST a = lap > s_log(half) ? (ST)1.0 : (ST)-1.0;
        while (true) {
            const bool again = need && (a * lap > -a * s_log((ST)2.0));
            if (!__any_sync(kFull, again)) break;
            if (again) { epsilon = epsilon * s_pow((ST)2.0, a); ++n_grad; }
            ulogp_prime = leap(epsilon);
            lap_f = A::sub(A::sub(ulogp_prime, ulogp), A::mul(A::sub(sumsq(m), pp0), 0.5f));
            lap = (ST)(double)lap_f;
        }
        return epsilon;
    }

    // one leaf of a subtree (BuildTree base case, src/nuts.rs:780-826): a leapfrog for the groups with `act` (the
    // others integrate with step size 0 and ignore the results), then n', s' and min(1, exp(joint - joint0))
    // min(1, exp(.)) only feeds the dual-averaging statistic: native throughput runs use the 2-instruction ex2 form
    static __device__ __forceinline__ ST accept_exp(ST v) {
        if constexpr (kFusedKick && sizeof(ST) == 4) return (ST)__expf((float)v);
        else return s_exp(v);
    }
    __device__ __forceinline__ void leaf(float (&cx)[E], float (&cm)[E], float (&cg)[E], float veps, bool act, ST logu,
                                         ST joint0, int &ln, bool &ls, ST &la) {
        float lp = leapfrog(cx, cm, cg, act ? veps : 0.0f);
        const float ss = finish(lp, cm);
        leaf_scalars(lp, ss, act, logu, joint0, ln, ls, la);
    }
    // the second leaf of a pair: its energy sums and the U-turn sums against the pair's first leaf (fx, fm) share one
    // butterfly; kg = stop_criterion between the two leaves
    __device__ __forceinline__ void leaf_turn(float (&cx)[E], float (&cm)[E], float (&cg)[E], float veps, bool act, ST logu,
                                              ST joint0, const float (&fx)[E], const float (&fm)[E], bool plus, int &ln,
                                              bool &ls, ST &la, bool &kg) {
        float lp = leapfrog(cx, cm, cg, act ? veps : 0.0f);
        float ss = local_sumsq(cm), a, b;
        local_turn(cx, fx, cm, fm, a, b);
        if (Target::kPartial) {
            group_sum4<A, G>(lp, ss, a, b);
        } else {
            group_sum2<A, G>(a, b);
            ss = group_sum<A, G>(ss);
        }
        kg = turn_ok(a, b, plus);
        leaf_scalars(lp, ss, act, logu, joint0, ln, ls, la);
    }
    __device__ __forceinline__ void leaf_scalars(float lp, float ss, bool act, ST logu, ST joint0, int &ln, bool &ls, ST &la) {
        const float joint_f = A::sub(lp, A::mul(ss, 0.5f));
        const ST joint = (ST)(double)joint_f;
        const ST ex = accept_exp(joint - joint0);
        ln = (logu < joint) ? 1 : 0;
        ls = (logu - (ST)1000.0) < joint;
        la = ((ST)1.0 < ex || ex != ex) ? (ST)1.0 : ex;  // T::min(1, e): NaN -> 1
        if (act) ++n_grad;
    }

    // One doubling = build_tree(edge, v, j), src/nuts.rs:764-946, iteratively, for the groups with `run`.
    // cx/cm/cg: the edge to extend (in) and the new edge (out).  For the `run` groups the outputs are the subtree
    // proposal (prop), n', s', alpha and n_alpha; other groups keep their scalars and never read their prop.
    __device__ void doubling(float (&cx)[E], float (&cm)[E], float (&cg)[E], bool plus, int j, ST logu, ST eps, ST joint0,
                             bool run, float (&prop)[E], int &n_out, bool &s_out, ST &alpha_out, int &nalpha_out) {
        const float veps = (float)(plus ? eps : -eps);
        if (j == 0) {
            int ln; bool ls; ST la;
            leaf(cx, cm, cg, veps, run, logu, joint0, ln, ls, la);
            if (run) {
#pragma unroll
                for (int k = 0; k < E; ++k) prop[k] = cx[k];
                n_out = ln; s_out = ls; alpha_out = la; nalpha_out = 1;
            }
            return;
        }
        const uint32_t n_pairs = 1u << (j - 1);
        int tn = 0, tna = 0;
        ST ta = (ST)0.0;
        bool ts = true;
        bool building = run;
        for (uint32_t pr = 0; pr < n_pairs; ++pr) {
            // ---- leaves 2 pr and 2 pr + 1; their merge (level 0 of the binary counter) stays in registers
            float tfx[E], tfm[E];  // first leaf of the subtree being merged upward (its proposal: prop)
            const bool pending = building;  // groups that have to merge / park the subtree ending in this pair
            {
                int ln; bool ls; ST la;
                leaf(cx, cm, cg, veps, pending, logu, joint0, ln, ls, la);
                if (pending) { tn = ln; ts = ls; ta = la; tna = 1; }
#pragma unroll
                for (int k = 0; k < E; ++k) { tfx[k] = cx[k]; tfm[k] = cm[k]; }
                if (pending) {
#pragma unroll
                    for (int k = 0; k < E; ++k) prop[k] = cx[k];
                }
            }
            const bool second = pending && ts;  // a failed first leaf is returned unchanged (src/nuts.rs:858)
            {
                int ln; bool ls, kg; ST la;
                leaf_turn(cx, cm, cg, veps, second, logu, joint0, tfx, tfm, plus, ln, ls, la, kg);
                const bool take_b = draw_take_later(tn, ln, second);
                if (second) {
                    if (take_b) {
#pragma unroll
                        for (int k = 0; k < E; ++k) prop[k] = cx[k];
                    }
                    tn += ln;
                    ta = ta + la;
                    tna = 2;
                    ts = ls && kg;  // s' = s'_1 && s'_2 && stop_criterion(minus, plus), s'_1 = true here
                }
            }
            // ---- levels >= 1: merge at the set bits of the pair index, park at the first clear one
            bool pend = pending;
            for (int lvl = 1;; ++lvl) {
                if (!__any_sync(kFull, pend)) break;
                if (lvl == j) {  // merged through the top: the subtree of depth j is complete (or failed and fully unwound)
                    if (pend) building = false;
                    break;
                }
                if ((pr >> (lvl - 1)) & 1u) {
                    // merge the parked first half A = stack[lvl] with the later half T
                    const int an = s_n[lvl * NG + grp], ana = s_na[lvl * NG + grp];
                    const ST aa = (ST)s_a[lvl * NG + grp];
                    const bool take_b = draw_take_later(an, tn, pend);
                    {
                        float lprop[E];
                        load_level(lvl, 2, lprop);
                        if (pend && !take_b) {
#pragma unroll
                            for (int k = 0; k < E; ++k) prop[k] = lprop[k];
                        }
                    }
                    // the first leaf is only read by groups that are still merging (a group that has parked or
                    // finished never looks at tfx / tfm again), so it is loaded without a select
                    load_level(lvl, 0, tfx);
                    load_level(lvl, 1, tfm);
                    if (pend) {
                        tn += an;
                        ta = aa + ta;
                        tna += ana;
                    }
                    // s' = s'_1 && s'_2 && stop_criterion(minus, plus); parked halves always have s' = true
                    {
                        const bool kg = keep_going(cx, tfx, cm, tfm, plus);
                        if (pend && ts) ts = kg;
                    }
                } else {
                    // first half at this level: a valid subtree parks here and the group builds the next pair;
                    // a failed one is returned unchanged by the parent (src/nuts.rs:858) and keeps unwinding
                    __syncwarp();  // every lane has consumed the previous occupant's scalars (WAR)
                    if (pend && ts) {
                        store_level(lvl, 0, tfx);
                        store_level(lvl, 1, tfm);
                        store_level(lvl, 2, prop);
                        if (gl == 0) { s_n[lvl * NG + grp] = tn; s_na[lvl * NG + grp] = tna; s_a[lvl * NG + grp] = (double)ta; }
                    }
                    __syncwarp();
                    pend = pend && !ts;
                }
            }
            if (!__any_sync(kFull, building)) break;
        }
        if (run) { n_out = tn; s_out = ts; alpha_out = ta; nalpha_out = tna; }
    }
};

template <class Target, class A, class ST, int E, int G, bool kReplay>
__global__ void __launch_bounds__(kGrpWarps * 32, GrpTune<E>::kMinBlocks) nuts_group_kernel(const Target tgt, const NutsParams p) {
    extern __shared__ __align__(16) float nuts_smem[];
    using W = NutsGroup<Target, A, ST, E, G, kReplay>;
    constexpr int NG = W::NG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, grp = lane / G;
    const int64_t warp_slot = (int64_t)blockIdx.x * kGrpWarps + warp;
    const int n_glob = p.max_depth > W::kL ? p.max_depth - W::kL : 0;
    float *g_stack = p.scratch + warp_slot * (int64_t)n_glob * 3 * W::kVec;
    unsigned char *s_scal = reinterpret_cast<unsigned char *>(nuts_smem + kGrpWarps * W::kWarpFloats) + warp * W::kScalBytes;
    int *s_hist = reinterpret_cast<int *>(s_scal + kGrpMaxLevels * NG * 16);
    s_hist[lane] = 0;
    __syncwarp();
    W w(tgt, p, lane, nuts_smem + warp * W::kWarpFloats, g_stack, s_scal);
    unsigned long long n_trans = 0, tot_grad = 0, tot_unif = 0;

    // Work items: (group of NG chains, slice of the run).  Tickets are dispensed in order, every group's slice k before
    // any slice k + 1, so the warp that holds (g, k - 1) is always resident and a warp that draws (g, k) only has to
    // wait for its completion flag.  Slicing evens out the tail: with G = 8 a C5 shard of 8,192 chains is just 1.15
    // waves of whole-run tasks.
    const int64_t total = p.n_collect + p.n_discard;
    const int64_t first = (p.progress || p.resume) ? 0 : 1;
    const int64_t n_groups = (p.chains + NG - 1) / NG;
    const int64_t it_lo = p.it_lo > first ? p.it_lo : first;                        // this launch: iterations [it_lo, it_hi)
    const int64_t it_hi = (p.it_hi >= 0 && p.it_hi < total) ? p.it_hi : total;
    const int64_t n_iter = it_hi > it_lo ? it_hi - it_lo : 0;
    const int64_t slice_len = (p.slice_steps > 0 && p.slice_steps < n_iter) ? p.slice_steps : (n_iter > 0 ? n_iter : 1);
    const int64_t n_slices = n_iter > 0 ? (n_iter + slice_len - 1) / slice_len : 1;
    while (true) {
        long long ticket = 0;
        if (lane == 0) ticket = (long long)atomicAdd(&p.counters[0], 1ULL);
        ticket = __shfl_sync(kFull, ticket, 0);
        if (ticket >= n_groups * n_slices) break;
        const int64_t slice = ticket / n_groups, group = ticket - slice * n_groups;
        if (slice > 0) {
            if (lane == 0) {
                const volatile int *flag = p.flags + group;
                while (*flag < (int)slice) __nanosleep(256);
            }
            __syncwarp();
            __threadfence();  // acquire: the previous slice's positions / state (read with ld.cg below)
        }
        const long long c0 = group * NG;
        const bool has = c0 + grp < p.chains;
        // idle groups read the last entry and write nothing
        const long long ci = has ? c0 + grp : p.chains - 1;
        const long long c = p.perm ? (long long)p.perm[ci] : ci;
        w.chain = c;
        w.gchain = (uint64_t)(c + p.chain_offset);
        w.cur_n = w.cur_e = w.cur_u = 0;

        if constexpr (kReplay) {
            if (p.tree_scal) {  // build_tree debug mode: one doubling of depth tree_j per chain, src/nuts.rs:764-946
                float cx[E], cm[E], cg[E], prop[E], gp[E];
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    const int i = gl * E + W::off(k);
                    cx[k] = i < p.D ? p.positions[c * p.D + i] : 0.0f;
                    cm[k] = i < p.D ? p.tree_mom[c * p.D + i] : 0.0f;
                    cg[k] = i < p.D ? p.tree_grad[c * p.D + i] : 0.0f;
                    prop[k] = 0.0f;
                }
                const double *sc = p.tree_scal + c * 4;
                int n_prime = 0, n_alpha = 0;
                bool s_prime = false;
                ST alpha = (ST)0.0;
                w.doubling(cx, cm, cg, !(sc[1] < 0.0), p.tree_j, (ST)sc[0], (ST)sc[2], (ST)sc[3], has, prop, n_prime, s_prime,
                           alpha, n_alpha);
                const float lpp = w.full_logp(tgt.logp_grad(prop, gp, gl));
                if (has) {
                    float *ov = p.tree_out_vec + c * 5 * p.D;
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        const int i = gl * E + W::off(k);
                        if (i < p.D) {
                            ov[i] = cx[k]; ov[p.D + i] = cm[k]; ov[2 * p.D + i] = cg[k]; ov[3 * p.D + i] = prop[k];
                            ov[4 * p.D + i] = gp[k];
                        }
                    }
                    if (gl == 0) {
                        double *os = p.tree_out_scal + c * 6;
                        os[0] = (double)lpp; os[1] = (double)n_prime; os[2] = s_prime ? 1.0 : 0.0; os[3] = (double)alpha;
                        os[4] = (double)n_alpha; os[5] = (double)w.cur_u;
                    }
                }
                w.n_grad = 0; w.n_unif = 0;
                continue;
            }
        }
        double *st = p.state + c * 5;
        ST epsilon = (ST)__ldcg(st + 0), epsilon_bar = (ST)__ldcg(st + 1), h_bar = (ST)__ldcg(st + 2), mu;
        long long m = (long long)__ldcg(st + 4);
        const bool resume = p.resume || slice > 0 || it_lo > first;
        const ST gamma = (ST)0.05, kappa = (ST)0.75, delta = (ST)p.target_accept;
        const long long t_0 = 10;

        auto store_row = [&](float *o, const float (&v)[E]) {
            if (!has) return;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = gl * E + W::off(k);
                if (i < p.D) o[i] = v[k];
            }
        };

        // ---- init_chain, src/nuts.rs:528-545
        {
            float pos[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = gl * E + W::off(k);
                pos[k] = i < p.D ? __ldcg(p.positions + c * p.D + i) : 0.0f;
            }
            w.store_parked(3, pos);
            if (resume) {
                mu = (ST)__ldcg(st + 3);
            } else {
                if (p.n_collect > 0) store_row(p.out + (c * p.out_pitch) * p.D, pos);
                float m0[E];
                w.step_word = 0;
                w.q = 0; w.q_batch = 0xffffffffu;
                w.draw_normals(m0);
                ST d = epsilon + (ST)1.0;
                if (d < (ST)0.0) d = -d;
                const ST tiny = sizeof(ST) == 8 ? (ST)2.220446049250313e-16 : (ST)1.1920929e-07;
                const bool need = has && d <= tiny;
                if (__any_sync(kFull, need)) {
                    const ST found = w.find_reasonable_epsilon(pos, m0, need);
                    if (need) epsilon = found;
                }
                mu = s_log((ST)10.0 * epsilon);
            }
        }

        const int64_t it_begin = it_lo + slice * slice_len;
        const int64_t it_end = (n_iter > 0 && it_begin + slice_len < it_hi) ? it_begin + slice_len : it_hi;
        for (int64_t it = it_begin; it < it_end; ++it) {
            // ---- NUTSChain::step, src/nuts.rs:550-691
            m += 1;
            w.step_word = (uint32_t)m;
            w.q = 0; w.q_batch = 0xffffffffu;
            const ST eps_used = epsilon;
            const int64_t unifs_before = w.cur_u;
            float cx[E], cm[E], cg[E];  // the edge being extended; the opposite edge is parked in shared memory
            w.load_parked(3, cx);
            w.draw_normals(cm);
            float ulogp = tgt.logp_grad(cx, cg, gl);
            if (has) ++w.n_grad;
            const float ss0 = w.finish(ulogp, cm);
            const float joint_f = A::sub(ulogp, A::mul(ss0, 0.5f));
            const ST joint = (ST)(double)joint_f;
            const ST logu = joint - w.draw_exp1();
            w.store_parked(0, cx);
            w.store_parked(1, cm);
            w.store_parked(2, cg);
            int j = 0;          // warp-uniform: every group still in its transition has done j doublings
            int depth = 0;      // this group's number of doublings
            int n = 1;
            bool s = has;
            bool side = true;   // which edge the working registers hold (both edges coincide before the first doubling)
            ST alpha = (ST)0.0;
            int n_alpha = 0;
            while (__any_sync(kFull, s)) {
                const ST u1 = (ST)w.draw_uniform(false, s);
                const bool plus = u1 < (ST)0.5;
                if (j > 0) {  // change of direction: exchange the working edge with the parked one
                    const bool turn = s && plus != side;
                    if (__any_sync(kFull, turn)) {
                        w.swap_parked(0, cx, turn);
                        w.swap_parked(1, cm, turn);
                        w.swap_parked(2, cg, turn);
                    }
                }
                if (s) side = plus;
                float prop[E];
                int n_prime = 0;
                bool s_prime = false;
                w.doubling(cx, cm, cg, side, j, logu, epsilon, joint, s, prop, n_prime, s_prime, alpha, n_alpha);
                const ST ratio = (ST)n_prime / (ST)n;
                const ST tmp = ((ST)1.0 < ratio) ? (ST)1.0 : ratio;
                const ST u2 = (ST)w.draw_uniform(false, s);
                if (s && s_prime && (u2 < tmp)) w.store_parked(3, prop);
                bool kg;
                {
                    float ox[E], om[E];
                    w.load_parked(0, ox);
                    w.load_parked(1, om);
                    kg = w.keep_going(cx, ox, cm, om, side);
                }
                j += 1;
                if (s) {
                    n += n_prime;
                    depth = j;
                    s = s_prime && kg && j < p.max_depth;
                }
            }
            if (has && gl == 0) atomicAdd(&s_hist[depth < 31 ? depth : 31], 1);
            if (has && gl == 0) ++n_trans;
            if (p.trace && has && gl == 0) {
                double *tr = p.trace + (c * p.trace_pitch + it) * 8;
                tr[0] = (double)joint; tr[1] = (double)logu; tr[2] = (double)n; tr[3] = (double)alpha; tr[4] = (double)n_alpha;
                tr[5] = (double)depth; tr[6] = (double)eps_used; tr[7] = kReplay ? (double)(w.cur_u - unifs_before) : (double)w.q;
            }
            // dual averaging, src/nuts.rs:676-690
            ST eta = (ST)1.0 / (ST)(m + t_0);
            h_bar = ((ST)1.0 - eta) * h_bar + eta * (delta - alpha / (ST)n_alpha);
            if (m <= p.adapt_until) {
                const ST _m = (ST)m;
                epsilon = s_exp(mu - s_sqrt(_m) / gamma * h_bar);
                eta = s_pow(_m, -kappa);
                epsilon_bar = s_exp(((ST)1.0 - eta) * s_log(epsilon_bar) + eta * s_log(epsilon));
            } else {
                epsilon = epsilon_bar;
            }
            if (it >= p.n_discard) {
                float pos[E];
                w.load_parked(3, pos);
                store_row(p.out + (c * p.out_pitch + (it - p.n_discard)) * p.D, pos);
            }
        }
        {
            float pos[E];
            w.load_parked(3, pos);
            store_row(p.positions + c * p.D, pos);
        }
        if (has && gl == 0) {
            st[0] = (double)epsilon; st[1] = (double)epsilon_bar; st[2] = (double)h_bar; st[3] = (double)mu;
            st[4] = (double)m;
            tot_grad += w.n_grad; tot_unif += w.n_unif;
        }
        w.n_grad = 0; w.n_unif = 0;
        if (n_slices > 1) {  // release: publish the slice
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(p.flags + group, (int)slice + 1);
        }
    }
    if (gl == 0) {
        if (tot_grad) atomicAdd(&p.counters[1], tot_grad);
        if (n_trans) atomicAdd(&p.counters[2], n_trans);
        if (tot_unif) atomicAdd(&p.counters[3], tot_unif);
    }
    __syncwarp();
    if (s_hist[lane]) atomicAdd(&p.counters[8 + lane], (unsigned long long)s_hist[lane]);
}

}  // namespace mmc
