// K4b: NUTS with several chains per warp ("group" layout; see include/minimcmc.h "NUTS", mmc_nuts_set_layout).
//
// Same algorithm, RNG contract and outputs as the one-chain-per-warp kernel of mmc_nuts.cuh
// (NUTSChain::{init_chain, step, run, run_progress} src/nuts.rs:457-691, build_tree :764-946, leapfrog :979-996,
// stop_criterion :963-977, find_reasonable_epsilon :695-761), but a warp carries 32 / G chains: G lanes per chain,
// E consecutive vector elements per lane (G E >= D).  The tree bookkeeping of the warp kernel (merge walk, level
// addressing, exp, uniform draws; about two thirds of its ~300 warp-instructions per leapfrog at D = 100) is issued
// once per warp and therefore shared by 32 / G chains, the reductions are log2 G butterfly steps instead of five, and
// with E = 13 at D = 100 all but 4 of the 104 element slots carry data (the warp kernel: 100 of 128).
//
// The chains of a warp advance in lock step, one transition at a time.  Control flow is warp-uniform (loops run
// while ANY group of the warp still needs them, decided by votes) and every state update is predicated with the
// group's own flags, so the full-mask shuffles are always executed convergently:
//   * a group whose transition has ended (U-turn, divergence, max depth) idles until the others end theirs;
//   * inside a doubling a group that is not building (finished, or its subtree failed) integrates with step size 0,
//     which leaves (x, p) untouched, and none of its results are committed;
//   * the binary-counter merge walk visits level l for every leaf: groups still carrying a valid subtree merge at the
//     set bits of the leaf index and park at the first clear bit; a failed subtree keeps merging at all set bits and
//     passes through the clear ones - exactly the RNG consumption and alpha / n_alpha sums of the recursion.
// Direction: the edge being extended is swapped into the "plus" registers for the doubling (the U-turn test
// (x+ - x-).p- >= 0 && (x+ - x-).p+ >= 0 is symmetric in the two momenta, and -(a - b) == b - a exactly).
#pragma once

#include "mmc_nuts.cuh"

namespace mmc {

#ifndef MMC_NUTS_GROUP_MIN_BLOCKS
#define MMC_NUTS_GROUP_MIN_BLOCKS 2   // <= 255 registers at E = 13 (ten E-vectors live in registers)
#endif
constexpr int kGrpWarps = 4;
constexpr int kGrpSmemLevels = 3;
constexpr int kGrpMaxLevels = 16;

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

// ---------------------------------------------------------------- group-form targets
// interface: float logp_grad(const float (&x)[E], float (&g)[E], int gl) const; gl = lane inside the group, which
// owns elements gl*E .. gl*E+E-1 (zero padded).  kPartial as in the warp form.

// RosenbrockND (src/distributions.rs:531-547) for D <= G E.
template <class A, int E, int G>
struct GRosenbrockND {
    static constexpr bool kPartial = true;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        const float xn = __shfl_down_sync(kFull, x[0], 1, G);
        float t[E];
        float acc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = gl * E + e;
            const bool valid = i + 1 < D;
            const float xnext = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : xn;
            const float tt = valid ? cms<A>(xnext, x[e], x[e]) : 0.0f;
            const float u = valid ? A::sub(1.0f, x[e]) : 0.0f;
            t[e] = tt;
            acc = A::add(acc, A::mad(A::mul(tt, tt), 100.0f, A::mul(u, u)));
            g[e] = A::mad(A::mul(400.0f, x[e]), tt, A::mul(2.0f, u));
        }
        float tprev = __shfl_up_sync(kFull, t[E - 1], 1, G);
        if (gl == 0) tprev = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float tp = e == 0 ? tprev : t[e == 0 ? 0 : e - 1];
            g[e] = A::add(A::mul(-200.0f, tp), g[e]);
        }
        return -acc;
    }
};

// StdNormal (src/nuts.rs:1024-1037); padding elements are zero.
template <class A, int E, int G>
struct GStdNormal {
    static constexpr bool kPartial = true;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int gl) const {
        float acc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc = A::mad(A::mul(x[e], x[e]), 0.5f, acc);
            g[e] = -x[e];
        }
        return -acc;
    }
};

// Any small thread-form target (kDim <= G E): every lane of the group gathers the full vector and evaluates it.
template <class T, int E, int G>
struct GSmall {
    static constexpr bool kPartial = false;
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
    static constexpr int Q4 = (E + 3) / 4;        // float4 per lane and vector in the level stacks
    static constexpr int kVec = Q4 * 4 * 32;      // floats of one parked vector of the whole warp
    static constexpr int kScalBytes = kGrpMaxLevels * NG * 16 + 32 * 4;  // per warp: (double alpha, int n, int n_alpha) per level and group + depth histogram

    const Target &tgt;
    const NutsParams &p;
    const int lane, gl, grp;
    float *s_stack;      // shared: [kGrpSmemLevels][3][Q4][32] float4 of this warp; every lane only touches its own column
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

    __device__ NutsGroup(const Target &t, const NutsParams &pp, int ln, float *ss, float *gs, void *scal)
        : tgt(t), p(pp), lane(ln), gl(ln % G), grp(ln / G), s_stack(ss), g_stack(gs) {
        s_a = reinterpret_cast<double *>(scal);
        s_n = reinterpret_cast<int *>(s_a + kGrpMaxLevels * NG);
        s_na = s_n + kGrpMaxLevels * NG;
        ubatch = make_uint4(0, 0, 0, 0);
    }

    // ---- parked subtree vectors (which: 0 = first-leaf x, 1 = first-leaf p, 2 = proposal)
    __device__ __forceinline__ void load_level(int lvl, int which, float (&v)[E]) {
        if (lvl < kGrpSmemLevels) {
            const float4 *src = reinterpret_cast<const float4 *>(s_stack + (lvl * 3 + which) * kVec) + lane;
#pragma unroll
            for (int k4 = 0; k4 < Q4; ++k4) {
                const float4 t = src[k4 * 32];
                if (4 * k4 + 0 < E) v[(4 * k4 + 0) % E] = t.x;
                if (4 * k4 + 1 < E) v[(4 * k4 + 1) % E] = t.y;
                if (4 * k4 + 2 < E) v[(4 * k4 + 2) % E] = t.z;
                if (4 * k4 + 3 < E) v[(4 * k4 + 3) % E] = t.w;
            }
        } else {
            const float4 *src = reinterpret_cast<const float4 *>(g_stack + ((lvl - kGrpSmemLevels) * 3 + which) * kVec) + lane;
#pragma unroll
            for (int k4 = 0; k4 < Q4; ++k4) {
                const float4 t = __ldcg(src + k4 * 32);
                if (4 * k4 + 0 < E) v[(4 * k4 + 0) % E] = t.x;
                if (4 * k4 + 1 < E) v[(4 * k4 + 1) % E] = t.y;
                if (4 * k4 + 2 < E) v[(4 * k4 + 2) % E] = t.z;
                if (4 * k4 + 3 < E) v[(4 * k4 + 3) % E] = t.w;
            }
        }
    }
    __device__ __forceinline__ void store_level(int lvl, int which, const float (&v)[E]) {
        float4 t[Q4];
#pragma unroll
        for (int k4 = 0; k4 < Q4; ++k4)
            t[k4] = make_float4(v[(4 * k4 + 0) % E], (4 * k4 + 1 < E) ? v[(4 * k4 + 1) % E] : 0.0f,
                                (4 * k4 + 2 < E) ? v[(4 * k4 + 2) % E] : 0.0f, (4 * k4 + 3 < E) ? v[(4 * k4 + 3) % E] : 0.0f);
        if (lvl < kGrpSmemLevels) {
            float4 *dst = reinterpret_cast<float4 *>(s_stack + (lvl * 3 + which) * kVec) + lane;
#pragma unroll
            for (int k4 = 0; k4 < Q4; ++k4) dst[k4 * 32] = t[k4];
        } else {
            float4 *dst = reinterpret_cast<float4 *>(g_stack + ((lvl - kGrpSmemLevels) * 3 + which) * kVec) + lane;
#pragma unroll
            for (int k4 = 0; k4 < Q4; ++k4) __stcg(dst + k4 * 32, t[k4]);
        }
    }

    // ---- random draws: the counters of the warp kernel (minimcmc.h "RNG contract"): normal i of a step comes from
    // word i & 3 of Philox block i >> 2, uniform q from words (q & 1 ? zw : xy) of block kSubUnif + (q >> 1)
    __device__ __forceinline__ void draw_normals(float (&m)[E]) {
        if (kReplay) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = gl * E + e;
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
            for (int e = 0; e < E; ++e) {
                float v = nb[e];
                if (E % 4 != 0) {
                    if (r == 1) v = nb[(e + 1) % (NQ * 4)];
                    if (r == 2) v = nb[(e + 2) % (NQ * 4)];
                    if (r == 3) v = nb[(e + 3) % (NQ * 4)];
                }
                m[e] = (i0 + e < p.D) ? v : 0.0f;
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
            const uint4 nb = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word,
                                                            kSubUnif + batch * G + (uint32_t)gl));
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
    __device__ __forceinline__ uint64_t draw_u53(bool take) {  // native mode only
        uint32_t lo, hw;
        next_words(take, lo, hw);
        return (((uint64_t)hw << 32) | lo) >> 11;
    }

    // ---- leapfrog, src/nuts.rs:979-996 (in place); returns logp' (partial when Target::kPartial)
    __device__ __forceinline__ float leapfrog(float (&x)[E], float (&m)[E], float (&g)[E], float e) {
#pragma unroll
        for (int k = 0; k < E; ++k) {
            m[k] = A::mad(A::mul(g[k], e), 0.5f, m[k]);
            x[k] = A::mad(m[k], e, x[k]);
        }
        const float lp = tgt.logp_grad(x, g, gl);
#pragma unroll
        for (int k = 0; k < E; ++k) m[k] = A::mad(A::mul(g[k], e), 0.5f, m[k]);
        return lp;
    }
    __device__ __forceinline__ float sumsq(const float (&m)[E]) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) s = A::mad(m[k], m[k], s);
        return group_sum<A, G>(s);
    }
    __device__ __forceinline__ float finish(float &lp_io, const float (&m)[E]) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) s = A::mad(m[k], m[k], s);
        if (Target::kPartial) group_sum2<A, G>(lp_io, s);
        else s = group_sum<A, G>(s);
        return s;
    }
    __device__ __forceinline__ float full_logp(float lp) { return Target::kPartial ? group_sum<A, G>(lp) : lp; }
    // stop_criterion, src/nuts.rs:963-977 (true = keep going) between the edge being extended (xw, pw) and another
    // state (xo, po); plus: the extended edge is the plus side
    __device__ __forceinline__ bool keep_going(const float (&xw)[E], const float (&xo)[E], const float (&pw)[E],
                                               const float (&po)[E], bool plus) {
        float dw = 0.0f, dt = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            float diff = A::sub(xw[k], xo[k]);
            diff = plus ? diff : -diff;
            dw = A::mad(diff, pw[k], dw);
            dt = A::mad(diff, po[k], dt);
        }
        group_sum2<A, G>(dw, dt);
        return dw >= 0.0f && dt >= 0.0f;
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

    // One doubling = build_tree(edge, v, j), src/nuts.rs:764-946, iteratively, for the groups with `run`.
    // cx/cm/cg: the edge to extend (in) and the new edge (out).  For the `run` groups the outputs are the subtree
    // proposal (prop), n', s', alpha and n_alpha; other groups keep their scalars and never read their prop.
    __device__ void doubling(float (&cx)[E], float (&cm)[E], float (&cg)[E], bool plus, int j, ST logu, ST eps, ST joint0,
                             bool run, float (&prop)[E], int &n_out, bool &s_out, ST &alpha_out, int &nalpha_out) {
        const float veps = (float)(plus ? eps : -eps);
        const uint32_t n_leaves = 1u << j;
        float tfx[E], tfm[E];  // first leaf of the subtree currently being merged upward (its proposal: `prop`)
#pragma unroll
        for (int k = 0; k < E; ++k) { tfx[k] = cx[k]; tfm[k] = cm[k]; }
        int tn = 0, tna = 0;
        ST ta = (ST)0.0;
        bool ts = true;
        bool building = run;
        for (uint32_t leaf = 0; leaf < n_leaves; ++leaf) {
            float lp = leapfrog(cx, cm, cg, building ? veps : 0.0f);
            const float ss = finish(lp, cm);
            const float joint_f = A::sub(lp, A::mul(ss, 0.5f));
            const ST joint = (ST)(double)joint_f;
            const ST ex = s_exp(joint - joint0);
            if (building) {
                ++n_grad;
                tn = (logu < joint) ? 1 : 0;
                ts = (logu - (ST)1000.0) < joint;
                ta = ((ST)1.0 < ex || ex != ex) ? (ST)1.0 : ex;  // T::min(1, e): NaN -> 1
                tna = 1;
#pragma unroll
                for (int k = 0; k < E; ++k) { tfx[k] = cx[k]; tfm[k] = cm[k]; prop[k] = cx[k]; }
            }
            bool pending = building;  // this group still has to merge / park the subtree that ends in this leaf
            for (int lvl = 0;; ++lvl) {
                if (!__any_sync(kFull, pending)) break;
                if (lvl == j) {  // merged through the top: the subtree of depth j is complete (or failed and fully unwound)
                    if (pending) building = false;
                    break;
                }
                if ((leaf >> lvl) & 1u) {
                    // merge the parked first half A = stack[lvl] with the later half T
                    const int an = s_n[lvl * NG + grp], ana = s_na[lvl * NG + grp];
                    const ST aa = (ST)s_a[lvl * NG + grp];
                    // u < n'' / max(n' + n'', 1) in f64 (src/nuts.rs:910-911)
                    bool take_b;
                    if (kReplay) {
                        const double u = draw_uniform(true, pending);
                        if (tn == 0) take_b = false;
                        else if (an == 0) take_b = u < 1.0;
                        else take_b = u < ((double)tn / (double)(an + tn));
                    } else {
                        // native draws are k 2^-53: k (n' + n'') < n'' 2^53 is the same test evaluated exactly
                        const uint64_t k53 = draw_u53(pending);
                        take_b = tn != 0 && (an == 0 || k53 * (uint64_t)(an + tn) < ((uint64_t)tn << 53));
                    }
                    float lx[E], lm[E];
                    load_level(lvl, 0, lx);
                    load_level(lvl, 1, lm);
                    if (__any_sync(kFull, pending && !take_b)) {
                        float lprop[E];
                        load_level(lvl, 2, lprop);
                        if (pending && !take_b) {
#pragma unroll
                            for (int k = 0; k < E; ++k) prop[k] = lprop[k];
                        }
                    }
                    if (pending) {
#pragma unroll
                        for (int k = 0; k < E; ++k) { tfx[k] = lx[k]; tfm[k] = lm[k]; }
                        tn += an;
                        ta = aa + ta;
                        tna += ana;
                    }
                    // s' = s'_1 && s'_2 && stop_criterion(minus, plus); parked halves always have s' = true
                    if (__any_sync(kFull, pending && ts)) {
                        const bool kg = keep_going(cx, tfx, cm, tfm, plus);
                        if (pending && ts) ts = kg;
                    }
                } else {
                    // first half at this level: a valid subtree parks here and the group builds the next leaf;
                    // a failed one is returned unchanged by the parent (src/nuts.rs:858) and keeps unwinding
                    __syncwarp();  // every lane has consumed the previous occupant's scalars (WAR)
                    if (pending && ts) {
                        store_level(lvl, 0, tfx);
                        store_level(lvl, 1, tfm);
                        store_level(lvl, 2, prop);
                        if (gl == 0) { s_n[lvl * NG + grp] = tn; s_na[lvl * NG + grp] = tna; s_a[lvl * NG + grp] = (double)ta; }
                    }
                    __syncwarp();
                    pending = pending && !ts;
                }
            }
            if (!__any_sync(kFull, building)) break;
        }
        if (run) { n_out = tn; s_out = ts; alpha_out = ta; nalpha_out = tna; }
    }
};

template <class Target, class A, class ST, int E, int G, bool kReplay>
__global__ void __launch_bounds__(kGrpWarps * 32, MMC_NUTS_GROUP_MIN_BLOCKS) nuts_group_kernel(const Target tgt, const NutsParams p) {
    extern __shared__ __align__(16) float nuts_smem[];
    using W = NutsGroup<Target, A, ST, E, G, kReplay>;
    constexpr int NG = W::NG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, grp = lane / G;
    float *s_stack = nuts_smem + warp * kGrpSmemLevels * 3 * W::kVec;
    const int64_t warp_slot = (int64_t)blockIdx.x * kGrpWarps + warp;
    const int n_glob = p.max_depth > kGrpSmemLevels ? p.max_depth - kGrpSmemLevels : 0;
    float *g_stack = p.scratch + warp_slot * (int64_t)n_glob * 3 * W::kVec;
    unsigned char *s_scal = reinterpret_cast<unsigned char *>(nuts_smem + kGrpWarps * kGrpSmemLevels * 3 * W::kVec) + warp * W::kScalBytes;
    int *s_hist = reinterpret_cast<int *>(s_scal + kGrpMaxLevels * NG * 16);
    s_hist[lane] = 0;
    __syncwarp();
    W w(tgt, p, lane, s_stack, g_stack, s_scal);
    unsigned long long n_trans = 0, tot_grad = 0, tot_unif = 0;

    while (true) {
        long long c0 = 0;
        if (lane == 0) c0 = (long long)atomicAdd(&p.counters[0], (unsigned long long)NG);
        c0 = __shfl_sync(kFull, c0, 0);
        if (c0 >= p.chains) break;
        const bool has = c0 + grp < p.chains;
        const long long c = has ? c0 + grp : p.chains - 1;  // idle groups read the last chain and write nothing
        w.chain = c;
        w.gchain = (uint64_t)(c + p.chain_offset);
        w.cur_n = w.cur_e = w.cur_u = 0;

        float pos[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int i = gl * E + k;
            pos[k] = i < p.D ? p.positions[c * p.D + i] : 0.0f;
        }
        double *st = p.state + c * 5;
        ST epsilon = (ST)st[0], epsilon_bar = (ST)st[1], h_bar = (ST)st[2], mu;
        long long m = (long long)st[4];
        const ST gamma = (ST)0.05, kappa = (ST)0.75, delta = (ST)p.target_accept;
        const long long t_0 = 10;

        auto store_draw = [&](int64_t slot) {
            if (!has) return;
            float *o = p.out + (c * p.out_pitch + slot) * p.D;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = gl * E + k;
                if (i < p.D) o[i] = pos[k];
            }
        };

        // ---- init_chain, src/nuts.rs:528-545
        if (p.resume) {
            mu = (ST)st[3];
        } else {
            if (p.n_collect > 0) store_draw(0);
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

        const int64_t total = p.n_collect + p.n_discard;
        const int64_t first = (p.progress || p.resume) ? 0 : 1;
        for (int64_t it = first; it < total; ++it) {
            // ---- NUTSChain::step, src/nuts.rs:550-691
            m += 1;
            w.step_word = (uint32_t)m;
            w.q = 0; w.q_batch = 0xffffffffu;
            float xm[E], pm[E], gm[E], xp[E], pp[E], gp[E];
            w.draw_normals(pp);
            float ulogp = tgt.logp_grad(pos, gp, gl);
            if (has) ++w.n_grad;
            const float ss0 = w.finish(ulogp, pp);
            const float joint_f = A::sub(ulogp, A::mul(ss0, 0.5f));
            const ST joint = (ST)(double)joint_f;
            const ST logu = joint - w.draw_exp1();
#pragma unroll
            for (int k = 0; k < E; ++k) {
                xm[k] = xp[k] = pos[k];
                pm[k] = pp[k];
                gm[k] = gp[k];
            }
            int j = 0;          // warp-uniform: every group still in its transition has done j doublings
            int depth = 0;      // this group's number of doublings
            int n = 1;
            bool s = has;
            ST alpha = (ST)0.0;
            int n_alpha = 0;
            while (__any_sync(kFull, s)) {
                const ST u1 = (ST)w.draw_uniform(false, s);
                const bool plus = u1 < (ST)0.5;
                if (!plus) {  // extend the minus edge: bring it into the working registers
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        float t;
                        t = xm[k]; xm[k] = xp[k]; xp[k] = t;
                        t = pm[k]; pm[k] = pp[k]; pp[k] = t;
                        t = gm[k]; gm[k] = gp[k]; gp[k] = t;
                    }
                }
                float prop[E];
#pragma unroll
                for (int k = 0; k < E; ++k) prop[k] = pos[k];
                int n_prime = 0;
                bool s_prime = false;
                w.doubling(xp, pp, gp, plus, j, logu, epsilon, joint, s, prop, n_prime, s_prime, alpha, n_alpha);
                const ST ratio = (ST)n_prime / (ST)n;
                const ST tmp = ((ST)1.0 < ratio) ? (ST)1.0 : ratio;
                const ST u2 = (ST)w.draw_uniform(false, s);
                if (s && s_prime && (u2 < tmp)) {
#pragma unroll
                    for (int k = 0; k < E; ++k) pos[k] = prop[k];
                }
                const bool kg = w.keep_going(xp, xm, pp, pm, plus);
                if (!plus) {
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        float t;
                        t = xm[k]; xm[k] = xp[k]; xp[k] = t;
                        t = pm[k]; pm[k] = pp[k]; pp[k] = t;
                        t = gm[k]; gm[k] = gp[k]; gp[k] = t;
                    }
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
            if (it >= p.n_discard) store_draw(it - p.n_discard);
        }
        if (has) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = gl * E + k;
                if (i < p.D) p.positions[c * p.D + i] = pos[k];
            }
            if (gl == 0) {
                st[0] = (double)epsilon; st[1] = (double)epsilon_bar; st[2] = (double)h_bar; st[3] = (double)mu;
                st[4] = (double)m;
                tot_grad += w.n_grad; tot_unif += w.n_unif;
            }
        }
        w.n_grad = 0; w.n_unif = 0;
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
