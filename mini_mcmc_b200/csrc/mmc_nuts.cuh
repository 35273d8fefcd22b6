// K4: NUTS, one chain per warp (see include/minimcmc.h "NUTS").
//
// Reproduces NUTSChain::{init_chain, step, run, run_progress} (src/nuts.rs:457-691), build_tree (:764-946),
// leapfrog (:979-996), stop_criterion (:963-977) and find_reasonable_epsilon (:695-761).
//
// Layout: the D-vector of a chain is blocked over the 32 lanes (lane l owns elements l*E .. l*E+E-1,
// zero padded), so a leapfrog is E independent FMAs per lane plus two neighbour shuffles for the
// Rosenbrock stencil, and every dot product (kinetic energy, U-turn test) is a 5-step butterfly.
// Accept / U-turn / divergence decisions are warp-uniform, so the whole tree walk is divergence free
// inside a warp; different chains (warps) take different numbers of leapfrogs without waiting for each
// other (persistent warps pull chains from an atomic counter).
//
// build_tree is recursive in the reference; here it is the equivalent iterative binary-counter walk
// (SURVEY.md appendix B1): after leaf #k the pending subtrees are merged once per trailing 1-bit of k,
// drawing one f64 uniform per merge; a failed subtree (divergence or U-turn) still merges with every
// pending sibling above it (set bits of k) and passes through unchanged where it is a first half, which is
// exactly the RNG consumption and alpha / n_alpha accumulation of the recursion.  Only O(depth) states are
// kept: per level the first leaf (x, p) and the current proposal x'.  Levels < kSmemLevels live in shared
// memory, deeper (exponentially rarer) levels in an L2-resident per-warp scratch.
#pragma once

#include "mmc_common.cuh"
#include "mmc_targets.cuh"

namespace mmc {

#ifndef MMC_NUTS_MIN_BLOCKS
#define MMC_NUTS_MIN_BLOCKS 5   // 96 registers: 20 warps / SM (measured best on B200: 4 -> 544 ms, 5 -> 497 ms, 6 -> 514 ms at C5)
#endif
constexpr int kNutsSmemLevels = 3;
constexpr int kNutsWarps = 4;
constexpr unsigned kFull = 0xffffffffu;

template <class A>
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = A::add(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// two independent sums in one butterfly (the shuffles of a step pipeline instead of serialising)
template <class A>
__device__ __forceinline__ void warp_sum2(float &a, float &b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ta = __shfl_xor_sync(kFull, a, o);
        const float tb = __shfl_xor_sync(kFull, b, o);
        a = A::add(a, ta);
        b = A::add(b, tb);
    }
}

// ---------------------------------------------------------------- warp-form targets
// interface: float logp_grad(const float (&x)[E], float (&g)[E], int lane) const  -> logp (same in all lanes)
//            kPartial = true: the returned value is this lane's partial sum of logp (the caller reduces it
//            together with the kinetic energy in one butterfly); false: already the full logp in every lane.

// RosenbrockND (src/distributions.rs:531-547) for any D <= 32 E.
template <class A, int E>
struct WRosenbrockND {
    static constexpr bool kPartial = true;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int lane) const {
        const float xn = __shfl_down_sync(kFull, x[0], 1);
        float t[E];
        float acc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = lane * E + e;
            const bool valid = i + 1 < D;
            const float xnext = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : xn;
            const float tt = valid ? cms<A>(xnext, x[e], x[e]) : 0.0f;
            const float u = valid ? A::sub(1.0f, x[e]) : 0.0f;
            t[e] = tt;
            acc = A::add(acc, A::mad(A::mul(tt, tt), 100.0f, A::mul(u, u)));
            g[e] = A::mad(A::mul(400.0f, x[e]), tt, A::mul(2.0f, u));
        }
        float tprev = __shfl_up_sync(kFull, t[E - 1], 1);
        if (lane == 0) tprev = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float tp = e == 0 ? tprev : t[e == 0 ? 0 : e - 1];
            g[e] = A::add(A::mul(-200.0f, tp), g[e]);
        }
        return -acc;
    }
};

// StdNormal (src/nuts.rs:1024-1037) for any D <= 32 E (padding elements are zero).
template <class A, int E>
struct WStdNormal {
    static constexpr bool kPartial = true;
    int D;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int lane) const {
        float acc = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            acc = A::mad(A::mul(x[e], x[e]), 0.5f, acc);
            g[e] = -x[e];
        }
        return -acc;
    }
};

// Any small thread-form target (kDim <= 4): every lane gathers the full vector and evaluates it.
template <class T, int E>
struct WSmall {
    static constexpr bool kPartial = false;
    T t;
    __device__ __forceinline__ float logp_grad(const float (&x)[E], float (&g)[E], int lane) const {
        constexpr int K = T::kDim;
        float xf[K], gf[K];
#pragma unroll
        for (int i = 0; i < K; ++i) xf[i] = __shfl_sync(kFull, x[i % E], i / E);
        const float lp = t.logp_grad(xf, gf);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            g[e] = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (lane * E + e == i) g[e] = gf[i];
        }
        return lp;
    }
};

// k 2^-53 < num / den evaluated exactly for a 53-bit k: k den < num 2^53 in 128 bits (k den needs 53 + log2 den bits,
// which passes 64 from tree depth 11 on; max_depth goes up to 16)
__device__ __forceinline__ bool u53_below_ratio(uint64_t k53, uint32_t num, uint32_t den) {
    const uint64_t lo = k53 * (uint64_t)den, hi = __umul64hi(k53, (uint64_t)den);
    const uint64_t rlo = (uint64_t)num << 53, rhi = (uint64_t)num >> 11;
    return hi < rhi || (hi == rhi && lo < rlo);
}

// ---------------------------------------------------------------- scalar helpers (type T of the reference)
__device__ __forceinline__ float s_exp(float v) { return expf(v); }
__device__ __forceinline__ double s_exp(double v) { return exp(v); }
__device__ __forceinline__ float s_log(float v) { return logf(v); }
__device__ __forceinline__ double s_log(double v) { return log(v); }
__device__ __forceinline__ float s_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double s_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float s_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double s_pow(double a, double b) { return pow(a, b); }

struct NutsParams {
    float *positions;       // [chains, D] in/out
    float *out;             // [chains, n_collect, D]
    double *state;          // [chains, 5] = epsilon, epsilon_bar, h_bar, mu, m   in/out
    const double *normals, *exps, *unifs;  // replay tapes [chains, cap_*]
    int64_t cap_normals, cap_exps, cap_unifs;
    float *scratch;         // [resident warps][max_depth - kNutsSmemLevels][3][32 E]
    unsigned long long *counters;  // [0] next chain, [1] n_grad, [2] n_transitions, [3] uniforms consumed, [8..] depth histogram
    int64_t chains, chain_offset;
    int64_t n_collect, n_discard;
    int64_t out_pitch;      // draws per chain row of `out` (>= n_collect)
    int64_t adapt_until;    // dual averaging adapts while m <= adapt_until (= n_discard in the reference, src/nuts.rs:681)
    int32_t progress, max_depth, D;
    int32_t resume;         // 1: continue a run split over several launches (skip init_chain, keep mu)
    // group kernel only: a group's run is cut into slices of slice_steps iterations that are dispensed by the ticket
    // counter (all groups' slice 0 first); flags[group] = number of completed slices (0 = slicing off)
    int64_t slice_steps;
    int *flags;
    // group kernel only: this launch covers iterations [it_lo, it_hi) of the run (it_hi < 0: to the end) and takes the
    // chains of its warps from perm (groups of similar step size, see mmc_nuts.cu; nullptr = index order)
    int64_t it_lo, it_hi;
    const int *perm;
    double target_accept;
    uint2 key;
    // optional per-transition trace [chains, trace_pitch, 8] = joint_0, logu, n, alpha, n_alpha, depth, epsilon used,
    // uniforms consumed; row = iteration index of the launch (mmc_nuts_set_trace_dev)
    double *trace;
    int64_t trace_pitch;
    // build_tree debug mode (mmc_nuts_build_tree; replay instantiations only): one doubling of depth tree_j per chain
    // from (positions, tree_mom, tree_grad) with tree_scal [chains, 4] = logu, v, epsilon, joint_0 and the uniforms of the
    // replay tape; tree_out_vec [chains, 5, D] = new edge x, p, grad, proposal x', grad(x'); tree_out_scal [chains, 6] =
    // logp(x'), n', s', alpha', n_alpha', uniforms consumed
    const float *tree_mom, *tree_grad;
    const double *tree_scal;
    float *tree_out_vec;
    double *tree_out_scal;
    int32_t tree_j;
};

template <class Target, class A, class ST, int E, bool kReplay>
struct NutsWarp {
    const Target &tgt;
    const NutsParams &p;
    const int lane;
    float *s_stack;      // shared: [kNutsSmemLevels][3][32 E] for this warp
    float *g_stack;      // global scratch for deeper levels
    // RNG state
    uint64_t gchain = 0;
    int64_t chain = 0;
    uint32_t step_word = 0;
    uint32_t q = 0;           // uniforms consumed in this step
    uint32_t q_batch = 0xffffffffu;
    uint4 ubatch;
    int64_t cur_n = 0, cur_e = 0, cur_u = 0;  // replay cursors
    uint32_t n_grad = 0, n_unif = 0;      // per chain, flushed into 64-bit totals by the caller
    // per-level scalars of the pending halves live in shared memory (warp-uniform broadcast reads)
    int *s_n, *s_na;
    double *s_a;

    __device__ NutsWarp(const Target &t, const NutsParams &pp, int ln, float *ss, float *gs, void *scal)
        : tgt(t), p(pp), lane(ln), s_stack(ss), g_stack(gs) {
        s_a = reinterpret_cast<double *>(scal);
        s_n = reinterpret_cast<int *>(s_a + 16);
        s_na = s_n + 16;
    }

    // pending-subtree vectors (which: 0 = first-leaf x, 1 = first-leaf p, 2 = proposal): explicit shared / global
    // paths so the compiler emits LDS/STS and LDG/STG (vectorised when E == 4) instead of generic accesses
    __device__ __forceinline__ void load_level(int lvl, int which, float (&v)[E]) {
        constexpr int V = 32 * E;
        if (lvl < kNutsSmemLevels) {
            const float *src = s_stack + (lvl * 3 + which) * V + lane * E;
            if (E == 4) { const float4 t = *reinterpret_cast<const float4 *>(src); v[0] = t.x; v[1 % E] = t.y; v[2 % E] = t.z; v[3 % E] = t.w; }
            else {
#pragma unroll
                for (int k = 0; k < E; ++k) v[k] = src[k];
            }
        } else {
            const float *src = g_stack + ((lvl - kNutsSmemLevels) * 3 + which) * V + lane * E;
            if (E == 4) { const float4 t = __ldcg(reinterpret_cast<const float4 *>(src)); v[0] = t.x; v[1 % E] = t.y; v[2 % E] = t.z; v[3 % E] = t.w; }
            else {
#pragma unroll
                for (int k = 0; k < E; ++k) v[k] = __ldcg(src + k);
            }
        }
    }
    __device__ __forceinline__ void store_level(int lvl, int which, const float (&v)[E]) {
        constexpr int V = 32 * E;
        if (lvl < kNutsSmemLevels) {
            float *dst = s_stack + (lvl * 3 + which) * V + lane * E;
            if (E == 4) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1 % E], v[2 % E], v[3 % E]);
            else {
#pragma unroll
                for (int k = 0; k < E; ++k) dst[k] = v[k];
            }
        } else {
            float *dst = g_stack + ((lvl - kNutsSmemLevels) * 3 + which) * V + lane * E;
            if (E == 4) __stcg(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1 % E], v[2 % E], v[3 % E]));
            else {
#pragma unroll
                for (int k = 0; k < E; ++k) __stcg(dst + k, v[k]);
            }
        }
    }

    // ---- random draws (native Philox keying is documented in minimcmc.h)
    __device__ __forceinline__ void draw_normals(float (&m)[E]) {
        if (kReplay) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = lane * E + e;
                m[e] = i < p.D ? (float)p.normals[chain * p.cap_normals + cur_n + i] : 0.0f;
            }
            cur_n += p.D;
        } else {
            if (E == 4) {
                const uint4 w = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word, (uint32_t)lane));
                float n[4];
                box_muller_f32(w.x, w.y, n[0], n[1]);
                box_muller_f32(w.z, w.w, n[2], n[3]);
#pragma unroll
                for (int e = 0; e < E; ++e) m[e] = (lane * E + e < p.D) ? n[e & 3] : 0.0f;
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = lane * E + e;
                    const uint4 w = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word, (uint32_t)(i >> 2)));
                    float n0, n1;
                    if ((i & 2) == 0) box_muller_f32(w.x, w.y, n0, n1); else box_muller_f32(w.z, w.w, n0, n1);
                    m[e] = i < p.D ? ((i & 1) ? n1 : n0) : 0.0f;
                }
            }
        }
    }
    __device__ __forceinline__ ST draw_exp1() {
        if (kReplay) return (ST)p.exps[chain * p.cap_exps + cur_e++];
        const uint4 w = philox_scalar_words(p.key, gchain, step_word);
        return (ST)(-logf(u24_open(w.x)));
    }
    // next uniform of the step; f64 = 53-bit (tree merges), otherwise type T
    __device__ __forceinline__ double draw_uniform(bool f64) {
        ++n_unif;
        if (kReplay) return p.unifs[chain * p.cap_unifs + cur_u++];
        const uint32_t batch = q >> 6;
        if (batch != q_batch) {  // 32 lanes x 2 uniforms per Philox batch
            ubatch = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word,
                                                     kSubUnif + batch * 32 + (uint32_t)lane));
            q_batch = batch;
        }
        const int src = (q >> 1) & 31;
        const bool hi = q & 1;
        const uint32_t lo = __shfl_sync(kFull, hi ? ubatch.z : ubatch.x, src);
        const uint32_t hw = __shfl_sync(kFull, hi ? ubatch.w : ubatch.y, src);
        ++q;
        if (f64 || sizeof(ST) == 8) return u53_half_open(lo, hw);
        return (double)u24_half_open(hw);
    }

    // the 53 random bits behind the next f64 uniform (native mode only; same counter stream as draw_uniform)
    __device__ __forceinline__ uint64_t draw_u53() {
        ++n_unif;
        const uint32_t batch = q >> 6;
        if (batch != q_batch) {
            ubatch = philox4x32_10(p.key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step_word,
                                                     kSubUnif + batch * 32 + (uint32_t)lane));
            q_batch = batch;
        }
        const int src = (q >> 1) & 31;
        const bool hi = q & 1;
        const uint32_t lo = __shfl_sync(kFull, hi ? ubatch.z : ubatch.x, src);
        const uint32_t hw = __shfl_sync(kFull, hi ? ubatch.w : ubatch.y, src);
        ++q;
        return (((uint64_t)hw << 32) | lo) >> 11;
    }

    // ---- leapfrog, src/nuts.rs:979-996 (in place); returns logp'
    __device__ __forceinline__ float leapfrog(float (&x)[E], float (&m)[E], float (&g)[E], ST eps) {
        const float e = (float)eps;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            m[k] = A::mad(A::mul(g[k], e), 0.5f, m[k]);
            x[k] = A::mad(m[k], e, x[k]);
        }
        const float lp = tgt.logp_grad(x, g, lane);
        ++n_grad;
#pragma unroll
        for (int k = 0; k < E; ++k) m[k] = A::mad(A::mul(g[k], e), 0.5f, m[k]);
        return lp;
    }
    __device__ __forceinline__ float sumsq(const float (&m)[E]) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) s = A::mad(m[k], m[k], s);
        return warp_sum<A>(s);
    }
    // completes a target evaluation: full logp (lp_io) and sum m^2, one fused butterfly when logp is partial
    __device__ __forceinline__ float finish(float &lp_io, const float (&m)[E]) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) s = A::mad(m[k], m[k], s);
        if (Target::kPartial) warp_sum2<A>(lp_io, s);
        else s = warp_sum<A>(s);
        return s;
    }
    __device__ __forceinline__ float full_logp(float lp) { return Target::kPartial ? warp_sum<A>(lp) : lp; }
    // stop_criterion, src/nuts.rs:963-977 (true = keep going)
    __device__ __forceinline__ bool keep_going(const float (&xm)[E], const float (&xp)[E], const float (&pm)[E],
                                               const float (&pp)[E]) {
        float dm = 0.0f, dp = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const float diff = A::sub(xp[k], xm[k]);
            dm = A::mad(diff, pm[k], dm);
            dp = A::mad(diff, pp[k], dp);
        }
        warp_sum2<A>(dm, dp);
        return dm >= 0.0f && dp >= 0.0f;
    }
    __device__ __forceinline__ bool all_finite(const float (&v)[E]) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < E; ++k) ok = ok && isfinite(v[k]);
        return __all_sync(kFull, ok);
    }

    // find_reasonable_epsilon, src/nuts.rs:695-761
    __device__ ST find_reasonable_epsilon(const float (&x0)[E], const float (&m0)[E]) {
        float g0[E], x[E], m[E], g[E];
        const ST half = (ST)0.5;
        ST epsilon = (ST)1.0;
        const float ulogp = full_logp(tgt.logp_grad(x0, g0, lane));
        ++n_grad;
        auto leap = [&](ST e) {
#pragma unroll
            for (int k = 0; k < E; ++k) { x[k] = x0[k]; m[k] = m0[k]; g[k] = g0[k]; }
            return full_logp(leapfrog(x, m, g, e));
        };
        float ulogp_prime = leap(epsilon);
        ST k = (ST)1.0;
        while (!isfinite(ulogp_prime) && !all_finite(g)) {
            k = k * half;
            ulogp_prime = leap(epsilon * k);
        }
        epsilon = half * k * epsilon;
        const float pp0 = sumsq(m0);
        float lap_f = A::sub(A::sub(ulogp_prime, ulogp), A::mul(A::sub(sumsq(m), pp0), 0.5f));
        ST lap = (ST)(double)lap_f;
        const ST a = lap > s_log(half) ? (ST)1.0 : (ST)-1.0;
        while (a * lap > -a * s_log((ST)2.0)) {
            epsilon = epsilon * s_pow((ST)2.0, a);
            ulogp_prime = leap(epsilon);
            lap_f = A::sub(A::sub(ulogp_prime, ulogp), A::mul(A::sub(sumsq(m), pp0), 0.5f));
            lap = (ST)(double)lap_f;
        }
        return epsilon;
    }

    // One doubling = build_tree(edge, v, j), src/nuts.rs:764-946, iteratively.
    // cx/cm/cg: the edge to extend (in) and the new edge (out).  Outputs the subtree proposal, n', s', alpha, n_alpha.
    __device__ void doubling(float (&cx)[E], float (&cm)[E], float (&cg)[E], int v, int j, ST logu, ST eps, ST joint0,
                             float (&prop)[E], int &n_out, bool &s_out, ST &alpha_out, int &nalpha_out) {
        const ST veps = (ST)v * eps;
        const uint32_t n_leaves = 1u << j;
        float tfx[E], tfm[E];  // first leaf of the subtree currently being merged upward
        int tn = 0, tna = 0;
        ST ta = (ST)0.0;
        bool ts = true;
        for (uint32_t leaf = 0; leaf < n_leaves; ++leaf) {
            float lp = leapfrog(cx, cm, cg, veps);
            const float ss = finish(lp, cm);
            const float joint_f = A::sub(lp, A::mul(ss, 0.5f));
            const ST joint = (ST)(double)joint_f;
            tn = (logu < joint) ? 1 : 0;
            ts = (logu - (ST)1000.0) < joint;
            const ST ex = s_exp(joint - joint0);
            ta = ((ST)1.0 < ex || ex != ex) ? (ST)1.0 : ex;  // T::min(1, e): NaN -> 1
            tna = 1;
#pragma unroll
            for (int k = 0; k < E; ++k) { tfx[k] = cx[k]; tfm[k] = cm[k]; prop[k] = cx[k]; }
            int lvl = 0;
            bool pushed = false;
            while (true) {
                while (lvl < j && ((leaf >> lvl) & 1u)) {
                    // merge pending first half A = stack[lvl] with the later half T
                    const int an = s_n[lvl], ana = s_na[lvl];
                    const ST aa = (ST)s_a[lvl];
                    // u < n'' / max(n' + n'', 1) in f64 (src/nuts.rs:910-911)
                    bool take_b;
                    if (kReplay) {
                        const double u = draw_uniform(true);
                        if (tn == 0) take_b = false;
                        else if (an == 0) take_b = u < 1.0;
                        else take_b = u < ((double)tn / (double)(an + tn));
                    } else {
                        // native draws are k 2^-53 with a 53-bit integer k: k (n' + n'') < n'' 2^53 is the same test
                        // evaluated exactly (the f64 quotient is rounded, which can only matter when u equals the rounded
                        // quotient itself, a 2^-53 event) and keeps the FP64 division out of the merge path
                        const uint64_t k53 = draw_u53();
                        take_b = tn != 0 && (an == 0 || u53_below_ratio(k53, (uint32_t)tn, (uint32_t)(an + tn)));
                    }
                    load_level(lvl, 0, tfx);
                    load_level(lvl, 1, tfm);
                    if (!take_b) load_level(lvl, 2, prop);
                    tn += an;
                    // s' = s'_1 && s'_2 && stop_criterion(minus, plus); pending halves always have s' = true
                    if (ts) ts = (v == 1) ? keep_going(tfx, cx, tfm, cm) : keep_going(cx, tfx, cm, tfm);
                    ta = aa + ta;
                    tna += ana;
                    ++lvl;
                }
                if (lvl == j) break;
                if (ts) {  // park the finished first half at this level and build the next leaf
#ifndef MMC_NUTS_NO_LEVEL_SYNC
                    __syncwarp();  // every lane has consumed the previous occupant of this level (s_n / s_na / s_a reads above)
#endif
                    store_level(lvl, 0, tfx);
                    store_level(lvl, 1, tfm);
                    store_level(lvl, 2, prop);
                    if (lane == 0) { s_n[lvl] = tn; s_na[lvl] = tna; s_a[lvl] = (double)ta; }
                    __syncwarp();
                    pushed = true;
                    break;
                }
                ++lvl;  // failed first half: the parent returns it unchanged (src/nuts.rs:858)
            }
            if (!pushed) break;  // whole subtree of depth j is complete (or failed and fully unwound)
        }
        n_out = tn; s_out = ts; alpha_out = ta; nalpha_out = tna;
    }
};

template <class Target, class A, class ST, int E, bool kReplay>
__global__ void __launch_bounds__(kNutsWarps * 32, MMC_NUTS_MIN_BLOCKS) nuts_run_kernel(const Target tgt, const NutsParams p) {
    extern __shared__ __align__(16) float nuts_smem[];
    constexpr int V = 32 * E;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_stack = nuts_smem + warp * kNutsSmemLevels * 3 * V;
    const int64_t warp_slot = (int64_t)blockIdx.x * kNutsWarps + warp;
    const int n_glob = p.max_depth > kNutsSmemLevels ? p.max_depth - kNutsSmemLevels : 0;
    float *g_stack = p.scratch + warp_slot * (int64_t)n_glob * 3 * V;
    // per-warp scalar area after the vector stacks: 16 x (double alpha, int n, int n_alpha) = 256 B
    void *s_scal = reinterpret_cast<unsigned char *>(nuts_smem + kNutsWarps * kNutsSmemLevels * 3 * V) + warp * 256;
    NutsWarp<Target, A, ST, E, kReplay> w(tgt, p, lane, s_stack, g_stack, s_scal);
    unsigned long long my_depth_count = 0, n_trans = 0, tot_grad = 0, tot_unif = 0;

    while (true) {
        long long c = 0;
        if (lane == 0) c = (long long)atomicAdd(&p.counters[0], 1ULL);
        c = __shfl_sync(kFull, c, 0);
        if (c >= p.chains) break;
        w.chain = c;
        w.gchain = (uint64_t)(c + p.chain_offset);
        w.cur_n = w.cur_e = w.cur_u = 0;

        float pos[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int i = lane * E + k;
            pos[k] = i < p.D ? p.positions[c * p.D + i] : 0.0f;
        }
        if constexpr (kReplay) {
            if (p.tree_scal) {  // build_tree debug mode: one doubling, src/nuts.rs:764-946
                float cm[E], cg[E], prop[E], gp[E];
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    const int i = lane * E + k;
                    cm[k] = i < p.D ? p.tree_mom[c * p.D + i] : 0.0f;
                    cg[k] = i < p.D ? p.tree_grad[c * p.D + i] : 0.0f;
                }
                const double *sc = p.tree_scal + c * 4;
                int n_prime = 0, n_alpha = 0;
                bool s_prime = false;
                ST alpha = (ST)0.0;
                w.doubling(pos, cm, cg, sc[1] < 0.0 ? -1 : 1, p.tree_j, (ST)sc[0], (ST)sc[2], (ST)sc[3], prop, n_prime, s_prime,
                           alpha, n_alpha);
                const float lpp = w.full_logp(tgt.logp_grad(prop, gp, lane));
                float *ov = p.tree_out_vec + c * 5 * p.D;
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    const int i = lane * E + k;
                    if (i < p.D) {
                        ov[i] = pos[k]; ov[p.D + i] = cm[k]; ov[2 * p.D + i] = cg[k]; ov[3 * p.D + i] = prop[k];
                        ov[4 * p.D + i] = gp[k];
                    }
                }
                if (lane == 0) {
                    double *os = p.tree_out_scal + c * 6;
                    os[0] = (double)lpp; os[1] = (double)n_prime; os[2] = s_prime ? 1.0 : 0.0; os[3] = (double)alpha;
                    os[4] = (double)n_alpha; os[5] = (double)w.cur_u;
                }
                w.n_grad = 0; w.n_unif = 0;
                continue;
            }
        }
        double *st = p.state + c * 5;
        ST epsilon = (ST)st[0], epsilon_bar = (ST)st[1], h_bar = (ST)st[2], mu;
        long long m = (long long)st[4];
        const ST gamma = (ST)0.05, kappa = (ST)0.75, delta = (ST)p.target_accept;
        const long long t_0 = 10;

        auto store_draw = [&](int64_t slot) {
            float *o = p.out + (c * p.out_pitch + slot) * p.D;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int i = lane * E + k;
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
            if (d <= tiny) epsilon = w.find_reasonable_epsilon(pos, m0);
            mu = s_log((ST)10.0 * epsilon);
        }

        const int64_t total = p.n_collect + p.n_discard;
        const int64_t first = (p.progress || p.resume) ? 0 : 1;
        for (int64_t it = first; it < total; ++it) {
            // ---- NUTSChain::step, src/nuts.rs:550-691
            m += 1;
            w.step_word = (uint32_t)m;
            w.q = 0; w.q_batch = 0xffffffffu;
            const ST eps_used = epsilon;
            const int64_t unifs_before = w.cur_u;
            float mom0[E], grad[E];
            w.draw_normals(mom0);
            float ulogp = tgt.logp_grad(pos, grad, lane);
            ++w.n_grad;
            const float ss0 = w.finish(ulogp, mom0);
            const float joint_f = A::sub(ulogp, A::mul(ss0, 0.5f));
            const ST joint = (ST)(double)joint_f;
            const ST logu = joint - w.draw_exp1();
            float xm[E], pm[E], gm[E], xp[E], pp[E], gp[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                xm[k] = xp[k] = pos[k];
                pm[k] = pp[k] = mom0[k];
                gm[k] = gp[k] = grad[k];
            }
            int j = 0;
            int n = 1;
            bool s = true;
            ST alpha = (ST)0.0;
            int n_alpha = 0;
            while (s) {
                const ST u1 = (ST)w.draw_uniform(false);
                const int v = (u1 < (ST)0.5) ? 1 : -1;
                float prop[E];
                int n_prime;
                bool s_prime;
                if (v == -1) w.doubling(xm, pm, gm, v, j, logu, epsilon, joint, prop, n_prime, s_prime, alpha, n_alpha);
                else w.doubling(xp, pp, gp, v, j, logu, epsilon, joint, prop, n_prime, s_prime, alpha, n_alpha);
                const ST ratio = (ST)n_prime / (ST)n;
                const ST tmp = ((ST)1.0 < ratio) ? (ST)1.0 : ratio;
                const ST u2 = (ST)w.draw_uniform(false);
                if (s_prime && (u2 < tmp)) {
#pragma unroll
                    for (int k = 0; k < E; ++k) pos[k] = prop[k];
                }
                n += n_prime;
                s = s_prime && w.keep_going(xm, xp, pm, pp);
                j += 1;
                if (j >= p.max_depth) s = false;
            }
            if (lane == (j < 31 ? j : 31)) ++my_depth_count;
            ++n_trans;
            if (p.trace && lane == 0) {
                double *tr = p.trace + (c * p.trace_pitch + it) * 8;
                tr[0] = (double)joint; tr[1] = (double)logu; tr[2] = (double)n; tr[3] = (double)alpha; tr[4] = (double)n_alpha;
                tr[5] = (double)j; tr[6] = (double)eps_used; tr[7] = kReplay ? (double)(w.cur_u - unifs_before) : (double)w.q;
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
            const int64_t idx = p.progress ? it : it;  // run: m-th step fills slot m - n_discard; run_progress: i - n_discard
            if (idx >= p.n_discard) store_draw(idx - p.n_discard);
        }
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int i = lane * E + k;
            if (i < p.D) p.positions[c * p.D + i] = pos[k];
        }
        if (lane == 0) {
            st[0] = (double)epsilon; st[1] = (double)epsilon_bar; st[2] = (double)h_bar; st[3] = (double)mu;
            st[4] = (double)m;
        }
        tot_grad += w.n_grad; tot_unif += w.n_unif;
        w.n_grad = 0; w.n_unif = 0;
    }
    if (lane == 0) {
        atomicAdd(&p.counters[1], tot_grad);
        atomicAdd(&p.counters[2], n_trans);
        atomicAdd(&p.counters[3], tot_unif);
    }
    if (my_depth_count) atomicAdd(&p.counters[8 + lane], my_depth_count);
}

}  // namespace mmc
