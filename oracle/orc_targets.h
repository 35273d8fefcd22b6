/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_rng.h header).
 *
 * CPU restatement of the reference's target / proposal arithmetic.  Every function cites the
 * reference lines it follows (paths relative to /root/reference).  Tensor maths of the burn
 * NdArray backend is f32 regardless of the scalar type T (src/nuts.rs:1022), so the gradient
 * targets are float; the MH targets are double (configs C1/C2 use f64).
 *
 * Gradients: the reference obtains them by burn autodiff (src/distributions.rs:81-87), a third-party
 * crate (burn 0.18.0, Cargo.lock:675) that is not in the tree.  They are restated analytically; the
 * restatement is pinned by the NUTS golden vectors (src/nuts.rs:1099-1120, 1150-1191).
 *
 * Build with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 */
#ifndef ORC_TARGETS_H
#define ORC_TARGETS_H

#include <math.h>
#include <stdint.h>

enum {
    ORC_T_GAUSSIAN2D = 1,      /* src/distributions.rs:159-206  (MH target, f64)                */
    ORC_T_ISO_GAUSSIAN = 2,    /* src/distributions.rs:394-402  (MH target, f64)                */
    ORC_T_POISSON = 3,         /* examples/poisson_mh.rs:10-26  (integer state)                 */
    ORC_T_ROSENBROCK_ND = 4,   /* src/distributions.rs:527-547                                   */
    ORC_T_ROSENBROCK_2D = 5,   /* src/distributions.rs:491-524                                   */
    ORC_T_DIFF_GAUSSIAN2D = 6, /* src/distributions.rs:213-316                                   */
    ORC_T_DENSE_GAUSSIAN = 7,  /* D-dim generalisation of :262-288 (BASELINE config C4)          */
    ORC_T_STD_NORMAL = 8       /* test-only target of src/nuts.rs:1024-1037                      */
};

typedef struct {
    int kind;
    int dim;
    double p[8];        /* kind-specific scalars, see orc_target_prepare */
    const float *vec;   /* DENSE: mean[D] */
    const float *mat;   /* DENSE: precision (inverse covariance) [D,D] row-major */
    /* derived (filled by orc_target_prepare) */
    float inv_cov[4];
    float mean2[2];
    float norm_const;
    float a, b;
} orc_target;

/* DiffableGaussian2D::new, src/distributions.rs:227-251 (T = f64 in the goldens), then the
 * from_floats casts of :296-310.  p = {mean0, mean1, c00, c01, c10, c11}. */
static inline void orc_target_prepare(orc_target *t) {
    if (t->kind == ORC_T_DIFF_GAUSSIAN2D) {
        double c00 = t->p[2], c01 = t->p[3], c10 = t->p[4], c11 = t->p[5];
        double det = c00 * c11 - c01 * c10;
        double inv_det = 1.0 / det;
        t->inv_cov[0] = (float)(c11 * inv_det);
        t->inv_cov[1] = (float)(-c01 * inv_det);
        t->inv_cov[2] = (float)(-c10 * inv_det);
        t->inv_cov[3] = (float)(c00 * inv_det);
        double logdet = log(det);
        double two = 2.0;
        t->norm_const = (float)(-(two * log(two * M_PI) + logdet) / two);
        t->mean2[0] = (float)t->p[0];
        t->mean2[1] = (float)t->p[1];
    } else if (t->kind == ORC_T_ROSENBROCK_2D) {
        t->a = (float)t->p[0];
        t->b = (float)t->p[1];
    } else if (t->kind == ORC_T_DENSE_GAUSSIAN) {
        t->norm_const = (float)t->p[0];
    }
}

/* (logp, grad) of a gradient target at x[D], all f32.  Returns logp. */
static inline float orc_logp_grad_f32(const orc_target *t, const float *x, float *g) {
    const int D = t->dim;
    switch (t->kind) {
    case ORC_T_ROSENBROCK_ND: {
        /* src/distributions.rs:536-546: -(sum_i 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2) */
        float acc = 0.0f;
        for (int i = 0; i < D; ++i) g[i] = 0.0f;
        for (int i = 0; i + 1 < D; ++i) {
            float tt = x[i + 1] - x[i] * x[i];
            float u = 1.0f - x[i];
            acc = acc + (tt * tt * 100.0f + u * u);
            g[i] = g[i] + (400.0f * x[i] * tt + 2.0f * u);
            g[i + 1] = g[i + 1] + (-200.0f * tt);
        }
        return -acc;
    }
    case ORC_T_ROSENBROCK_2D: {
        /* src/distributions.rs:515-523: -((a - x)^2 + b (y - x^2)^2) */
        float xx = x[0], y = x[1];
        float u = t->a - xx;
        float tt = y - xx * xx;
        g[0] = 2.0f * u + 4.0f * t->b * xx * tt;
        g[1] = -2.0f * t->b * tt;
        return -(u * u + tt * tt * t->b);
    }
    case ORC_T_DIFF_GAUSSIAN2D: {
        /* src/distributions.rs:296-315: z = delta^T P ; quad = z . delta ; -0.5 quad + norm_const.
         * autodiff gradient of that graph: -0.5 (z + P delta). */
        float d0 = x[0] - t->mean2[0], d1 = x[1] - t->mean2[1];
        const float *P = t->inv_cov;
        float z0 = d0 * P[0] + d1 * P[2];
        float z1 = d0 * P[1] + d1 * P[3];
        float w0 = P[0] * d0 + P[1] * d1;
        float w1 = P[2] * d0 + P[3] * d1;
        float quad = z0 * d0 + z1 * d1;
        g[0] = -0.5f * (z0 + w0);
        g[1] = -0.5f * (z1 + w1);
        return -(quad * 0.5f) + t->norm_const;
    }
    case ORC_T_DENSE_GAUSSIAN: {
        /* generalisation of src/distributions.rs:262-288 to D dims with a symmetric precision P:
         * z = delta P, logp = norm_const - 0.5 z.delta, grad = -z.  The contraction is accumulated in
         * f64 and rounded once (matrixmultiply's blocked f32 order is not restated). */
        double quad = 0.0;
        for (int j = 0; j < D; ++j) {
            double z = 0.0;
            for (int i = 0; i < D; ++i) z += (double)(x[i] - t->vec[i]) * (double)t->mat[(size_t)i * D + j];
            float zf = (float)z;
            g[j] = -zf;
            quad += (double)zf * (double)(x[j] - t->vec[j]);
        }
        return t->norm_const - 0.5f * (float)quad;
    }
    case ORC_T_STD_NORMAL: {
        /* src/nuts.rs:1032-1036: -(sum 0.5 x^2) */
        float acc = 0.0f;
        for (int i = 0; i < D; ++i) {
            acc = acc + x[i] * x[i] * 0.5f;
            g[i] = -x[i];
        }
        return -acc;
    }
    default:
        for (int i = 0; i < D; ++i) g[i] = NAN;
        return NAN;
    }
}

/* ---- MH targets / proposals (f64) ---- */

/* Gaussian2D::unnorm_logp, src/distributions.rs:193-205.  p = {mean0, mean1, a, b, c, d}.
 * diff.dot(&inv_cov) is row-vector x matrix, then .dot(&diff). */
static inline double orc_gaussian2d_unnorm_logp(const double *p, const double *x) {
    double a = p[2], b = p[3], c = p[4], d = p[5];
    double det = a * d - b * c;
    double i00 = d / det, i01 = -b / det, i10 = -c / det, i11 = a / det;
    double d0 = x[0] - p[0], d1 = x[1] - p[1];
    double r0 = d0 * i00 + d1 * i10;
    double r1 = d0 * i01 + d1 * i11;
    return -0.5 * (r0 * d0 + r1 * d1);
}

/* Gaussian2D::logp (Normalized), src/distributions.rs:164-186. */
static inline double orc_gaussian2d_logp(const double *p, const double *x) {
    double a = p[2], b = p[3], c = p[4], d = p[5];
    double term1 = -log(2.0 * M_PI);
    double det = a * d - b * c;
    double term2 = -0.5 * log(fabs(det));
    return term1 + term2 + orc_gaussian2d_unnorm_logp(p, x);
}

/* IsotropicGaussian as Target, src/distributions.rs:394-402. */
static inline double orc_iso_unnorm_logp(double std, const double *x, int D) {
    double sum = 0.0;
    for (int i = 0; i < D; ++i) sum = sum + x[i] * x[i];
    return -0.5 * sum / (std * std);
}

/* IsotropicGaussian::logp (proposal density), src/distributions.rs:374-386.  Note the
 * normaliser ln(var * pi * std * std), kept verbatim. */
static inline double orc_iso_proposal_logp(double std, const double *from, const double *to, int D) {
    double lp = 0.0;
    double d = (double)D;
    double two = 2.0;
    double var = std * std;
    for (int i = 0; i < D; ++i) {
        double diff = to[i] - from[i];
        double exponent = -(diff * diff) / (two * var);
        lp += exponent;
    }
    lp += -d * 0.5 * log(var * M_PI * std * std);
    return lp;
}

/* ln_factorial, examples/poisson_mh.rs:79-89: sum_{i=1..k} ln(i) in that order. */
static inline double orc_ln_factorial(uint64_t k) {
    if (k < 2) return 0.0;
    double acc = 0.0;
    for (uint64_t i = 1; i <= k; ++i) acc += log((double)i);
    return acc;
}
/* PoissonTarget::unnorm_logp, examples/poisson_mh.rs:19-25. */
static inline double orc_poisson_logp(double lambda, uint64_t k) {
    double kf = (double)k;
    return -lambda + kf * log(lambda) - orc_ln_factorial(k);
}
/* NonnegativeProposal::logp, examples/poisson_mh.rs:53-71. */
static inline double orc_nonneg_logq(uint64_t x, uint64_t y) {
    if (x == 0) return y == 1 ? 0.0 : -INFINITY;
    return (y == x + 1 || y + 1 == x) ? log(0.5) : -INFINITY;
}

#endif
