// K3 host driver + FP32 SIMT GEMM path (see mmc_dense.cuh).  The tensor-core path lives in mmc_dense_tc.cu.
//
// One transition = 1 + (L + 1) + 1 launches on the caller's stream:
//   dense_begin   : momenta (Philox / replay), Delta = X - mu, kinetic energy of the fresh momenta
//   dense_gemm x (L+1): Z = Delta . P (P = Sigma^-1, symmetric) with the leapfrog fused into the epilogue:
//        FIRST: quad_cur += rowdot(Z, Delta);  gh = -Z eps/2;  p = p0 + gh;             Delta' = Delta + eps p
//        MID  :                                gh = -Z eps/2;  p = (p + gh) + gh;       Delta' = Delta + eps p
//        LAST : quad_prop += rowdot(Z, Delta); gh = -Z eps/2;  p = p + gh;  ke_prop += rowsum(p^2)
//     (the two half-kicks around a gradient use the same cached g eps/2, src/hmc.rs:408,420-425; Delta is double
//      buffered because other tiles still read the old Delta as their A operand)
//   dense_accept  : H = -logp + ke/2, accept iff H_cur - H_prop >= ln(u) (src/hmc.rs:343-376), X = Delta' + mu, draw out
#include <vector>

#include "mmc_dense.cuh"

namespace mmc {


// ---------------------------------------------------------------- begin / accept
__global__ void dense_begin_kernel(const float *__restrict__ pos, const float *__restrict__ mean, float *__restrict__ delta,
                                   float *__restrict__ mom, float *__restrict__ scal, const float *__restrict__ rp_mom,
                                   const float *__restrict__ rp_u, int64_t chains, int D, int Dp, int64_t chain_offset,
                                   uint32_t gstep, int64_t local_step, uint2 key) {
    // one warp per chain; internal rows have pitch Dp >= D, columns >= D are zero
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= chains) return;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    float ke = 0.f;
    for (int j = lane; j < Dp / 4; j += 32) {
        const int i = 4 * j;
        if (i >= D) {
            *reinterpret_cast<float4 *>(delta + c * Dp + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(mom + c * Dp + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const float4 x = *reinterpret_cast<const float4 *>(pos + c * D + i);
        const float4 m = *reinterpret_cast<const float4 *>(mean + i);
        float4 p;
        if (rp_mom) {
            p = *reinterpret_cast<const float4 *>(rp_mom + (local_step * chains + c) * D + i);
        } else {
            const uint4 w = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)j));
            box_muller_f32(w.x, w.y, p.x, p.y);
            box_muller_f32(w.z, w.w, p.z, p.w);
        }
        *reinterpret_cast<float4 *>(delta + c * Dp + i) = make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
        *reinterpret_cast<float4 *>(mom + c * Dp + i) = p;
        ke += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ke += __shfl_xor_sync(0xffffffffu, ke, o);
    if (lane == 0) {
        scal[0 * chains + c] = ke;
        scal[1 * chains + c] = 0.f;
        scal[2 * chains + c] = 0.f;
        scal[3 * chains + c] = 0.f;
        float u;
        if (rp_u) u = rp_u[local_step * chains + c];
        else u = u24_half_open(philox_scalar_words(key, gchain, gstep).x);
        scal[4 * chains + c] = u;
    }
}

__global__ void dense_accept_kernel(float *__restrict__ pos, const float *__restrict__ mean, const float *__restrict__ delta,
                                    const float *__restrict__ delta_lo, const float *__restrict__ scal, float norm_const, float *__restrict__ out,
                                    float *__restrict__ trace, unsigned long long *accept_count, int64_t chains, int D, int Dp,
                                    int64_t out_pitch, int64_t slot, int64_t local_step) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= chains) return;
    const float ke_cur = scal[0 * chains + c], quad_cur = scal[1 * chains + c];
    const float ke_prop = scal[2 * chains + c], quad_prop = scal[3 * chains + c], u = scal[4 * chains + c];
    const float logp_cur = norm_const - 0.5f * quad_cur, logp_prop = norm_const - 0.5f * quad_prop;
    const float h_cur = -logp_cur + 0.5f * ke_cur, h_prop = -logp_prop + 0.5f * ke_prop;
    const float accept_logp = h_cur - h_prop;
    const bool acc = accept_logp >= logf(u);
    for (int j = lane; j < D / 4; j += 32) {
        const int i = 4 * j;
        float4 x = *reinterpret_cast<const float4 *>(pos + c * D + i);
        if (acc) {
            float4 d = *reinterpret_cast<const float4 *>(delta + c * Dp + i);
            if (delta_lo) {  // tensor-core path keeps Delta as an exact hi + lo split
                const float4 l = *reinterpret_cast<const float4 *>(delta_lo + c * Dp + i);
                d = make_float4(d.x + l.x, d.y + l.y, d.z + l.z, d.w + l.w);
            }
            const float4 m = *reinterpret_cast<const float4 *>(mean + i);
            x = make_float4(d.x + m.x, d.y + m.y, d.z + m.z, d.w + m.w);
            *reinterpret_cast<float4 *>(pos + c * D + i) = x;
        }
        if (out) *reinterpret_cast<float4 *>(out + (c * out_pitch + slot) * D + i) = x;
    }
    if (lane == 0) {
        if (acc) atomicAdd(accept_count, 1ULL);
        if (trace) reinterpret_cast<float4 *>(trace)[local_step * chains + c] =
            make_float4(logp_cur, logp_prop, accept_logp, acc ? 1.f : 0.f);
    }
}

// ---------------------------------------------------------------- FP32 SIMT GEMM with the leapfrog epilogue
// C tile 128 x 128, K step 16, 256 threads, 8 x 8 outputs per thread, double-buffered shared memory.
// A = Delta [M, D] row-major (K contiguous), B = P [D, D] row-major: Z[m][n] = sum_k A[m][k] B[k][n].
constexpr int kBM = 128, kBN = 128, kBK = 16;

__device__ __forceinline__ void dense_epilogue(int mode, float z, float dlt, float &p, float &dnext, float eps,
                                               float eps_half, float &quad, float &ke) {
    const float gh = -z * eps_half;
    if (mode != kModeMid) quad = fmaf(z, dlt, quad);
    if (mode == kModeFirst) {
        p = p + gh;
        dnext = fmaf(eps, p, dlt);
    } else if (mode == kModeMid) {
        p = (p + gh) + gh;
        dnext = fmaf(eps, p, dlt);
    } else {
        p = p + gh;
        ke = fmaf(p, p, ke);
    }
}

__global__ void __launch_bounds__(256) dense_gemm_simt_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                              float *__restrict__ mom, float *__restrict__ dnext,
                                                              float *__restrict__ scal, int64_t M, int D, float eps, int mode) {
    __shared__ __align__(16) float As[2][kBK][kBM + 4];
    __shared__ __align__(16) float Bs[2][kBK][kBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 rows x 8 cols (interleaved by 16)
    const int64_t m0 = (int64_t)blockIdx.y * kBM;
    const int n0 = blockIdx.x * kBN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loaders: A tile 128 x 16 (each thread 2 float4 along K), B tile 16 x 128 (each thread 2 float4 along N)
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;       // 128 rows x 2 halves of 8 k
    const int b_k = tid >> 4, b_n = (tid & 15) * 8;        // 16 k x 16 groups of 8 n
    auto load_tile = [&](int buf, int k0) {
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
        if (m0 + a_row < M) {
            const float *src = A + (m0 + a_row) * D + k0 + a_k;
            a0 = *reinterpret_cast<const float4 *>(src);
            a1 = *reinterpret_cast<const float4 *>(src + 4);
        }
        As[buf][a_k + 0][a_row] = a0.x; As[buf][a_k + 1][a_row] = a0.y; As[buf][a_k + 2][a_row] = a0.z; As[buf][a_k + 3][a_row] = a0.w;
        As[buf][a_k + 4][a_row] = a1.x; As[buf][a_k + 5][a_row] = a1.y; As[buf][a_k + 6][a_row] = a1.z; As[buf][a_k + 7][a_row] = a1.w;
        const float *bs = B + (int64_t)(k0 + b_k) * D + n0 + b_n;
        *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = *reinterpret_cast<const float4 *>(bs);
        *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n + 4]) = *reinterpret_cast<const float4 *>(bs + 4);
    };
    load_tile(0, 0);
    __syncthreads();
    const int nk = D / kBK;
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tile(buf ^ 1, (kb + 1) * kBK);
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            float a[8], b[8];
            const float4 av0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 av1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float4 bv0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float4 bv1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
            a[0] = av0.x; a[1] = av0.y; a[2] = av0.z; a[3] = av0.w; a[4] = av1.x; a[5] = av1.y; a[6] = av1.z; a[7] = av1.w;
            b[0] = bv0.x; b[1] = bv0.y; b[2] = bv0.z; b[3] = bv0.w; b[4] = bv1.x; b[5] = bv1.y; b[6] = bv1.z; b[7] = bv1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue: rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; cols tx*4 + {0..3} and 64 + tx*4 + {0..3}
    const float eps_half = eps * 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        float quad = 0.f, ke = 0.f;
        if (m < M) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + h * 64 + tx * 4;
                const int64_t off = m * D + n;
                const float4 dl = *reinterpret_cast<const float4 *>(A + off);
                float4 p = *reinterpret_cast<const float4 *>(mom + off);
                float4 dn = dl;
                dense_epilogue(mode, acc[i][h * 4 + 0], dl.x, p.x, dn.x, eps, eps_half, quad, ke);
                dense_epilogue(mode, acc[i][h * 4 + 1], dl.y, p.y, dn.y, eps, eps_half, quad, ke);
                dense_epilogue(mode, acc[i][h * 4 + 2], dl.z, p.z, dn.z, eps, eps_half, quad, ke);
                dense_epilogue(mode, acc[i][h * 4 + 3], dl.w, p.w, dn.w, eps, eps_half, quad, ke);
                *reinterpret_cast<float4 *>(mom + off) = p;
                if (mode != kModeLast) *reinterpret_cast<float4 *>(dnext + off) = dn;
            }
        }
        if (mode != kModeMid) {
            // the 16 threads sharing this row are 16 consecutive lanes (tx = lane & 15)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                quad += __shfl_xor_sync(0xffffffffu, quad, o);
                ke += __shfl_xor_sync(0xffffffffu, ke, o);
            }
            if (tx == 0 && m < M) {
                atomicAdd(scal + (mode == kModeFirst ? 1 : 3) * M + m, quad);
                if (mode == kModeLast) atomicAdd(scal + 2 * M + m, ke);
            }
        }
    }
}

int dense_gemm_tc(DenseState *st, int cur, int64_t M, int D, float eps, int mode, cudaStream_t stream);  // mmc_dense_tc.cu
int dense_gemm_tc_chain(DenseState *st, int64_t M, int D, float eps, int L, cudaStream_t stream);
int dense_tc_prepare(DenseState *st);
int dense_tc_split_delta(DenseState *st, cudaStream_t stream);
void dense_tc_destroy(DenseState *st);

// ---------------------------------------------------------------- host driver
int dense_create(DenseState **out, const mmc_target_desc *t, int64_t chains) {
    MMC_REQUIRE(t->vec && t->mat, "dense Gaussian target needs mean (vec) and precision (mat)");
    MMC_REQUIRE(t->dim % 4 == 0 && t->dim >= 4, "dense Gaussian: dim must be a multiple of 4, got %d", t->dim);
    DenseState *st = new DenseState();
    st->D = t->dim;
    // internal pitch: whole 256-column tiles (tcgen05 path; the FP32 SIMT path needs 128), zero padded; D = 128 stays as it is
    st->Dp = (t->dim % 256 == 0 || t->dim == 128) ? t->dim : (t->dim + 255) / 256 * 256;
    st->chains = chains;
    st->norm_const = (float)t->params[0];
    const size_t D = (size_t)t->dim, Dp = (size_t)st->Dp, md = (size_t)chains * Dp * sizeof(float);
    auto fail = [&](cudaError_t e) { dense_destroy(st); return cuda_fail(e, "dense_create", __FILE__, __LINE__); };
    cudaError_t e;
    if ((e = cudaMalloc((void **)&st->d_mean, Dp * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void **)&st->d_prec, Dp * Dp * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void **)&st->d_delta[0], md)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void **)&st->d_delta[1], md)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void **)&st->d_mom, md)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void **)&st->d_scal, 6 * (size_t)chains * 4)) != cudaSuccess) return fail(e);
    // both Delta buffers start at zero: the epilogue only ever writes 0 into the padding columns, but it reads them first
    if ((e = cudaMemset(st->d_delta[0], 0, md)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(st->d_delta[1], 0, md)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(st->d_mom, 0, md)) != cudaSuccess) return fail(e);
    std::vector<float> mean(Dp, 0.0f);
    for (size_t i = 0; i < D; ++i) mean[i] = t->vec[i];
    if ((e = cudaMemcpy(st->d_mean, mean.data(), Dp * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    // symmetrise on the host (grad = -Z assumes P = P^T; also lets the tensor path use P as its own transpose)
    std::vector<float> P(Dp * Dp, 0.0f);
    for (size_t i = 0; i < D; ++i)
        for (size_t j = 0; j < D; ++j) P[i * Dp + j] = 0.5f * (t->mat[i * D + j] + t->mat[j * D + i]);
    if ((e = cudaMemcpy(st->d_prec, P.data(), Dp * Dp * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    *out = st;
    return MMC_OK;
}

void dense_destroy(DenseState *st) {
    if (!st) return;
    dense_tc_destroy(st);
    cudaFree(st->d_mean);
    cudaFree(st->d_prec);
    cudaFree(st->d_delta[0]);
    cudaFree(st->d_delta[1]);
    cudaFree(st->d_mom);
    cudaFree(st->d_scal);
    cudaFree(st->d_prec_split);
    cudaFree(st->d_delta_split[0]);
    cudaFree(st->d_delta_split[1]);
    cudaFree(st->d_prec_x);
    cudaFree(st->d_delta_x[0]);
    cudaFree(st->d_delta_x[1]);
    cudaFree(st->d_ready);
    delete st;
}

int dense_run(DenseState *st, const DenseRunArgs &a, cudaStream_t stream) {
    MMC_REQUIRE(a.n_leapfrog >= 1, "dense Gaussian HMC needs n_leapfrog >= 1");
    MMC_REQUIRE(a.chains == st->chains, "chain count changed");
    const int D = st->D, Dp = st->Dp;
    const int64_t M = a.chains;
    const int64_t steps = a.n_collect + a.n_discard;
    const uint2 key = seed_key(a.seed);
    const unsigned wgrid = (unsigned)((M * 32 + 255) / 256);
    const dim3 ggrid((unsigned)(Dp / kBN), (unsigned)((M + kBM - 1) / kBM));
    MMC_REQUIRE(Dp % 256 == 0 || a.gemm_path == 0, "tcgen05 GEMM paths need whole 256-column tiles (dim 128 runs on the FP32 path), got %d", D);
    if (a.gemm_path >= 1) {
        int rc = dense_tc_prepare(st);
        if (rc) return rc;
        st->tc_pair = a.gemm_path >= 2;
        st->tc_mixed = a.gemm_path == 3;
    }
    for (int64_t s = 0; s < steps; ++s) {
        dense_begin_kernel<<<wgrid, 256, 0, stream>>>(a.positions, st->d_mean, st->d_delta[0], st->d_mom, st->d_scal,
                                                      a.momenta, a.u, M, D, Dp, a.chain_offset, (uint32_t)(a.step_base + s), s, key);
        if (a.gemm_path >= 1) {
            int rc = dense_tc_split_delta(st, stream);
            if (rc) return rc;
        }
        int cur = 0;
        bool chained = false;
        if (a.gemm_path == 3) {   // all L + 1 GEMMs of the transition in one launch where the chain kernel applies
            const int rc = dense_gemm_tc_chain(st, M, Dp, a.eps, a.n_leapfrog, stream);
            if (rc == MMC_OK) { chained = true; cur = a.n_leapfrog & 1; }
            else if (rc != MMC_ERR_UNSUPPORTED) return rc;
        }
        for (int l = 0; l <= a.n_leapfrog && !chained; ++l) {
            const int mode = l == 0 ? kModeFirst : (l == a.n_leapfrog ? kModeLast : kModeMid);
            if (a.gemm_path >= 1) {
                int rc = dense_gemm_tc(st, cur, M, Dp, a.eps, mode, stream);
                if (rc) return rc;
            } else {
                dense_gemm_simt_kernel<<<ggrid, 256, 0, stream>>>(st->d_delta[cur], st->d_prec, st->d_mom,
                                                                  st->d_delta[cur ^ 1], st->d_scal, M, Dp, a.eps, mode);
            }
            if (mode != kModeLast) cur ^= 1;
        }
        const bool collect = s >= a.n_discard && a.out;
        const bool split = a.gemm_path == 1 || a.gemm_path == 2;   // 3xTF32 paths keep Delta as hi + lo; the others in full
        const float *fin = split ? st->d_delta_split[cur] : st->d_delta[cur];
        const float *fin_lo = split ? st->d_delta_split[cur] + (size_t)M * Dp : nullptr;
        dense_accept_kernel<<<wgrid, 256, 0, stream>>>(a.positions, st->d_mean, fin, fin_lo, st->d_scal, st->norm_const,
                                                       collect ? a.out : nullptr, a.trace, a.accept_count, M, D, Dp, a.out_pitch,
                                                       collect ? s - a.n_discard : 0, s);
    }
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

__global__ void dense_export_tape_kernel(uint2 key, int64_t chains, int D, int64_t chain_offset, int64_t step_base,
                                         int64_t steps, float *momenta, float *u) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (step, chain, group of 4 dims)
    const int64_t groups = D / 4;
    if (idx >= steps * chains * groups) return;
    const int64_t j = idx % groups, sc = idx / groups, c = sc % chains, s = sc / chains;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    const uint32_t gstep = (uint32_t)(step_base + s);
    const uint4 w = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), gstep, (uint32_t)j));
    float4 p;
    box_muller_f32(w.x, w.y, p.x, p.y);
    box_muller_f32(w.z, w.w, p.z, p.w);
    *reinterpret_cast<float4 *>(momenta + sc * D + 4 * j) = p;
    if (j == 0) u[sc] = u24_half_open(philox_scalar_words(key, gchain, gstep).x);
}

int dense_export_tape(int64_t chains, int D, int64_t chain_offset, uint64_t seed, int64_t step_base, int64_t steps,
                      float *momenta, float *u, cudaStream_t stream) {
    const int64_t total = steps * chains * (D / 4);
    dense_export_tape_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(seed_key(seed), chains, D, chain_offset,
                                                                                  step_base, steps, momenta, u);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace mmc
