// K3 tensor-core path: Z = Delta . P on tcgen05 (5th-gen tensor cores), operands staged by TMA, fp32
// accumulators in TMEM, leapfrog fused into the epilogue (see mmc_dense.cu for the step structure).
//
// Precision: the reference computes this contraction in fp32 (burn NdArray matmul).  Plain TF32 (10-bit mantissa)
// misses the 1e-5 parity bar, so every operand is split x = hi + lo with hi = x truncated to TF32 and
// lo = x - hi (exact), and three MMAs accumulate into the same TMEM tile:  hi.hi + hi.lo + lo.hi  (the dropped
// lo.lo term is < 2^-20 relative).  The splits are produced by the previous epilogue, so they cost no extra pass.
//
// Tile: BLOCK_M 128 chains x BLOCK_N 256 columns x BLOCK_K 16/32 (one swizzle row of fp32), UMMA 128 x 256 x 8
// (kind::tf32, cta_group::1).  Persistent CTAs (one per SM) walk the tile list; warp roles (320 threads):
// warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2-9 = epilogue (two
// warps per 32-lane TMEM quarter, one per 128-column half).  Pipelines: full/empty mbarriers per shared-memory
// stage (TMA -> MMA -> tcgen05.commit) and tmem_full/tmem_empty per accumulator (MMA <-> epilogue); the
// accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of one tile overlaps the main loop
// of the next.
//
// Mixed split (gemm_path 3, the default; CTA-pair kernel only): the two cross terms hi.lo + lo.hi only need ~9 bits of
// each factor (lo is already < 2^-10 of its operand), so they run as ONE K-concatenated kind::f16 MMA on bf16 copies,
// [bf16(a_hi) | bf16(a_lo)] . [bf16(b_lo) | bf16(b_hi)]^T, at twice the TF32 rate: 1 TF32 + 2 BF16 MMAs of cost 1 + 2 x 1/2
// = 2 TF32-equivalents per product instead of 3.  Operand bytes are unchanged (hi fp32 + one bf16 pair per element), the
// full-precision Delta lives in the FP32 path's ping-pong buffers and is what the epilogue reads and advances.
#include <cuda.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "mmc_dense.cuh"

namespace mmc {


namespace tc {

#ifndef MMC_TC_BK
#define MMC_TC_BK 32
#endif
// BLOCK_K 32 = one 128-byte swizzle row per operand row (2 stages of 96 KB; 3 of 64 KB in the CTA-pair kernel),
// BLOCK_K 16 = 64-byte swizzle rows (4 / 6 stages).  With the 256-bit epilogue both run within 3 % of each other;
// 32 is slightly ahead (measured).
constexpr int BM = 128, BN = 256, BK = MMC_TC_BK, kStages = (BK == 32 ? 2 : 4), UMMA_K = 8;
constexpr uint32_t kRowBytes = BK * 4;                     // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
static_assert(BK == 32 || BK == 16, "BLOCK_K must be 32 or 16");
constexpr uint32_t kABytes = BM * BK * 4;   // 16 KB (one of hi / lo)
constexpr uint32_t kBBytes = BN * BK * 4;   // 32 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;  // 96 KB
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quarter, one per 128-column half of the tile
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kTmemCols = 512;          // two 256-column fp32 accumulators

struct Maps {
    CUtensorMap a_hi[2], a_lo[2], b_hi, b_lo;
    CUtensorMap b_hi_half, b_lo_half;   // 128-row boxes for the CTA-pair kernel
    CUtensorMap a_x[2], b_x_half;       // mixed split: bf16 [rows, 2 D] cross-term operands, boxes of 64 x 128
    CUtensorMap b_hi_q, b_x_q;          // 64-row boxes of B for the 4-CTA cluster kernel (each CTA loads half of its half)
    CUtensorMap a_full[2];              // MMC_TC_HW_TRUNC experiment: the full-precision Delta as the TF32 operand
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major tile, rows of kRowBytes with the matching swizzle, 8-row groups
// 8 * kRowBytes apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)((8u * kRowBytes) >> 4) << 32;    // stride byte offset = 8 rows, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell)
    d |= (uint64_t)(BK == 32 ? 2 : 4) << 61;         // layout type SWIZZLE_128B (2) / SWIZZLE_64B (4)
    return d;
}
// instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, M = 128, N = 256
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// 256-bit global accesses (LDG/STG.E.256): a lane moves one full 32-byte sector of its row per instruction, which
// halves the L1->XBAR request count of the row-per-lane epilogue
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ldg256(const float *p) {
    f8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(float *p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// Persistent kernel: CTA b processes tiles b, b + grid, ...; the fp32 accumulator is double buffered in TMEM
// (2 x 256 columns) so the epilogue of tile i overlaps the TMA/MMA main loop of tile i + 1.
__global__ void __launch_bounds__(kThreads, 1)
dense_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     const float *__restrict__ a_hi, const float *__restrict__ a_lo, float *__restrict__ n_hi,
                     float *__restrict__ n_lo, float *__restrict__ mom, float *__restrict__ scal, int64_t M, int D, float eps,
                     int mode, int n_tiles, int n_nblocks) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzled tiles need 1024 B alignment
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = D / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;  // running k-block counter across tiles (stage = it % kStages)
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int n0 = (tile % n_nblocks) * BN;
                const int m0 = (tile / n_nblocks) * BM;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t st = smem_base + s * kStageBytes;
                    mbar_expect_tx(full_bar(s), kStageBytes);
                    tma_load_2d(st, &map_a_hi, full_bar(s), kb * BK, m0);
                    tma_load_2d(st + kABytes, &map_a_lo, full_bar(s), kb * BK, m0);
                    tma_load_2d(st + 2 * kABytes, &map_b_hi, full_bar(s), kb * BK, n0);
                    tma_load_2d(st + 2 * kABytes + kBBytes, &map_b_lo, full_bar(s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            uint32_t it = 0, lt = 0;  // lt = local tile counter (accumulator stage = lt & 1)
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                mbar_wait(tmem_empty_bar(as), aph ^ 1u);   // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * (uint32_t)BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t st = smem_base + s * kStageBytes;
                    const uint64_t da_hi = make_smem_desc(st), da_lo = make_smem_desc(st + kABytes);
                    const uint64_t db_hi = make_smem_desc(st + 2 * kABytes), db_lo = make_smem_desc(st + 2 * kABytes + kBBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // advance inside the swizzle row
                        tcgen05_mma_tf32(tmem_d, da_hi + koff, db_hi + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                        tcgen05_mma_tf32(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                        tcgen05_mma_tf32(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
                    }
                    tcgen05_commit(empty_bar(s));   // frees the stage once the MMAs above have read it
                }
                tcgen05_commit(tmem_full_bar(as));  // accumulator complete
            }
        }
    } else {
        // ===== epilogue (8 warps): TMEM -> registers -> fused leapfrog -> global =====
        const int ew = warp - 2;
        const int q = warp & 3;                 // TMEM lane quarter this warp may access (warp id % 4)
        const int half = ew >> 2;               // which 128-column half of the tile this warp handles
        const int row = q * 32 + lane;
        const float eps_half = eps * 0.5f;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int n0 = (tile % n_nblocks) * BN;
            const int64_t m = (int64_t)(tile / n_nblocks) * BM + row;
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            const bool row_ok = m < M;
            float quad = 0.f, ke = 0.f;
            bool waited = false;
#pragma unroll 1
            for (int chunk = 0; chunk < BN / 64; ++chunk) {
                const int col = half * (BN / 2) + chunk * 32;
                const int64_t off = (row_ok ? m : 0) * D + n0 + col;
                // issue the global loads of this chunk before waiting on the accumulator
                f8 h8[4], l8[4], p8[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h8[j] = ldg256(a_hi + off + 8 * j);
                    l8[j] = ldg256(a_lo + off + 8 * j);
                    p8[j] = ldg256(mom + off + 8 * j);
                }
                if (!waited) {
                    mbar_wait(tmem_full_bar(as), aph);
                    tcgen05_fence_after();
                    waited = true;
                }
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN + (uint32_t)col, r);
                if (chunk == BN / 64 - 1) {
                    // all TMEM reads of this warp for this accumulator are done: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(as)) : "memory");
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float pp[8], hi[8], lo[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float z = __uint_as_float(r[8 * j + e]);
                            const float dl = h8[j].v[e] + l8[j].v[e];
                            const float gh = -z * eps_half;
                            float dn;
                            if (mode != kModeMid) quad = fmaf(z, dl, quad);
                            if (mode == kModeFirst) {
                                pp[e] = p8[j].v[e] + gh;
                                dn = fmaf(eps, pp[e], dl);
                            } else if (mode == kModeMid) {
                                pp[e] = (p8[j].v[e] + gh) + gh;
                                dn = fmaf(eps, pp[e], dl);
                            } else {
                                pp[e] = p8[j].v[e] + gh;
                                ke = fmaf(pp[e], pp[e], ke);
                                dn = dl;
                            }
                            hi[e] = tf32_hi(dn);
                            lo[e] = dn - hi[e];
                        }
                        stg256(mom + off + 8 * j, pp);
                        if (mode != kModeLast) {
                            stg256(n_hi + off + 8 * j, hi);
                            stg256(n_lo + off + 8 * j, lo);
                        }
                    }
                }
            }
            if (row_ok && mode != kModeMid) {
                atomicAdd(scal + (mode == kModeFirst ? 1 : 3) * M + m, quad);
                if (mode == kModeLast) atomicAdd(scal + 2 * M + m, ke);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ================================================================== CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) share one 256 x 256 tile: CTA r stages its own 128 rows of A and
// rows [128 r, 128 r + 128) of the B tile, the leader issues tcgen05.mma.cta_group::2 (M = 256), each tensor core reads
// both halves of B, and every CTA keeps its 128 x 256 accumulator in its own TMEM.  Per SM this cuts the bytes staged
// per k-block from 48 KB to 32 KB (L2 -> SM traffic and shared-memory fill bandwidth, the limiters of the 1-CTA
// kernel), which also leaves room for 6 instead of 4 stages.
constexpr uint32_t kBHalfBytes = kBBytes / 2;
constexpr uint32_t kStageBytes2 = 2 * kABytes + 2 * kBHalfBytes;
constexpr int kStages2 = (BK == 32 ? 3 : 6);
// + one 32 x 32 float transposition tile (8 column groups of 132 floats) per epilogue warp (coalesced epilogue of the mixed-split kernel)
constexpr uint32_t kEpiTileFloats = 8 * 132;
constexpr uint32_t kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + kEpiWarps * kEpiTileFloats * 4;
static_assert(kSmemBytes2 <= 232448, "CTA-pair kernel: shared memory over the 227 KB limit");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar, uint16_t cta_mask = 3) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}
// the same load delivered to every CTA of cta_mask at the same CTA-relative offset; the completion bytes of each copy are
// signalled on the full barrier of the RECEIVING CTA's pair leader (barrier address with the peer bit cleared, as CUTLASS'
// SM100_TMA_2SM_LOAD_MULTICAST does)
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1,
                                                    uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                      uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}
// kind::f16 with BF16 operands (format 1), F32 accumulator; UMMA_K = 16 elements = the same 32 bytes per K step
__device__ __forceinline__ void tcgen05_mma_bf16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                      uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair_bf16() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_addr, float hi_addr) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo_addr, hi_addr);   // .x (low half) = element at the lower address
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ void stg256u(uint32_t *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// kMixed: map_a_lo / map_b_lo are the bf16 cross-term operands ([rows, 2 D], 64-element boxes), a_hi = the full-precision
// Delta (read by the epilogue), a_lo unused, n_hi = next TF32 hi (nullptr: not stored, MMC_TC_HW_TRUNC), n_lo = next
// full-precision Delta, n_x = next bf16 cross-term operand
// kChain: ONE launch runs all L + 1 GEMMs of a transition.  A tile of leapfrog l only depends on the column tiles of its own
// 256-row block at leapfrog l - 1 (they wrote the block's next Delta), so instead of a launch boundary the TMA producers wait
// on a per-(leapfrog, row block) counter that the epilogue warps bump after their stores.  The TMA / MMA / epilogue pipeline
// (stage ring, double-buffered TMEM) then runs across leapfrogs without draining: no launch gap, no prologue and no
// un-overlapped last epilogue per GEMM (about 11 us of the 42 us a 4,096-chain shard spends per leapfrog).
struct ChainCtl {
    const float *dfull[2];   // full-precision Delta ping-pong (leapfrog l reads [l & 1], writes [(l & 1) ^ 1])
    float *dfull_w[2];
    float *hi[2];            // TF32 copies of Delta (nullptr: the tensor core truncates, see tc_hw_truncates)
    uint32_t *x[2];          // bf16 cross-term operands
    int *ready;              // [n_leap + 2][n_mblocks] epilogue-warp arrivals per 256-row block
    int n_leap;              // L: leapfrog l = 0 is the first gradient, l = L the last half kick
    int n_mblocks;
};
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// kCoal (with kMixed): the epilogue transposes the accumulator chunk through shared memory so that every global access of
// a warp covers whole 128-byte lines of ONE row (lane = column) instead of one 32-byte sector of 32 different rows
// (lane = row, the TMEM layout): ncu showed the L1TEX -> XBAR request path as the busiest unit (67 %) with the row-per-lane
// epilogue issuing 20,480 sector requests per tile next to the 16,384 line requests of the TMA operand loads.
// kQuad: clusters of FOUR CTAs = two CTA pairs that work on the same 256 columns of two consecutive 256-row blocks.  The
// B half a CTA needs is the same in both pairs, so each CTA loads 64 of its 128 rows and multicasts them to its twin in the
// other pair: 48 instead of 64 KB per CTA and k-block come from L2, the feed that bounds the kernel.  A
// stage is reused only after BOTH pairs have consumed it (empty barriers count two multicast commits).
template <bool kMixed, bool kCoal, bool kQuad = false, bool kChain = false>
__global__ void __launch_bounds__(kThreads, 1)
dense_gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                          const float *__restrict__ a_hi, const float *__restrict__ a_lo, float *__restrict__ n_hi,
                          float *__restrict__ n_lo, uint32_t *__restrict__ n_x, float *__restrict__ mom, float *__restrict__ scal,
                          int64_t M, int D, float eps, int mode, int n_tiles, int n_nblocks,
                          const __grid_constant__ CUtensorMap map_a_hi1, const __grid_constant__ CUtensorMap map_a_lo1,
                          const ChainCtl ctl) {
    static_assert(!kChain || (kMixed && kCoal && !kQuad), "the leapfrog chain exists for the mixed-split CTA-pair kernel");
    const int n_l = kChain ? ctl.n_leap + 1 : 1;   // GEMMs of this launch
    static_assert(!kMixed || BK == 32, "the mixed split is written for 128-byte swizzle rows");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages2 * kStageBytes2;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages2 + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages2 + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages2 + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages2 + 4);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = D / BK;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = kQuad ? (crank & 1u) : crank;   // rank inside the CTA pair
    const uint32_t psel = kQuad ? (crank >> 1) : 0u;      // which pair of the cluster
    const uint32_t lrank = crank & ~1u;                   // cluster rank of this pair's leader
    const bool leader = rank == 0;
    constexpr int kCl = kQuad ? 4 : 2;
    const int pair = blockIdx.x / kCl, n_pairs = gridDim.x / kCl;   // cluster index / count: the tile loop strides by clusters
    auto pair_block = [&](int tile) { return (tile / n_nblocks) * (kQuad ? 2 : 1) + (int)psel; };   // 256-row block of this pair

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(full_bar(s), 1);    // the leader's producer arrives with the bytes of BOTH CTAs
            mbar_init(empty_bar(s), kQuad ? 2 : 1);   // multicast tcgen05.commit of the leader (kQuad: of both pairs' leaders)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 2 * kEpiWarps);   // epilogue warps of both CTAs (used in the leader only)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer (one per CTA; both signal the leader's full barrier) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int l = 0; l < n_l; ++l) {
            const CUtensorMap *ma_hi = (kChain && (l & 1)) ? &map_a_hi1 : &map_a_hi;
            const CUtensorMap *ma_lo = (kChain && (l & 1)) ? &map_a_lo1 : &map_a_lo;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int n0 = (tile % n_nblocks) * BN + (int)rank * (BN / 2);
                const int m0 = pair_block(tile) * (2 * BM) + (int)rank * BM;
                if (kChain && l > 0) {
                    // every epilogue warp of both CTAs of every column tile of this row block has stored leapfrog l - 1
                    const int *flag = ctl.ready + (int64_t)l * ctl.n_mblocks + pair_block(tile);
                    const int target = n_nblocks * 2 * kEpiWarps;
                    while (ld_acquire_gpu(flag) < target) __nanosleep(40);
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores observed above -> the TMA (async proxy) reads below
                }
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % kStages2;
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t st = smem_base + s * kStageBytes2;
                    const uint32_t lbar = mapa_shared(full_bar(s), lrank);
                    if (leader) mbar_expect_tx(full_bar(s), 2 * kStageBytes2);
                    const int kx = kMixed ? kb * 2 * BK : kb * BK;   // bf16 operands: 2 BK elements (hi | lo) per k-block
                    tma_load_2d_pair(st, ma_hi, lbar, kb * BK, m0);
                    tma_load_2d_pair(st + kABytes, ma_lo, lbar, kx, m0);
                    if constexpr (kQuad) {
                        // rows [64 psel, 64 psel + 64) of this CTA's B half, delivered to this CTA and to its twin in the other pair
                        const uint16_t mc = (uint16_t)((1u << crank) | (1u << (crank ^ 2u)));
                        const uint32_t qoff = psel * (kBHalfBytes / 2);
                        tma_load_2d_pair_mc(st + 2 * kABytes + qoff, &map_b_hi, lbar, kb * BK, n0 + (int)psel * (BN / 4), mc);
                        tma_load_2d_pair_mc(st + 2 * kABytes + kBHalfBytes + qoff, &map_b_lo, lbar, kx, n0 + (int)psel * (BN / 4), mc);
                    } else {
                        tma_load_2d_pair(st + 2 * kABytes, &map_b_hi, lbar, kb * BK, n0);
                        tma_load_2d_pair(st + 2 * kABytes + kBHalfBytes, &map_b_lo, lbar, kx, n0);
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_pair();
            uint32_t it = 0, lt = 0;
            for (int l = 0; l < n_l; ++l)
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
                const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                mbar_wait(tmem_empty_bar(as), aph ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * (uint32_t)BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % kStages2;
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t st = smem_base + s * kStageBytes2;
                    const uint64_t da_hi = make_smem_desc(st), da_lo = make_smem_desc(st + kABytes);
                    const uint64_t db_hi = make_smem_desc(st + 2 * kABytes), db_lo = make_smem_desc(st + 2 * kABytes + kBHalfBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
                        tcgen05_mma_tf32_pair(tmem_d, da_hi + koff, db_hi + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                        if constexpr (kMixed) {
                            // [bf16 a_hi | bf16 a_lo] . [bf16 b_lo | bf16 b_hi]^T: 16 bf16 = the same 32-byte K step
                            tcgen05_mma_bf16_pair(tmem_d, da_lo + koff, db_lo + koff, make_idesc_pair_bf16(), 1u);
                        } else {
                            tcgen05_mma_tf32_pair(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                            tcgen05_mma_tf32_pair(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
                        }
                    }
                    tcgen05_commit_pair(empty_bar(s), kQuad ? (uint16_t)0xF : (uint16_t)3);   // frees the stage in every CTA that fills it
                }
                tcgen05_commit_pair(tmem_full_bar(as), (uint16_t)(3u << lrank));   // accumulators of both CTAs of this pair complete
            }
        }
    } else {
        // ===== epilogue (8 warps per CTA): own 128 rows x 256 columns =====
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const float eps_half = eps * 0.5f;
        uint32_t lt = 0;
        if constexpr (kMixed && kCoal) {
            // transposition tile of this warp: 8 column groups x (32 rows x 4 columns + 4 floats of padding); lane = row writes
            // float4 (conflict-free per quarter warp), lane = (row % 4 group, column group) reads float4 (conflict-free)
            float *zt = reinterpret_cast<float *>(smem_raw + (bar_base + 256u - smem_u32(smem_raw))) + ew * kEpiTileFloats;
            const int cg = lane & 7, rg = lane >> 3;   // this lane's 4 columns (4 cg ..) and its row inside a group of 4 rows
            const float *const a_hi_arg = a_hi;
            float *const n_lo_arg = n_lo, *const n_hi_arg = n_hi;
            uint32_t *const n_x_arg = n_x;
            auto run_tiles = [&](auto mode_c, int l) {
                constexpr int kMode = decltype(mode_c)::value;
                // this GEMM's operands: the launch arguments, or leapfrog l of the chain
                const float *const a_hi = kChain ? ctl.dfull[l & 1] : a_hi_arg;
                float *const n_lo = kChain ? ctl.dfull_w[(l & 1) ^ 1] : n_lo_arg;
                float *const n_hi = kChain ? ctl.hi[(l & 1) ^ 1] : n_hi_arg;
                uint32_t *const n_x = kChain ? ctl.x[(l & 1) ^ 1] : n_x_arg;
                for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
                    const int n0 = (tile % n_nblocks) * BN;
                    const int64_t m_base = (int64_t)pair_block(tile) * (2 * BM) + (int64_t)rank * BM + q * 32;   // row of TMEM lane 0
                    const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                    float quad = 0.f, ke = 0.f;   // lane = row (m_base + lane), as in the accumulator layout
                    bool waited = false;
#pragma unroll 1
                    for (int chunk = 0; chunk < BN / 64; ++chunk) {
                        const int col = half * (BN / 2) + chunk * 32;
                        // iteration it handles rows 4 it + rg: one 128-byte line per row and array, 4 rows per warp access
                        const int64_t off0 = (m_base + rg) * D + n0 + col + 4 * cg;
                        const int64_t rows_left = M - m_base - rg;   // row 4 it + rg exists iff 4 it < rows_left
                        float4 dl[8], pm[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int64_t o = (4 * it < rows_left) ? off0 + (int64_t)(4 * it) * D : (int64_t)(n0 + col + 4 * cg);
                            dl[it] = __ldg(reinterpret_cast<const float4 *>(a_hi + o));
                            pm[it] = *reinterpret_cast<const float4 *>(mom + o);
                        }
                        if (!waited) {
                            mbar_wait(tmem_full_bar(as), aph);
                            tcgen05_fence_after();
                            waited = true;
                        }
                        {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN + (uint32_t)col, r);
                            if (chunk == BN / 64 - 1) {
                                tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) {
                                    const uint32_t lbar = mapa_shared(tmem_empty_bar(as), lrank);
                                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(lbar) : "memory");
                                }
                            }
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4)
                                *reinterpret_cast<float4 *>(zt + c4 * 132 + lane * 4) =
                                    make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]), __uint_as_float(r[4 * c4 + 2]),
                                                __uint_as_float(r[4 * c4 + 3]));
                        }
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const float4 z4 = *reinterpret_cast<const float4 *>(zt + cg * 132 + (4 * it + rg) * 4);
                            const float z[4] = {z4.x, z4.y, z4.z, z4.w};
                            const float d[4] = {dl[it].x, dl[it].y, dl[it].z, dl[it].w};
                            const float pin[4] = {pm[it].x, pm[it].y, pm[it].z, pm[it].w};
                            float pp[4], dn[4], hi[4], lo[4];
                            float qs = 0.f, ks = 0.f;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float gh = -z[e] * eps_half;
                                if (kMode == kModeMid) pp[e] = (pin[e] + gh) + gh;
                                else pp[e] = pin[e] + gh;
                                dn[e] = kMode == kModeLast ? d[e] : fmaf(eps, pp[e], d[e]);
                                hi[e] = tf32_hi(dn[e]);
                                lo[e] = dn[e] - hi[e];
                                if (kMode != kModeMid) qs = fmaf(z[e], d[e], qs);
                                if (kMode == kModeLast) ks = fmaf(pp[e], pp[e], ks);
                            }
                            if (kMode != kModeMid) {
                                // row sums over the 8 lanes of a row; lane L owns row L: computed in iteration L / 4 by group L % 4
#pragma unroll
                                for (int o = 4; o > 0; o >>= 1) {
                                    qs += __shfl_xor_sync(0xffffffffu, qs, o);
                                    if (kMode == kModeLast) ks += __shfl_xor_sync(0xffffffffu, ks, o);
                                }
                                const float qrow = __shfl_sync(0xffffffffu, qs, 8 * (lane & 3));
                                const float krow = kMode == kModeLast ? __shfl_sync(0xffffffffu, ks, 8 * (lane & 3)) : 0.f;
                                if ((lane >> 2) == it) { quad += qrow; ke += krow; }
                            }
                            if (4 * it < rows_left) {
                                const int64_t o = off0 + (int64_t)(4 * it) * D;
                                *reinterpret_cast<float4 *>(mom + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
                                if (kMode != kModeLast) {
                                    *reinterpret_cast<float4 *>(n_lo + o) = make_float4(dn[0], dn[1], dn[2], dn[3]);
                                    if (n_hi) *reinterpret_cast<float4 *>(n_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                                    // bf16 row block of this k-block: words 0..15 = (hi, hi) pairs, 16..31 = (lo, lo) pairs
                                    uint32_t *xw = n_x + (o - 4 * cg);
                                    *reinterpret_cast<uint2 *>(xw + 2 * cg) = make_uint2(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]));
                                    *reinterpret_cast<uint2 *>(xw + 16 + 2 * cg) = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
                                }
                            }
                        }
                        __syncwarp();   // the tile is rewritten by the next chunk
                    }
                    const int64_t m = m_base + lane;
                    if (m < M && kMode != kModeMid) {
                        atomicAdd(scal + (kMode == kModeFirst ? 1 : 3) * M + m, quad);
                        if (kMode == kModeLast) atomicAdd(scal + 2 * M + m, ke);
                    }
                    if constexpr (kChain) {
                        // release this warp's part of the row block to leapfrog l + 1 (every lane fences its own stores)
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) atomicAdd(ctl.ready + (int64_t)(l + 1) * ctl.n_mblocks + pair_block(tile), 1);
                    }
                }
            };
            for (int l = 0; l < n_l; ++l) {
                const int mode_l = kChain ? (l == 0 ? (int)kModeFirst : (l == ctl.n_leap ? (int)kModeLast : (int)kModeMid)) : mode;
                if (mode_l == kModeMid) run_tiles(std::integral_constant<int, kModeMid>{}, l);
                else if (mode_l == kModeFirst) run_tiles(std::integral_constant<int, kModeFirst>{}, l);
                else run_tiles(std::integral_constant<int, kModeLast>{}, l);
            }
        } else
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
            const int n0 = (tile % n_nblocks) * BN;
            const int64_t m = (int64_t)pair_block(tile) * (2 * BM) + (int64_t)rank * BM + row;
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            const bool row_ok = m < M;
            float quad = 0.f, ke = 0.f;
            bool waited = false;
#pragma unroll 1
            for (int chunk = 0; chunk < BN / 64; ++chunk) {
                const int col = half * (BN / 2) + chunk * 32;
                const int64_t off = (row_ok ? m : 0) * D + n0 + col;
                f8 h8[4], l8[4], p8[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h8[j] = ldg256(a_hi + off + 8 * j);
                    if constexpr (!kMixed) l8[j] = ldg256(a_lo + off + 8 * j);
                    p8[j] = ldg256(mom + off + 8 * j);
                }
                if (!waited) {
                    mbar_wait(tmem_full_bar(as), aph);
                    tcgen05_fence_after();
                    waited = true;
                }
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN + (uint32_t)col, r);
                if (chunk == BN / 64 - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t lbar = mapa_shared(tmem_empty_bar(as), lrank);
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(lbar) : "memory");
                    }
                }
                if (row_ok) {
                    uint32_t xh[16], xl[16];   // mixed split: this chunk's 32 bf16(hi) and 32 bf16(lo), packed in pairs
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float pp[8], hi[8], lo[8], dnx[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float z = __uint_as_float(r[8 * j + e]);
                            const float dl = kMixed ? h8[j].v[e] : h8[j].v[e] + l8[j].v[e];
                            const float gh = -z * eps_half;
                            float dn;
                            if (mode != kModeMid) quad = fmaf(z, dl, quad);
                            if (mode == kModeFirst) {
                                pp[e] = p8[j].v[e] + gh;
                                dn = fmaf(eps, pp[e], dl);
                            } else if (mode == kModeMid) {
                                pp[e] = (p8[j].v[e] + gh) + gh;
                                dn = fmaf(eps, pp[e], dl);
                            } else {
                                pp[e] = p8[j].v[e] + gh;
                                ke = fmaf(pp[e], pp[e], ke);
                                dn = dl;
                            }
                            hi[e] = tf32_hi(dn);
                            lo[e] = dn - hi[e];
                            dnx[e] = dn;
                        }
                        stg256(mom + off + 8 * j, pp);
                        if (mode != kModeLast) {
                            if constexpr (kMixed) {
                                stg256(n_lo + off + 8 * j, dnx);             // full-precision Delta
                                if (n_hi) stg256(n_hi + off + 8 * j, hi);   // TF32 operand
#pragma unroll
                                for (int e = 0; e < 8; e += 2) {
                                    xh[4 * j + e / 2] = pack_bf16x2(hi[e], hi[e + 1]);
                                    xl[4 * j + e / 2] = pack_bf16x2(lo[e], lo[e + 1]);
                                }
                            } else {
                                stg256(n_hi + off + 8 * j, hi);
                                stg256(n_lo + off + 8 * j, lo);
                            }
                        }
                    }
                    if constexpr (kMixed) {
                        if (mode != kModeLast) {
                            // bf16 row m of [M, 2 D]: k-block (n0 + col) / 32 holds 32 hi then 32 lo = 32 words at word m D + n0 + col
                            uint32_t *xw = n_x + off;
                            uint32_t v[8];
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = xh[8 * g + e];
                                stg256u(xw + 8 * g, v);
                            }
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = xl[8 * g + e];
                                stg256u(xw + 16 + 8 * g, v);
                            }
                        }
                    }
                }
            }
            if (row_ok && mode != kModeMid) {
                atomicAdd(scal + (mode == kModeFirst ? 1 : 3) * M + m, quad);
                if (mode == kModeLast) atomicAdd(scal + 2 * M + m, ke);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer can still read its shared memory / signal it
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

__global__ void split_kernel(const float *__restrict__ src, float *__restrict__ hi, float *__restrict__ lo, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4 *>(src)[i];
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    reinterpret_cast<float4 *>(hi)[i] = h;
    reinterpret_cast<float4 *>(lo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// mixed split of a [rows, D] matrix: hi (TF32 truncation, fp32, optional) and the bf16 cross-term operand [rows, 2 D] whose
// k-block kb (32 columns) holds 32 x bf16(hi) then 32 x bf16(lo) (b_order: lo first, so that the K-concatenated product
// of an A row and a B row is a_hi b_lo + a_lo b_hi).  One thread per 8 consecutive elements.
__global__ void split_mixed_kernel(const float *__restrict__ src, float *__restrict__ hi, uint32_t *__restrict__ x, int64_t n8,
                                   int b_order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 v0 = reinterpret_cast<const float4 *>(src)[2 * i], v1 = reinterpret_cast<const float4 *>(src)[2 * i + 1];
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        h[e] = tf32_hi(v[e]);
        l[e] = v[e] - h[e];
    }
    if (hi) {
        reinterpret_cast<float4 *>(hi)[2 * i] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4 *>(hi)[2 * i + 1] = make_float4(h[4], h[5], h[6], h[7]);
    }
    // element index 8 i = row D + c: word index of the pair (c, c + 1) inside the row's k-block = row D + (c / 32) 32 + (c % 32) / 2
    const int64_t e0 = 8 * i;
    const int64_t blk = e0 >> 5;           // (row, k-block) index: D % 32 == 0
    const int w = (int)(e0 & 31) >> 1;     // first of 4 words inside the 16-word half
    uint32_t *xb = x + blk * 32;
    uint4 ph = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    uint4 pl = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    *reinterpret_cast<uint4 *>(xb + (b_order ? 16 : 0) + w) = ph;
    *reinterpret_cast<uint4 *>(xb + (b_order ? 0 : 16) + w) = pl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_2d(EncodeTiledFn fn, CUtensorMap *map, float *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return MMC_ERR_CUDA;
    }
    return MMC_OK;
}

// bf16 [rows, 2 cols] cross-term operand: boxes of 2 BK = 64 elements (one 128-byte swizzle row) x box_rows
int encode_2d_bf16(EncodeTiledFn fn, CUtensorMap *map, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {2 * cols, rows};
    const cuuint64_t strides[1] = {2 * cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * BK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r);
        return MMC_ERR_CUDA;
    }
    return MMC_OK;
}

}  // namespace tc

int dense_tc_prepare(DenseState *st);

// One-off self-test per process: the same short mixed-split run with the pre-truncated TF32 copy and with the
// full-precision Delta as the kind::tf32 operand; 1 = bit-identical (the hardware truncates), 0 = not, cached.
static std::atomic<int> g_hw_trunc{-1};   // -1 unknown, -2 probing
static bool tc_hw_truncates() {
    const int seen = g_hw_trunc.load();
    if (seen >= 0) return seen == 1;
    if (seen == -2) return false;   // the probe's own handles (or a second thread while the probe runs: it keeps the TF32 copy)
    g_hw_trunc = -2;
    int verdict = 0;
    const int D = 256;
    const int64_t M = 256;
    std::vector<float> mean(D), prec((size_t)D * D), init((size_t)M * D);
    uint32_t lcg = 12345u;
    auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return (float)(lcg >> 8) * (1.0f / 16777216.0f) - 0.5f; };
    for (auto &v : mean) v = rnd();
    for (int i = 0; i < D; ++i)
        for (int j = 0; j <= i; ++j) prec[(size_t)i * D + j] = prec[(size_t)j * D + i] = (i == j ? 2.0f : 0.0f) + 0.1f * rnd();
    for (auto &v : init) v = 3.0f * rnd();
    mmc_target_desc t{};
    t.dim = D;
    t.vec = mean.data();
    t.mat = prec.data();
    DenseState *ps = nullptr;
    float *d_pos = nullptr, *d_out = nullptr;
    unsigned long long *d_acc = nullptr;
    std::vector<float> out[2];
    bool ok = dense_create(&ps, &t, M) == MMC_OK && cudaMalloc((void **)&d_pos, init.size() * 4) == cudaSuccess &&
              cudaMalloc((void **)&d_out, init.size() * 4) == cudaSuccess && cudaMalloc((void **)&d_acc, 8) == cudaSuccess;
    for (int v = 0; v < 2 && ok; ++v) {
        ok = cudaMemcpy(d_pos, init.data(), init.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemset(d_acc, 0, 8) == cudaSuccess;
        DenseRunArgs a{};
        a.positions = d_pos; a.out = d_out; a.accept_count = d_acc;
        a.chains = M; a.n_collect = 1; a.out_pitch = 1; a.eps = 0.05f; a.n_leapfrog = 3; a.seed = 99; a.gemm_path = 3;
        if (ok && v == 0) ok = dense_tc_prepare(ps) == MMC_OK;   // sees g_hw_trunc == -2 -> tc_hw_trunc = false
        if (ok) {
            ps->tc_hw_trunc = v == 1;
            ok = dense_run(ps, a, nullptr) == MMC_OK && cudaDeviceSynchronize() == cudaSuccess;
        }
        if (ok) {
            out[v].resize(init.size());
            ok = cudaMemcpy(out[v].data(), d_out, init.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
    }
    if (ok) {
        bool moved = false;
        for (size_t i = 0; i < init.size() && !moved; ++i) moved = out[0][i] != init[i];
        verdict = moved && memcmp(out[0].data(), out[1].data(), init.size() * 4) == 0 ? 1 : 0;
    }
    cudaFree(d_pos); cudaFree(d_out); cudaFree(d_acc);
    dense_destroy(ps);
    g_hw_trunc = verdict;
    return verdict == 1;
}
// 1 / 0: what the self-test found on this device (diagnostics: mmc_hmc_gemm_info)
int dense_tc_hw_trunc_state() { return g_hw_trunc.load(); }

int dense_tc_prepare(DenseState *st) {
    if (st->tc) return MMC_OK;
    const int D = st->Dp;   // internal pitch (whole tiles)
    const int64_t M = st->chains;
    MMC_REQUIRE(D % tc::BN == 0, "tcgen05 path needs dim %% 256 == 0, got %d", D);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MMC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MMC_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    const size_t md = (size_t)M * D * sizeof(float), dd = (size_t)D * D * sizeof(float);
    MMC_CUDA(cudaMalloc((void **)&st->d_prec_split, 2 * dd));
    MMC_CUDA(cudaMalloc((void **)&st->d_delta_split[0], 2 * md));
    MMC_CUDA(cudaMalloc((void **)&st->d_delta_split[1], 2 * md));
    // mixed split: bf16 cross-term operands, [rows, 2 D] bf16 = 4 bytes per element of the matrix
    MMC_CUDA(cudaMalloc((void **)&st->d_prec_x, dd));
    MMC_CUDA(cudaMalloc((void **)&st->d_delta_x[0], md));
    MMC_CUDA(cudaMalloc((void **)&st->d_delta_x[1], md));
    MMC_CUDA(cudaMemset(st->d_delta_x[0], 0, md));
    MMC_CUDA(cudaMemset(st->d_delta_x[1], 0, md));
    const int64_t n4 = (int64_t)D * D / 4;
    tc::split_kernel<<<(unsigned)((n4 + 255) / 256), 256>>>(st->d_prec, st->d_prec_split, st->d_prec_split + (size_t)D * D, n4);
    MMC_CUDA(cudaGetLastError());
    tc::split_mixed_kernel<<<(unsigned)((n4 / 2 + 255) / 256), 256>>>(st->d_prec, nullptr, st->d_prec_x, n4 / 2, 1);
    MMC_CUDA(cudaGetLastError());
    MMC_CUDA(cudaDeviceSynchronize());
    tc::Maps *maps = new tc::Maps();
    int rc = MMC_OK;
    const tc::EncodeTiledFn enc = (tc::EncodeTiledFn)fn;
    for (int b = 0; b < 2 && !rc; ++b) {
        rc = tc::encode_2d(enc, &maps->a_hi[b], st->d_delta_split[b], (uint64_t)M, (uint64_t)D, tc::BM);
        if (!rc) rc = tc::encode_2d(enc, &maps->a_lo[b], st->d_delta_split[b] + (size_t)M * D, (uint64_t)M, (uint64_t)D, tc::BM);
        if (!rc) rc = tc::encode_2d_bf16(enc, &maps->a_x[b], st->d_delta_x[b], (uint64_t)M, (uint64_t)D, tc::BM);
        if (!rc) rc = tc::encode_2d(enc, &maps->a_full[b], st->d_delta[b], (uint64_t)M, (uint64_t)D, tc::BM);
    }
    if (!rc) rc = tc::encode_2d(enc, &maps->b_hi, st->d_prec_split, (uint64_t)D, (uint64_t)D, tc::BN);
    if (!rc) rc = tc::encode_2d(enc, &maps->b_lo, st->d_prec_split + (size_t)D * D, (uint64_t)D, (uint64_t)D, tc::BN);
    if (!rc) rc = tc::encode_2d(enc, &maps->b_hi_half, st->d_prec_split, (uint64_t)D, (uint64_t)D, tc::BN / 2);
    if (!rc) rc = tc::encode_2d(enc, &maps->b_lo_half, st->d_prec_split + (size_t)D * D, (uint64_t)D, (uint64_t)D, tc::BN / 2);
    if (!rc) rc = tc::encode_2d_bf16(enc, &maps->b_x_half, st->d_prec_x, (uint64_t)D, (uint64_t)D, tc::BN / 2);
    if (!rc) rc = tc::encode_2d(enc, &maps->b_hi_q, st->d_prec_split, (uint64_t)D, (uint64_t)D, tc::BN / 4);
    if (!rc) rc = tc::encode_2d_bf16(enc, &maps->b_x_q, st->d_prec_x, (uint64_t)D, (uint64_t)D, tc::BN / 4);
    if (rc) { delete maps; return rc; }
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_pair_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    MMC_CUDA(cudaFuncSetAttribute(tc::dense_gemm_tc_pair_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    {   // the chain kernel's producers wait on other CTAs: all of its clusters must be resident at once
        cudaLaunchConfig_t qc{};
        qc.gridDim = dim3((unsigned)(sm_count() / 2 * 2));
        qc.blockDim = dim3(tc::kThreads);
        qc.dynamicSmemBytes = tc::kSmemBytes2;
        cudaLaunchAttribute qa{};
        qa.id = cudaLaunchAttributeClusterDimension;
        qa.val.clusterDim.x = 2; qa.val.clusterDim.y = 1; qa.val.clusterDim.z = 1;
        qc.attrs = &qa;
        qc.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tc::dense_gemm_tc_pair_kernel<true, true, false, true>, &qc) == cudaSuccess) st->tc_chain_pairs = n;
        else (void)cudaGetLastError();
    }
    {   // how many 4-CTA clusters of the quad kernel are resident at once (GPCs whose SM count is not a multiple of 4 strand SMs)
        cudaLaunchConfig_t qc{};
        qc.gridDim = dim3((unsigned)(sm_count() / 4 * 4));
        qc.blockDim = dim3(tc::kThreads);
        qc.dynamicSmemBytes = tc::kSmemBytes2;
        cudaLaunchAttribute qa{};
        qa.id = cudaLaunchAttributeClusterDimension;
        qa.val.clusterDim.x = 4; qa.val.clusterDim.y = 1; qa.val.clusterDim.z = 1;
        qc.attrs = &qa;
        qc.numAttrs = 1;
        int n = 0;
        const cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, tc::dense_gemm_tc_pair_kernel<true, true, true>, &qc);
        if (qe == cudaSuccess) st->tc_quad_clusters = n;
        else (void)cudaGetLastError();
        if (getenv("MMC_TC_VERBOSE"))
            fprintf(stderr, "[minimcmc] dense tcgen05: resident 4-CTA clusters = %d (%s), hardware tf32 truncation self-test pending\n", n,
                    cudaGetErrorString(qe));
    }
    st->tc = maps;
    // Mixed split only: kind::tf32 can read the full-precision Delta directly when the tensor core ignores the low 13
    // mantissa bits of its fp32 containers, which saves the separate TF32 copy (4 of the epilogue's 24 bytes per element).
    // PTX leaves unconverted inputs unspecified, so this is enabled only after a one-off self-test on this device has
    // produced bit-identical transitions both ways (MMC_TC_HW_TRUNC=0 / 1 overrides the test).
    const char *hw = getenv("MMC_TC_HW_TRUNC");
    if (hw && (hw[0] == '0' || hw[0] == '1')) st->tc_hw_trunc = hw[0] == '1';
    else st->tc_hw_trunc = tc_hw_truncates();
    return MMC_OK;
}

// splits the full-precision Delta written by dense_begin_kernel into buffer 0 of the operand ping-pong
int dense_tc_split_delta(DenseState *st, cudaStream_t stream) {
    if (st->tc_mixed) {
        const int64_t n8 = st->chains * st->Dp / 8;
        tc::split_mixed_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(st->d_delta[0], st->tc_hw_trunc ? nullptr : st->d_delta_split[0],
                                                                                st->d_delta_x[0], n8, 0);
        MMC_CUDA(cudaGetLastError());
        return MMC_OK;
    }
    const int64_t n4 = st->chains * st->Dp / 4;
    float *hi = st->d_delta_split[0];
    tc::split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(st->d_delta[0], hi, hi + (size_t)st->chains * st->Dp, n4);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

int dense_gemm_tc(DenseState *st, int cur, int64_t M, int D, float eps, int mode, cudaStream_t stream) {
    tc::Maps *maps = static_cast<tc::Maps *>(st->tc);
    const size_t md = (size_t)M * D;
    float *a_hi = st->d_delta_split[cur], *a_lo = a_hi + md;
    float *n_hi = st->d_delta_split[cur ^ 1], *n_lo = n_hi + md;
    const int n_nblocks = D / tc::BN;
    if (st->tc_pair) {
        // CTA-pair kernel: 256 x 256 tiles, clusters of two CTAs, persistent over min(tiles, SMs / 2) pairs
        const int n_tiles2 = n_nblocks * (int)((M + 2 * tc::BM - 1) / (2 * tc::BM));
        const int pairs = n_tiles2 < sm_count() / 2 ? n_tiles2 : sm_count() / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(tc::kThreads);
        cfg.dynamicSmemBytes = tc::kSmemBytes2;
        cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if (st->tc_mixed) {
            // 4-CTA clusters (B halves multicast to two pairs): opt-in with MMC_TC_QUAD=1.  Measured on C4 (profiles/
            // r3_dense_quad_cluster.log): 33 clusters are resident (132 of 148 SMs: GPCs whose SM count is not a multiple of 4
            // strand SMs) and 256 tiles of 512 x 256 take 8 waves instead of the 7 of the pair kernel, which cancels the 12 %
            // a tile gains from the smaller operand feed (60.0 vs 58.3-59.0 ms per 4 transitions).
            const char *quad_env = getenv("MMC_TC_QUAD");
            const int n_tiles4 = n_nblocks * (int)((M + 4 * tc::BM - 1) / (4 * tc::BM));
            const bool quad = quad_env && quad_env[0] == '1' && st->tc_quad_clusters > 0 && M >= 4 * tc::BM;
            if (quad) {
                const int clusters = n_tiles4 < st->tc_quad_clusters ? n_tiles4 : st->tc_quad_clusters;
                cfg.gridDim = dim3((unsigned)(4 * clusters));
                attr.val.clusterDim.x = 4;
                MMC_CUDA(cudaLaunchKernelEx(&cfg, tc::dense_gemm_tc_pair_kernel<true, true, true>,
                                            st->tc_hw_trunc ? maps->a_full[cur] : maps->a_hi[cur], maps->a_x[cur], maps->b_hi_q, maps->b_x_q,
                                            (const float *)st->d_delta[cur], (const float *)nullptr,
                                            st->tc_hw_trunc ? (float *)nullptr : n_hi, st->d_delta[cur ^ 1], st->d_delta_x[cur ^ 1], st->d_mom,
                                            st->d_scal, M, D, eps, mode, n_tiles4, n_nblocks, maps->a_full[cur], maps->a_x[cur], tc::ChainCtl{}));
                return MMC_OK;
            }
            static const bool coal = !(getenv("MMC_TC_EPI") && getenv("MMC_TC_EPI")[0] == '0');   // 0: row-per-lane epilogue (A/B)
            auto kern = coal ? tc::dense_gemm_tc_pair_kernel<true, true> : tc::dense_gemm_tc_pair_kernel<true, false>;
            MMC_CUDA(cudaLaunchKernelEx(&cfg, kern, st->tc_hw_trunc ? maps->a_full[cur] : maps->a_hi[cur],
                                        maps->a_x[cur], maps->b_hi_half, maps->b_x_half, (const float *)st->d_delta[cur],
                                        (const float *)nullptr, st->tc_hw_trunc ? (float *)nullptr : n_hi, st->d_delta[cur ^ 1],
                                        st->d_delta_x[cur ^ 1], st->d_mom, st->d_scal, M, D, eps, mode, n_tiles2, n_nblocks,
                                        maps->a_full[cur], maps->a_x[cur], tc::ChainCtl{}));
            return MMC_OK;
        }
        MMC_CUDA(cudaLaunchKernelEx(&cfg, tc::dense_gemm_tc_pair_kernel<false, false>, maps->a_hi[cur], maps->a_lo[cur], maps->b_hi_half,
                                    maps->b_lo_half, (const float *)a_hi, (const float *)a_lo, n_hi, n_lo, (uint32_t *)nullptr, st->d_mom,
                                    st->d_scal, M, D, eps, mode, n_tiles2, n_nblocks, maps->a_hi[cur], maps->a_lo[cur], tc::ChainCtl{}));
        return MMC_OK;
    }
    const int n_tiles = n_nblocks * (int)((M + tc::BM - 1) / tc::BM);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    tc::dense_gemm_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(maps->a_hi[cur], maps->a_lo[cur], maps->b_hi,
                                                                             maps->b_lo, a_hi, a_lo, n_hi, n_lo, st->d_mom,
                                                                             st->d_scal, M, D, eps, mode, n_tiles, n_nblocks);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// All L + 1 GEMMs of one transition in ONE launch of the mixed-split CTA-pair kernel (kChain); Delta ends in d_delta[L & 1].
// Returns MMC_ERR_UNSUPPORTED when the path does not apply (the caller then launches the GEMMs one by one).
int dense_gemm_tc_chain(DenseState *st, int64_t M, int D, float eps, int L, cudaStream_t stream) {
    const char *env = getenv("MMC_TC_CHAIN");
    if (!st->tc || !st->tc_mixed || !st->tc_pair || st->tc_chain_pairs <= 0 || (env && env[0] == '0')) return MMC_ERR_UNSUPPORTED;
    const char *quad_env = getenv("MMC_TC_QUAD");
    if (quad_env && quad_env[0] == '1') return MMC_ERR_UNSUPPORTED;
    tc::Maps *maps = static_cast<tc::Maps *>(st->tc);
    const int n_nblocks = D / tc::BN;
    const int n_mblocks = (int)((M + 2 * tc::BM - 1) / (2 * tc::BM));
    const int n_tiles2 = n_nblocks * n_mblocks;
    const int pairs = n_tiles2 < st->tc_chain_pairs ? n_tiles2 : st->tc_chain_pairs;
    const size_t need = (size_t)(L + 2) * n_mblocks * sizeof(int);
    if (need > st->ready_bytes) {
        cudaFree(st->d_ready);
        st->d_ready = nullptr;
        st->ready_bytes = 0;
        MMC_CUDA(cudaMalloc((void **)&st->d_ready, need));
        st->ready_bytes = need;
    }
    MMC_CUDA(cudaMemsetAsync(st->d_ready, 0, need, stream));
    tc::ChainCtl ctl{};
    for (int b = 0; b < 2; ++b) {
        ctl.dfull[b] = st->d_delta[b];
        ctl.dfull_w[b] = st->d_delta[b];
        ctl.hi[b] = st->tc_hw_trunc ? nullptr : st->d_delta_split[b];
        ctl.x[b] = st->d_delta_x[b];
    }
    ctl.ready = st->d_ready;
    ctl.n_leap = L;
    ctl.n_mblocks = n_mblocks;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = tc::kSmemBytes2;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    // producers of one CTA wait for epilogues of others: the whole grid has to be resident at once, which a cooperative
    // launch guarantees (the kernel never calls grid.sync(); a device that cannot co-schedule it refuses the launch and the
    // caller falls back to one launch per GEMM)
    cudaLaunchAttribute attrs[2];
    attrs[0] = attr;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = st->tc_chain_coop ? 2 : 1;
    const bool hw = st->tc_hw_trunc;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc::dense_gemm_tc_pair_kernel<true, true, false, true>, hw ? maps->a_full[0] : maps->a_hi[0],
                                       maps->a_x[0], maps->b_hi_half, maps->b_x_half, (const float *)nullptr, (const float *)nullptr,
                                       (float *)nullptr, (float *)nullptr, (uint32_t *)nullptr, st->d_mom, st->d_scal, M, D, eps,
                                       (int)kModeMid, n_tiles2, n_nblocks, hw ? maps->a_full[1] : maps->a_hi[1], maps->a_x[1], ctl);
    if (e != cudaSuccess && st->tc_chain_coop) {
        // cooperative + cluster launches are refused on this driver / device: remember it and use plain launches of the GEMMs
        (void)cudaGetLastError();
        st->tc_chain_coop = false;
        st->tc_chain_pairs = 0;
        if (getenv("MMC_TC_VERBOSE")) fprintf(stderr, "[minimcmc] dense tcgen05: cooperative chain launch refused (%s)\n", cudaGetErrorString(e));
        return MMC_ERR_UNSUPPORTED;
    }
    MMC_CUDA(e);
    return MMC_OK;
}

void dense_tc_destroy(DenseState *st) {
    if (st && st->tc) {
        delete static_cast<tc::Maps *>(st->tc);
        st->tc = nullptr;
    }
}

}  // namespace mmc
