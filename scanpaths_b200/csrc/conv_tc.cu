// Implicit-GEMM 3x3 / 5x5 convolution on Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// out[(img*1200 + p)*ldo + col] = inv_scale * sum_{tap,ci} a[img, p+tap, ci] * w[row(col), tap, ci] (+ bias)
//
// The convolutions of the decode path: the 3x3 gate convolutions of the ConvLSTM (ConvLSTM.forward,
// OSIE/models/baseline_attention.py:39-42; per image 1200 pixels x 2048 outputs (4 gates x 512) x
// K = 9 x 512, once per step for h and once per image for x) -- on the product path in their Winograd
// F(2x4,3x3) form (wino_gemm_tc_kernel, second half of this file), as a direct implicit GEMM with
// use_tensor_cores = 2 -- and the 1x1 / 5x5 convolutions of the head (ks = 1: the composed 5x5 -> 2 maps).
//
// Design
//   * D^T = W . A^T: the UMMA "M" side is a tile of 128 OUTPUT CHANNELS (weights), the "N" side a
//     tile of 240 PIXELS = 6 image rows x 40 (N = 240 is a legal UMMA shape), so no row of the
//     128 x 240 tile is padding (1200 = 5 x 240, 2048 = 16 x 128).  For every filter tap the
//     activation operand is ONE 4-D TMA box {64 ch, 40 w, 6 h, 1 img} shifted by the tap offset;
//     rows / columns outside the image are zero-filled by TMA -- that IS the conv padding, there
//     is no im2col and no padded copy.
//   * fp32-equivalent arithmetic on the fp16 pipe: every operand is a pair x = hi + lo / 2^11
//     (11 + 11 significand bits).  Three MMAs per k-step: hi*hi -> accumulator 0,
//     hi*lo + lo*hi -> accumulator 1 (scaled by 2^11), combined in the epilogue.  The dropped
//     lo*lo term is 2^-22 relative.  TMEM: 2 x 240 fp32 columns.
//   * the tensor core adds into its fp32 accumulator with truncation (measured on B200:
//     -1.6e-8 relative per accumulation step, i.e. -4.5e-6 after K = 4608 and -1.2e-5 after
//     K = 12800 -- outside the 1e-5 parity gate).  So accumulator 0 only ever holds ONE filter
//     tap (32 k-steps): after each tap the 8 drain warps read it from TMEM and add it to running
//     totals in registers (round-to-nearest), while the MMA thread already issues the correction
//     MMAs of the next k-block (accumulator 1 is never drained mid-loop: its values weigh 2^-11).
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread) + TMEM allocator,
//     warps 2-9 drain/epilogue (each thread: 1 TMEM lane = 1 output channel x 120 pixels of totals).
//   * persistent: grid = #SMs, every CTA walks tiles blockIdx.x, +gridDim.x, ... (the 16 channel
//     tiles of one pixel tile are adjacent, so the activation tile is shared through L2).  Because
//     the totals live in registers, TMEM is free again as soon as the last tap is drained: the MMA
//     thread starts the next tile while the drain warps run the epilogue of the previous one.
//   * operands staged by TMA with 128-byte swizzle, K-major; kStages-deep mbarrier ring.
#include <cuda.h>

#include "decoder.cuh"

namespace spb {

namespace tc {

#ifndef SPB_TC_BLOCKK
#define SPB_TC_BLOCKK 64
#endif
#ifndef SPB_TC_STAGES
#define SPB_TC_STAGES 2
#endif
constexpr int kTileCh = 128, kTilePix = 240, kBlockK = SPB_TC_BLOCKK, kStages = SPB_TC_STAGES;
constexpr int kSwizzleBytes = kBlockK * 2;                // one K-block row: 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)
static_assert(kSwizzleBytes == 128 || kSwizzleBytes == 64, "BLOCK_K must be 64 or 32 fp16 elements");
constexpr int kWBytes = kTileCh * kBlockK * 2;            // 16 KB per weight operand tile
constexpr int kActBytes = kTilePix * kBlockK * 2;         // 30 KB per activation operand tile
constexpr int kStageBytes = 2 * kWBytes + 2 * kActBytes;  // 92 KB: W_hi, W_lo, A_hi, A_lo
constexpr int kOutRows = 30;                              // pixels per TMA-store box
constexpr int kOutBytes = kOutRows * kTileCh * 4;         // 15 KB staging per pixel half
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2 * kOutBytes;
constexpr int kTmemCols = 512, kCorrCol = 256;
constexpr int kThreads = 320;
constexpr int kChunkKB = 512 / kBlockK;                   // k-blocks per drained chunk: 512 input channels of one filter tap
                                                          // (32 accumulation steps: what the truncation compensation is measured for)
constexpr int kHalfPix = kTilePix / 2;                    // 120 pixels of totals per drain thread
static_assert(kHalfPix % kOutRows == 0, "store boxes must tile the pixel half");

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

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cf. cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B between 8-row groups |
// version 1 | layout SWIZZLE_128B.
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * SWIZZLE_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SWIZZLE_BYTES == 128 ? 2 : 4) << 61;
    return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) { return umma_desc_kmajor<kSwizzleBytes>(saddr); }

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128 (channels), N = 240 (pixels)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTilePix >> 3) << 17) | ((uint32_t)(kTileCh >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int KS, int CIN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                    const __grid_constant__ CUtensorMap tmOut, ConvGemmArgs a, int num_tiles, int nct, int kPT, int rows_per_img) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bar0 + 8 * s; };
    auto empty_bar = [&](int s) { return bar0 + 8 * (kStages + s); };
    const uint32_t main_full_bar = bar0 + 8 * (2 * kStages);        // MMA -> drain warps: one tap accumulated
    const uint32_t main_empty_bar = bar0 + 8 * (2 * kStages + 1);   // drain warps -> MMA: accumulator 0 read out
    const uint32_t corr_empty_bar = bar0 + 8 * (2 * kStages + 2);   // drain warps -> MMA: accumulator 1 read out
    const uint32_t corr_full_bar = bar0 + 8 * (2 * kStages + 3);    // MMA -> drain warps: accumulator 1 of the tile complete
    const uint32_t tmem_slot = bar0 + 8 * (2 * kStages + 4);
    const uint32_t out_smem = bar0 + 256;                         // 2 x 15 KB TMA-store staging

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kKBperTap = CIN / kBlockK;              // CIN = 512 (decoder) or 2048 (the encoder's sal_conv)
    constexpr int kNumKB = KS * KS * kKBperTap;
    constexpr int kNumChunks = kNumKB / kChunkKB;
    constexpr int kPad = KS / 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(main_full_bar, 1);
        mbar_init(main_empty_bar, 8);
        mbar_init(corr_empty_bar, 8);
        mbar_init(corr_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int c_tile = tile % nct, p_tile = (tile / nct) % kPT, img = tile / (nct * kPT);
                const int row_base = (a.w_row_base ? a.w_row_base[img] / a.w_row_div : 0) + c_tile * kTileCh;
                const int y0 = p_tile * (kTilePix / kW);
                for (int kb = 0; kb < kNumKB; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
                    const int tap = kb / kKBperTap, cb = kb % kKBperTap;
                    const int ky = tap / KS, kx = tap % KS;
                    const uint32_t sa = base + s * kStageBytes;
                    mbar_expect_tx(full_bar(s), kStageBytes);
                    tma_load_2d(sa, &tmW_hi, full_bar(s), kb * kBlockK, row_base);
                    tma_load_2d(sa + kWBytes, &tmW_lo, full_bar(s), kb * kBlockK, row_base);
                    tma_load_4d(sa + 2 * kWBytes, &tmA_hi, full_bar(s), cb * kBlockK, kx - kPad, y0 + ky - kPad, img);
                    tma_load_4d(sa + 2 * kWBytes + kActBytes, &tmA_lo, full_bar(s), cb * kBlockK, kx - kPad,
                                y0 + ky - kPad, img);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t d_main = tmem_base, d_corr = tmem_base + kCorrCol;
            uint32_t it = 0, gch = 0, ti = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
                for (int kb = 0; kb < kNumKB; ++kb, ++it) {
                    const int s = it % kStages;
                    const int kc = kb % kChunkKB;
                    mbar_wait(full_bar(s), (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = base + s * kStageBytes;
                    const uint64_t w_hi = umma_desc_sw128(sa), w_lo = umma_desc_sw128(sa + kWBytes);
                    const uint64_t x_hi = umma_desc_sw128(sa + 2 * kWBytes), x_lo = umma_desc_sw128(sa + 2 * kWBytes + kActBytes);
                    if (kb == 0) {
                        // tile boundary: accumulator 0 is released by the last drain of the previous tile,
                        // accumulator 1 a little later (it is folded into the totals after that drain)
                        if (gch > 0) {
                            mbar_wait(main_empty_bar, (gch - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_main, w_hi + adv, x_hi + adv, k > 0 ? 1u : 0u);
                        }
                        if (ti > 0) {
                            mbar_wait(corr_empty_bar, (ti - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_corr, w_hi + adv, x_lo + adv, k > 0 ? 1u : 0u);
                            umma_f16(d_corr, w_lo + adv, x_hi + adv, 1u);
                        }
                    } else if (kb == kNumKB - 1) {
                        // last k-block of the tile: finish accumulator 0 first and publish it, so that its
                        // drain overlaps the last correction MMAs
                        if (kc == 0) {
                            mbar_wait(main_empty_bar, (gch - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_main, w_hi + adv, x_hi + adv, (kc > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(main_full_bar);
                        ++gch;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_corr, w_hi + adv, x_lo + adv, 1u);
                            umma_f16(d_corr, w_lo + adv, x_hi + adv, 1u);
                        }
                        umma_commit(corr_full_bar);
                    } else {
                        // correction products first: they do not touch accumulator 0, which the drain
                        // warps may still be reading at a chunk boundary
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_corr, w_hi + adv, x_lo + adv, 1u);
                            umma_f16(d_corr, w_lo + adv, x_hi + adv, 1u);
                        }
                        if (kc == 0) {
                            mbar_wait(main_empty_bar, (gch - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_main, w_hi + adv, x_hi + adv, (kc > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(s));            // frees this smem stage once the MMAs have read it
                    if (kc == kChunkKB - 1 && kb != kNumKB - 1) { umma_commit(main_full_bar); ++gch; }   // this tap's partial sum is complete
                }
            }
        }
    } else {
        // ===== drain + epilogue warps: TMEM -> registers (running totals) -> global =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // pixel half of the tile: warps 2-5 -> 0, warps 6-9 -> 1
        const int r = q * 32 + lane;                  // output channel (row of the weight tile)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + half * kHalfPix;
        uint32_t gch = 0, ti = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
            const int c_tile = tile % nct, p_tile = (tile / nct) % kPT, img = tile / (nct * kPT);
            float tot[kHalfPix];
            for (int chunk = 0; chunk < kNumChunks; ++chunk, ++gch) {
                mbar_wait(main_full_bar, gch & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (chunk == 0) {
                    // first tap: the totals ARE the accumulator -- all loads in flight, one wait
#pragma unroll
                    for (int c = 0; c < kHalfPix; c += 8) {
                        uint32_t v[8];
                        tmem_ld8(lane_addr + c, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) tot[c + j] = __uint_as_float(v[j]);     // (bias fix applied below)
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < kHalfPix; ++c) tot[c] = fmaf(tot[c], a.trunc_fix, tot[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < kHalfPix; c += 24) {
                        uint32_t v0[8], v1[8], v2[8];
                        tmem_ld8(lane_addr + c, v0);
                        tmem_ld8(lane_addr + c + 8, v1);
                        tmem_ld8(lane_addr + c + 16, v2);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float x0 = __uint_as_float(v0[j]), x1 = __uint_as_float(v1[j]), x2 = __uint_as_float(v2[j]);
                            tot[c + j] += fmaf(x0, a.trunc_fix, x0);
                            tot[c + 8 + j] += fmaf(x1, a.trunc_fix, x1);
                            tot[c + 16 + j] += fmaf(x2, a.trunc_fix, x2);
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(main_empty_bar) : "memory");
            }
            // accumulator 1 (all correction products of the tile): fold into the totals, hand TMEM back
            mbar_wait(corr_full_bar, ti & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < kHalfPix; c += 40) {
                uint32_t v0[8], v1[8], v2[8], v3[8], v4[8];
                tmem_ld8(lane_addr + kCorrCol + c, v0);
                tmem_ld8(lane_addr + kCorrCol + c + 8, v1);
                tmem_ld8(lane_addr + kCorrCol + c + 16, v2);
                tmem_ld8(lane_addr + kCorrCol + c + 24, v3);
                tmem_ld8(lane_addr + kCorrCol + c + 32, v4);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    tot[c + j] += __uint_as_float(v0[j]) * (1.0f / kLoScale);
                    tot[c + 8 + j] += __uint_as_float(v1[j]) * (1.0f / kLoScale);
                    tot[c + 16 + j] += __uint_as_float(v2[j]) * (1.0f / kLoScale);
                    tot[c + 24 + j] += __uint_as_float(v3[j]) * (1.0f / kLoScale);
                    tot[c + 32 + j] += __uint_as_float(v4[j]) * (1.0f / kLoScale);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(corr_empty_bar) : "memory");

            // epilogue: each pixel half (4 warps = 128 channels) stages 30 pixels x 128 channels in shared
            // memory and one thread hands the box to the TMA store engine; the 4-byte-per-lane global
            // stores this replaces kept the drain warps busy for ~6k cycles per tile, longer than a
            // K = 512 (ks = 1) tile can hide.
            const int row_base = (a.w_row_base ? a.w_row_base[img] / a.w_row_div : 0) + c_tile * kTileCh;
            const float bias = a.bias ? a.bias[row_base + r] : 0.0f;
            const uint32_t stage_out = out_smem + half * kOutBytes;
            const int64_t row0 = (int64_t)img * rows_per_img + p_tile * kTilePix + half * kHalfPix;
            if (a.nchw) {
                // channel-major output [img][col][pixel] (what the decoder takes as visual_feature): this thread's
                // 120 pixels of its channel are contiguous
                float *dst = a.out + ((int64_t)img * a.cols + c_tile * kTileCh + r) * rows_per_img + p_tile * kTilePix + half * kHalfPix;
#pragma unroll
                for (int j = 0; j < kHalfPix; j += 4) {
                    float4 v;
                    v.x = tot[j] * a.inv_scale + bias; v.y = tot[j + 1] * a.inv_scale + bias;
                    v.z = tot[j + 2] * a.inv_scale + bias; v.w = tot[j + 3] * a.inv_scale + bias;
                    if (a.relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
                    *reinterpret_cast<float4 *>(dst + j) = v;
                }
                continue;
            }
#pragma unroll
            for (int rr = 0; rr < kHalfPix / kOutRows; ++rr) {
                if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer free again
                named_bar_sync(1 + half, 128);
#pragma unroll
                for (int j = 0; j < kOutRows; ++j) {
                    float v = tot[rr * kOutRows + j] * a.inv_scale + bias;
                    if (a.relu) v = fmaxf(v, 0.0f);
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage_out + (uint32_t)(j * kTileCh + r) * 4), "f"(v) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1 + half, 128);
                if (r == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmOut), "r"(stage_out), "r"(c_tile * kTileCh), "r"((int)(row0 + rr * kOutRows))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (r == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all stores landed before exit
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------
// Winograd F(2x4, 3x3) gate GEMMs with the ROW half of the output transform in the epilogue.
//   m[i][j] = U[4j+i] . W[4j+i]^T     (K = 512; i = 0..3 the F(2,3) row position, j = 0..5 the F(4,3) column position)
//   out[2j + 0] = m[0][j] + m[1][j] + m[2][j]
//   out[2j + 1] = m[1][j] - m[2][j] - m[3][j]          (= A_2^T m; the cell kernel finishes with . A_4)
// 24 multiplies per 8 outputs (the direct convolution: 72, F(2x2): 32), and only 12 result planes leave the
// SM (and are read back by the cell kernel) instead of 24.
//   * tile = 128 gate columns (UMMA M) x 128 Winograd tiles (UMMA N) x the 4 row positions of one j; the K
//     loops of the 4 positions run back to back.  Operands are (hi, lo) fp16 pairs, x = hi + lo / 2^11, and
//     every position has TWO 128-column TMEM accumulators side by side: main = hi*hi, corr = hi*lo + lo*hi.
//     Per k-step ONE N = 256 MMA  W_hi x [X_hi ; X_lo] -> [main | corr]  (the two activation operand tiles are
//     adjacent in shared memory) and one N = 128 MMA  W_lo x X_hi -> corr.  The truncating adds of the tensor
//     core then act on 32 (main) and 64 (corr, weight 2^-11) accumulation steps instead of 96;
//   * two position buffers (2 x 256 = all 512 TMEM columns): the tensor core works on position p+1 while the
//     drain warps read position p (thread = 1 gate column x 64 tiles), fold main + corr / 2^11 into two running
//     sums in registers (fp32, round-to-nearest) and release the buffer; after i = 3 the two sums are stored
//     (a warp store = 32 adjacent gate columns of one tile row, evict-first: the results stream through L2 once);
//   * 64 KB stages {W_hi, W_lo, X_hi, X_lo} x 128 rows x 64 k, 128-byte swizzle, 3-deep mbarrier ring.
// Measured (B200, 256 images): the MMA thread spends ~85 % of the kernel blocked on MMA issue and the chip runs
// at its power cap (~1.4 GHz): the issued rate is 0.85-0.9 of what cuBLAS bf16 sustains on the same box.  A CTA-pair
// (cta_group::2) variant that halves the shared-memory operand traffic was built, verified and timed: same
// duration -- the bound is the number of issued MMA flops, which is why this is F(2x4) and not F(2x2).
// Result layout (tile-major): out[((2j + r) * cols/128 + col/128) * rows_pad + row][col % 128].
// ---------------------------------------------------------------------------
namespace wg {
constexpr int kTileRows = 128;                             // Winograd tiles per GEMM tile (UMMA N)
constexpr int kPosI = 4, kPosJ = 6;                        // F(2,3) positions folded in the epilogue, F(4,3) positions
constexpr int kBK = 64;                                    // k-block: 64 fp16 = one 128-byte swizzle row
constexpr int kOpBytes = 128 * kBK * 2;                    // 16 KB per operand tile
constexpr int kStageBytes = 4 * kOpBytes;                  // 64 KB: W_hi, W_lo, X_hi, X_lo
constexpr int kStages = 3;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr int kAccCols = 128, kNumBuf = 2;                 // per buffer: main | corr
constexpr uint32_t kIdescN256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kTileCh >> 4) << 24);
constexpr uint32_t kIdescN128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(kTileCh >> 4) << 24);
static_assert(kSmemBytes <= 232448, "shared memory budget");
}  // namespace wg

__device__ __forceinline__ void umma_f16_idesc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// CHUNKS: accumulator chunks per position.  1 = the whole K = 512 of a position in one [main | corr] buffer (32
// accumulation steps; the h-gates).  4 = a fresh buffer every 2 k-blocks (8 accumulation steps), folded into the
// running sums by the drain warps with round-to-nearest adds: less truncation noise for the output transform to
// amplify -- for the loop-invariant x-gates, whose error is coherent over the 16 steps (decode.cu).
template <int CHUNKS>
__global__ void __launch_bounds__(kThreads, 1)
wino_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmU_hi, const __grid_constant__ CUtensorMap tmU_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                    float *__restrict__ out, int num_tiles, int nct, int ntb, int cols, int64_t rows_pad,
                    float inv_scale, float trunc_fix) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + wg::kStages * wg::kStageBytes;
    auto full_bar = [&](int s) { return bar0 + 8 * s; };
    auto empty_bar = [&](int s) { return bar0 + 8 * (wg::kStages + s); };
    auto acc_full_bar = [&](int b) { return bar0 + 8 * (2 * wg::kStages + b); };                   // MMA -> drain warps
    auto acc_empty_bar = [&](int b) { return bar0 + 8 * (2 * wg::kStages + wg::kNumBuf + b); };    // drain warps -> MMA
    const uint32_t tmem_slot = bar0 + 8 * (2 * wg::kStages + 2 * wg::kNumBuf);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kNumKB = kE / wg::kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmU_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmU_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
        for (int s = 0; s < wg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < wg::kNumBuf; ++b) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    // tile -> (column tile fastest, then row block, then j): the 16 column tiles of one row block run at the
    // same time (its U rows come from HBM once), the weights of one j (16 MB) stay in L2
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int c_tile = tile % nct, tb = (tile / nct) % ntb, j = tile / (nct * ntb);
                for (int i = 0; i < wg::kPosI; ++i) {
                    const int pos = j * wg::kPosI + i;
                    const int w_row = pos * cols + c_tile * kTileCh;
                    for (int kb = 0; kb < kNumKB; ++kb, ++it) {
                        const int s = it % wg::kStages;
                        mbar_wait(empty_bar(s), ((it / wg::kStages) & 1) ^ 1);
                        const uint32_t sa = base + s * wg::kStageBytes;
                        mbar_expect_tx(full_bar(s), wg::kStageBytes);
                        tma_load_2d(sa, &tmW_hi, full_bar(s), kb * wg::kBK, w_row);
                        tma_load_2d(sa + wg::kOpBytes, &tmW_lo, full_bar(s), kb * wg::kBK, w_row);
                        tma_load_3d(sa + 2 * wg::kOpBytes, &tmU_hi, full_bar(s), kb * wg::kBK, tb * wg::kTileRows, pos);
                        tma_load_3d(sa + 3 * wg::kOpBytes, &tmU_lo, full_bar(s), kb * wg::kBK, tb * wg::kTileRows, pos);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, pc = 0;                       // pc: positions issued so far (buffer ring index)
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int ic = 0; ic < wg::kPosI * CHUNKS; ++ic, ++pc) {       // (position i, chunk) pairs
                    const int buf = pc & (wg::kNumBuf - 1);
                    const uint32_t d_main = tmem_base + buf * 2 * wg::kAccCols, d_corr = d_main + wg::kAccCols;
                    mbar_wait(acc_empty_bar(buf), ((pc / wg::kNumBuf) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < kNumKB / CHUNKS; ++kb, ++it) {
                        const int s = it % wg::kStages;
                        mbar_wait(full_bar(s), (it / wg::kStages) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = base + s * wg::kStageBytes;
                        const uint64_t w_hi = umma_desc_kmajor<128>(sa), w_lo = umma_desc_kmajor<128>(sa + wg::kOpBytes);
                        const uint64_t x_hilo = umma_desc_kmajor<128>(sa + 2 * wg::kOpBytes);   // 256 rows: X_hi then X_lo
#pragma unroll
                        for (int k = 0; k < wg::kBK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                            // [main | corr] (+)= W_hi . [X_hi ; X_lo]^T, then corr += W_lo . X_hi^T
                            umma_f16_idesc(d_main, w_hi + adv, x_hilo + adv, wg::kIdescN256, acc);
                            umma_f16_idesc(d_corr, w_lo + adv, x_hilo + adv, wg::kIdescN128, 1u);
                        }
                        umma_commit(empty_bar(s));
                    }
                    umma_commit(acc_full_bar(buf));
                }
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;    // TMEM lane quarter (gate columns), half of the 128 tiles
        constexpr int kMine = wg::kTileRows / 2;           // 64 tiles per thread
        uint64_t evict_first;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(evict_first));
        uint32_t pc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int c_tile = tile % nct, tb = (tile / nct) % ntb, j = tile / (nct * ntb);
            float sa[kMine], sb[kMine];
#pragma unroll
            for (int ic = 0; ic < wg::kPosI * CHUNKS; ++ic, ++pc) {
                const int i = ic / CHUNKS;
                const bool first = (ic % CHUNKS) == 0;        // first chunk of its position
                const int buf = pc & (wg::kNumBuf - 1);
                const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * wg::kAccCols + half * kMine;
                mbar_wait(acc_full_bar(buf), (pc / wg::kNumBuf) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < kMine; c += 16) {
                    uint32_t vm[2][8], vc[2][8];
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        tmem_ld8(lane_addr + c + 8 * b, vm[b]);
                        tmem_ld8(lane_addr + wg::kAccCols + c + 8 * b, vc[b]);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float mm = __uint_as_float(vm[e >> 3][e & 7]);
                        const float m = fmaf(__uint_as_float(vc[e >> 3][e & 7]), 1.0f / kLoScale, fmaf(mm, trunc_fix, mm));
                        if (i == 0) sa[c + e] = first ? m : sa[c + e] + m;
                        else if (i == 1) { sa[c + e] += m; sb[c + e] = first ? m : sb[c + e] + m; }
                        else if (i == 2) { sa[c + e] += m; sb[c + e] -= m; }
                        else sb[c + e] -= m;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty_bar(buf)) : "memory");
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int64_t row0 = ((int64_t)(2 * j + r) * nct + c_tile) * rows_pad + (int64_t)tb * wg::kTileRows + half * kMine;
                float *dst = out + row0 * kTileCh + q * 32 + lane;
#pragma unroll
                for (int e = 0; e < kMine; ++e) {
                    const float v = (r == 0 ? sa[e] : sb[e]) * inv_scale;
                    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(dst + (int64_t)e * kTileCh), "f"(v), "l"(evict_first)
                                 : "memory");
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

static int make_map_a(CUtensorMap *m, const __half *ptr, int n_images, int rows_per_img, int cin) {
    const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)kW, (cuuint64_t)(rows_per_img / kW), (cuuint64_t)n_images};
    const cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)kW * cin * 2, (cuuint64_t)rows_per_img * cin * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)kW, (cuuint32_t)(kTilePix / kW), 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)ptr, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, kSwizzleBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_map_b(CUtensorMap *m, const __half *ptr, int64_t rows, int64_t K, int block_k = kBlockK) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)kTileCh};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_map_out(CUtensorMap *m, float *ptr, int64_t ldo, int64_t rows, int box_cols) {
    const cuuint64_t dims[2] = {(cuuint64_t)ldo, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ldo * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)kOutRows};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace tc

int conv_gemm_tc(const ConvGemmArgs &a_in, cudaStream_t s) {
    using namespace tc;
    ConvGemmArgs a = a_in;
    a.trunc_fix = acc_trunc_fix();
    // ks = 1 / 3 / 5: convolution over 30 x 40 images, operand pairs x = hi + lo / 2^11; with ks = 1 an "image" may
    // be any multiple of 240 rows (plain batched GEMM: out[b][row][col] = sum_k a[b][row][k] w[base_b + col][k])
    const int rows = a.rows_per_img;
    if (a.cols % kTileCh != 0 || (a.ks != 1 && a.ks != 3 && a.ks != 5) || a.ldo % 4 != 0 || ((uintptr_t)a.out & 15) != 0 ||
        a.w_row_div < 1 || rows <= 0 || rows % kTilePix != 0 || (a.ks != 1 && rows != kHW) ||
        (a.cin != kE && !(a.cin == 2048 && a.ks == 3)) || (a.nchw && a.w_row_base != nullptr)) {
        set_error("conv_gemm_tc: cols must be a multiple of %d, rows per image of %d, ks 1, 3 or 5, out 16-byte aligned",
                  kTileCh, kTilePix);
        return SPB_ERR_ARG;
    }
    if (get_encode() == nullptr) {
        set_error("conv_gemm_tc: cuTensorMapEncodeTiled not available from the driver");
        return SPB_ERR_CUDA;
    }
    const int64_t K = (int64_t)a.ks * a.ks * a.cin;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo, mo;
    int rc = make_map_a(&ma_hi, a.a_hi, a.n_images, rows, a.cin);
    if (!rc) rc = make_map_a(&ma_lo, a.a_lo, a.n_images, rows, a.cin);
    if (!rc) rc = make_map_b(&mw_hi, a.w_hi, a.w_rows, K);
    if (!rc) rc = make_map_b(&mw_lo, a.w_lo, a.w_rows, K);
    if (!rc) rc = make_map_out(&mo, a.out, a.ldo, (int64_t)a.n_images * rows, kTileCh);
    if (rc) {
        set_error("conv_gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d", rc);
        return SPB_ERR_CUDA;
    }
    const int nct = a.cols / kTileCh, pt = rows / kTilePix;
    const int64_t tiles64 = (int64_t)nct * pt * a.n_images;
    if (tiles64 > 0x7fffffff) {
        set_error("conv_gemm_tc: too many tiles");
        return SPB_ERR_ARG;
    }
    const int num_tiles = (int)tiles64;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();       // persistent: one CTA per SM
#define SPB_LAUNCH_TC(KS_, CIN_)                                                                                          \
    SPB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<KS_, CIN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); \
    conv_gemm_tc_kernel<KS_, CIN_><<<grid, kThreads, kSmemBytes, s>>>(ma_hi, ma_lo, mw_hi, mw_lo, mo, a, num_tiles, nct, pt, rows)
    if (a.cin == 2048) { SPB_LAUNCH_TC(3, 2048); }
    else if (a.ks == 1) { SPB_LAUNCH_TC(1, 512); }
    else if (a.ks == 3) { SPB_LAUNCH_TC(3, 512); }
    else { SPB_LAUNCH_TC(5, 512); }
#undef SPB_LAUNCH_TC
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

// Winograd F(2x4,3x3) gate GEMMs + row output transform (wino_gemm_tc_kernel).
//   u  [24][rows_pad][512] fp16 pairs (hi + lo/2^11), position p = 4j + i;  w [24 * cols][512] fp16 pairs;
//   out [12][cols/128][rows_pad][128] fp32.  rows_pad % 128 == 0, cols % 128 == 0.
//   pos_j = 4: the same kernel on the 16 positions of F(2x2,3x3) (u [16][..], w [16 * cols][..], out [8][..]) --
//   the row positions i and their fold in the epilogue are those of F(2,3) either way.
int wino_gemm_tc(const __half *u_hi, const __half *u_lo, const __half *w_hi, const __half *w_lo, float *out,
                 int64_t rows_pad, int cols, float inv_scale, cudaStream_t s, bool fine_drain, int pos_j) {
    using namespace tc;
    if (pos_j != wg::kPosJ && pos_j != 4) {
        set_error("wino_gemm_tc: 6 (F(2x4)) or 4 (F(2x2)) column positions");
        return SPB_ERR_ARG;
    }
    const int kPos = wg::kPosI * pos_j;
    if (rows_pad <= 0 || rows_pad % wg::kTileRows != 0 || cols <= 0 || cols % kTileCh != 0 || ((uintptr_t)out & 15) != 0) {
        set_error("wino_gemm_tc: rows_pad must be a multiple of %d, cols of %d", wg::kTileRows, kTileCh);
        return SPB_ERR_ARG;
    }
    if (get_encode() == nullptr) {
        set_error("wino_gemm_tc: cuTensorMapEncodeTiled not available from the driver");
        return SPB_ERR_CUDA;
    }
    const int nct = cols / kTileCh;
    const int64_t ntb = rows_pad / wg::kTileRows;
    if ((int64_t)kPos * rows_pad > 0x7fffffff || pos_j * nct * ntb > 0x7fffffff) {
        set_error("wino_gemm_tc: too many rows");
        return SPB_ERR_ARG;
    }
    CUtensorMap mu_hi, mu_lo, mw_hi, mw_lo;
    auto make_u = [&](CUtensorMap *m, const __half *ptr) -> int {
        const cuuint64_t dims[3] = {(cuuint64_t)kE, (cuuint64_t)rows_pad, (cuuint64_t)kPos};
        const cuuint64_t strides[2] = {(cuuint64_t)kE * 2, (cuuint64_t)rows_pad * kE * 2};
        const cuuint32_t box[3] = {(cuuint32_t)wg::kBK, (cuuint32_t)wg::kTileRows, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void *)ptr, dims, strides, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS ? 0 : (int)r;
    };
    int rc = make_u(&mu_hi, u_hi);
    if (!rc) rc = make_u(&mu_lo, u_lo);
    if (!rc) rc = make_map_b(&mw_hi, w_hi, (int64_t)kPos * cols, kE, wg::kBK);
    if (!rc) rc = make_map_b(&mw_lo, w_lo, (int64_t)kPos * cols, kE, wg::kBK);
    if (rc) {
        set_error("wino_gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d", rc);
        return SPB_ERR_CUDA;
    }
    const int num_tiles = (int)(pos_j * nct * ntb);
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    if (fine_drain) {
        SPB_CUDA(cudaFuncSetAttribute(wino_gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes));
        wino_gemm_tc_kernel<4><<<grid, kThreads, wg::kSmemBytes, s>>>(mu_hi, mu_lo, mw_hi, mw_lo, out, num_tiles, nct, (int)ntb,
                                                                     cols, rows_pad, inv_scale, acc_trunc_fix_fine());
    } else {
        SPB_CUDA(cudaFuncSetAttribute(wino_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes));
        wino_gemm_tc_kernel<1><<<grid, kThreads, wg::kSmemBytes, s>>>(mu_hi, mu_lo, mw_hi, mw_lo, out, num_tiles, nct, (int)ntb,
                                                                     cols, rows_pad, inv_scale, acc_trunc_fix());
    }
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

}  // namespace spb
