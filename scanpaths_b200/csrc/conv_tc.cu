// Implicit-GEMM 3x3 / 5x5 convolution on Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// out[(img*1200 + p)*ldo + col] = inv_scale * sum_{tap,ci} a[img, p+tap, ci] * w[row(col), tap, ci] (+ bias)
//
// This is the dominant kernel of the decode path: the 3x3 gate convolutions of the
// ConvLSTM (ConvLSTM.forward, OSIE/models/baseline_attention.py:39-42; M = 1200 pixels per
// image, N = 2048 = 4 gates x 512, K = 9 x 512) and the 5x5 layer before the head
// (:202, :352; N = 512, K = 25 x 512).
//
// Design
//   * one CTA per (120-pixel, 256-column) output tile: 120 = 3 image rows x 40, so every
//     filter tap of the A operand is ONE 4-D TMA box {64 ch, 40 w, 3 h, 1 img} shifted by the
//     tap offset; out-of-image rows / columns are zero-filled by TMA (the conv padding).
//     The UMMA tile is 128 x 256; its last 8 rows are don't-care.
//   * fp32-equivalent arithmetic on the fp16 pipe: every operand is a pair
//     x = hi + lo / 2^11 (11 + 11 significand bits).  Three MMAs per k-step:
//     hi*hi -> accumulator 0, hi*lo + lo*hi -> accumulator 1 (scaled by 2^11), combined in
//     the epilogue.  The dropped lo*lo term is 2^-22 relative.  TMEM: 2 x 256 fp32 columns.
//   * the tensor core adds into its fp32 accumulator with truncation (measured on B200:
//     -1.6e-8 relative per accumulation step, i.e. -4.5e-6 after K = 4608 and -1.2e-5 after
//     K = 12800 -- outside the 1e-5 parity gate).  So accumulator 0 only ever holds ONE
//     filter tap (32 k-steps): after each tap the 8 epilogue warps drain it from TMEM and add
//     it to running totals in registers (round-to-nearest), while the MMA thread already
//     issues the correction MMAs of the next k-block (accumulator 1 is never drained
//     mid-loop: its values weigh 2^-11).
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread) + TMEM allocator,
//     warps 2-9 drain/epilogue (each thread: 1 TMEM lane x 128 columns of running totals).
//   * persistent: grid = #SMs, every CTA walks tiles blockIdx.x, +gridDim.x, ... (column tiles
//     of one pixel tile are adjacent, so the A tile is shared through L2).  Because the totals
//     live in registers, TMEM is free again as soon as the last tap is drained: the MMA thread
//     starts the next tile while the drain warps run the epilogue of the previous one.
//   * epilogue modes: (0) scale + bias + fp32 store; (1) the whole ConvLSTM cell
//     (ConvLSTM.forward :39-46): gate pre-activations = GEMM + x-convolution + rank-1 memory
//     term, sigmoid / tanh, c' = f c + i g, h' = o c' written as the next step's fp16 pair.
//   * operands staged by TMA with 128-byte swizzle, K-major; kStages-deep mbarrier ring.
#include <cuda.h>

#include "decoder.cuh"

namespace spb {

namespace tc {

constexpr int kBlockM = 128, kValidM = 120, kBlockN = 256, kBlockK = 64, kStages = 2;
constexpr int kABytes = kBlockM * kBlockK * 2;            // 16 KB (15 KB written by TMA)
constexpr int kATxBytes = kValidM * kBlockK * 2;          // 15360
constexpr int kBBytes = kBlockN * kBlockK * 2;            // 32 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;    // 96 KB
constexpr int kEpiFloats = 2 * (3 * 64 * 9 + 5 * 42);   // fused-cell staging: V tile + spatial halo, <= 2 streams
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiFloats * 4;
constexpr int kTmemCols = 512;
constexpr int kThreads = 320;
constexpr int kChunkKB = kE / kBlockK;                   // k-blocks per drained chunk: one filter tap

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cf. cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B between 8-row groups |
// version 1 | layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int KS, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    ConvGemmArgs a, int num_tiles, int ntn) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = base + kStages * kStageBytes;
    auto full_bar = [&](int s) { return bar0 + 8 * s; };
    auto empty_bar = [&](int s) { return bar0 + 8 * (kStages + s); };
    const uint32_t main_full_bar = bar0 + 8 * (2 * kStages);        // MMA -> drain warps: one tap accumulated
    const uint32_t main_empty_bar = bar0 + 8 * (2 * kStages + 1);   // drain warps -> MMA: accumulator 0 read out
    const uint32_t corr_empty_bar = bar0 + 8 * (2 * kStages + 2);   // drain warps -> MMA: accumulator 1 read out
    const uint32_t tmem_slot = bar0 + 8 * (2 * kStages + 3);
    float *epi = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)) + kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kNumKB = KS * KS * (kE / kBlockK);
    constexpr int kNumChunks = kNumKB / kChunkKB;
    constexpr int kPad = KS / 2;
    constexpr int kMT = kHW / kValidM;   // 10 pixel tiles per image

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(main_full_bar, 1);
        mbar_init(main_empty_bar, 8);
        mbar_init(corr_empty_bar, 8);
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
                const int n_tile = tile % ntn, m_tile = (tile / ntn) % kMT, img = tile / (ntn * kMT);
                const int row_base = (a.w_row_base ? a.w_row_base[img] : 0) + n_tile * kBlockN;
                const int y0 = m_tile * 3;
                for (int kb = 0; kb < kNumKB; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
                    const int tap = kb / (kE / kBlockK), cb = kb % (kE / kBlockK);
                    const int ky = tap / KS, kx = tap % KS;
                    const uint32_t sa = base + s * kStageBytes;
                    mbar_expect_tx(full_bar(s), 2 * kATxBytes + 2 * kBBytes);
                    tma_load_4d(sa, &tmA_hi, full_bar(s), cb * kBlockK, kx - kPad, y0 + ky - kPad, img);
                    tma_load_4d(sa + kABytes, &tmA_lo, full_bar(s), cb * kBlockK, kx - kPad, y0 + ky - kPad, img);
                    tma_load_2d(sa + 2 * kABytes, &tmB_hi, full_bar(s), kb * kBlockK, row_base);
                    tma_load_2d(sa + 2 * kABytes + kBBytes, &tmB_lo, full_bar(s), kb * kBlockK, row_base);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t d_main = tmem_base, d_corr = tmem_base + kBlockN;
            uint32_t it = 0, gch = 0, ti = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
                for (int kb = 0; kb < kNumKB; ++kb, ++it) {
                    const int s = it % kStages;
                    const int kc = kb % kChunkKB;
                    mbar_wait(full_bar(s), (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = base + s * kStageBytes;
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + kABytes);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * kABytes), b_lo = umma_desc_sw128(sa + 2 * kABytes + kBBytes);
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
                            umma_f16(d_main, a_hi + adv, b_hi + adv, k > 0 ? 1u : 0u);
                        }
                        if (ti > 0) {
                            mbar_wait(corr_empty_bar, (ti - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_corr, a_hi + adv, b_lo + adv, k > 0 ? 1u : 0u);
                            umma_f16(d_corr, a_lo + adv, b_hi + adv, 1u);
                        }
                    } else {
                        // correction products first: they do not touch accumulator 0, which the drain
                        // warps may still be reading at a chunk boundary
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_corr, a_hi + adv, b_lo + adv, 1u);
                            umma_f16(d_corr, a_lo + adv, b_hi + adv, 1u);
                        }
                        if (kc == 0) {
                            mbar_wait(main_empty_bar, (gch - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);
                            umma_f16(d_main, a_hi + adv, b_hi + adv, (kc > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(s));            // frees this smem stage once the MMAs have read it
                    if (kc == kChunkKB - 1) { umma_commit(main_full_bar); ++gch; }   // this tap's partial sum is complete
                }
            }
        }
    } else {
        // ===== drain + epilogue warps: TMEM -> registers (running totals) -> global =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // column half: warps 2-5 -> 0, warps 6-9 -> 1
        const int r = q * 32 + lane;
        const bool valid = r < kValidM;
        const int etid = threadIdx.x - 64;            // 0..255 among the epilogue threads
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + half * (kBlockN / 2);
        uint32_t gch = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_tile = tile % ntn, m_tile = (tile / ntn) % kMT, img = tile / (ntn * kMT);
            float tot[kBlockN / 2];
#pragma unroll
            for (int j = 0; j < kBlockN / 2; ++j) tot[j] = 0.0f;
            for (int chunk = 0; chunk < kNumChunks; ++chunk, ++gch) {
                mbar_wait(main_full_bar, gch & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < kBlockN / 2; c += 32) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16(lane_addr + c, v0);
                    tmem_ld16(lane_addr + c + 16, v1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        tot[c + j] += __uint_as_float(v0[j]);
                        tot[c + 16 + j] += __uint_as_float(v1[j]);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(main_empty_bar) : "memory");
            }
            // the last main_full commit also covers every correction MMA of this tile: fold accumulator 1
            // into the totals and hand TMEM back, then run the epilogue from registers
#pragma unroll
            for (int c = 0; c < kBlockN / 2; c += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16(lane_addr + kBlockN + c, v0);
                tmem_ld16(lane_addr + kBlockN + c + 16, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    tot[c + j] += __uint_as_float(v0[j]) * (1.0f / kLoScale);
                    tot[c + 16 + j] += __uint_as_float(v1[j]) * (1.0f / kLoScale);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(corr_empty_bar) : "memory");

            const int p = m_tile * kValidM + r;                       // pixel of this thread
            const int64_t pix = (int64_t)img * kHW + p;
            if (MODE == 0) {
                if (valid) {
                    const int row_base = (a.w_row_base ? a.w_row_base[img] : 0) + n_tile * kBlockN + half * (kBlockN / 2);
                    float *dst = a.out + pix * a.ldo + n_tile * kBlockN + half * (kBlockN / 2);
#pragma unroll
                    for (int c = 0; c < kBlockN / 2; c += 4) {
                        float4 o;
                        o.x = tot[c] * a.inv_scale; o.y = tot[c + 1] * a.inv_scale;
                        o.z = tot[c + 2] * a.inv_scale; o.w = tot[c + 3] * a.inv_scale;
                        if (a.bias) {
                            o.x += a.bias[row_base + c]; o.y += a.bias[row_base + c + 1];
                            o.z += a.bias[row_base + c + 2]; o.w += a.bias[row_base + c + 3];
                        }
                        *reinterpret_cast<float4 *>(dst + c) = o;
                    }
                }
            } else {
                // ---- fused ConvLSTM cell.  n_tile is the 64-channel block; this thread owns channels
                // ch0 .. ch0+31 of pixel p; tot[g*32 + j] is gate g (i, f, o, g) of channel ch0 + j.
                const int S = a.n_streams;
                const int ch0 = n_tile * 64 + half * 32;
                const int y0 = m_tile * 3;
                float *Vs = epi;                                   // [S][3][64][9]
                float *Hs = epi + S * 3 * 64 * 9;                  // [S][5][42] spatial halo, zero outside the image
                named_bar_sync(1, 256);                            // previous tile's epilogue is done with the staging area
                for (int i = etid; i < S * 3 * 64 * 9; i += 256) {
                    const int sg = i / (64 * 9), rem = i - sg * (64 * 9);      // sg = s*3 + g
                    Vs[i] = a.V[((int64_t)img * S * 3 + sg) * (kE * 9) + n_tile * 64 * 9 + rem];
                }
                for (int i = etid; i < S * 5 * 42; i += 256) {
                    const int st = i / 210, rem = i - st * 210, hy = rem / 42, hx = rem - hy * 42;
                    const int yy = y0 - 1 + hy, xx = hx - 1;
                    Hs[i] = (yy >= 0 && yy < kH && xx >= 0 && xx < kW) ? a.sp_mem[((int64_t)img * S + st) * kHW + yy * kW + xx]
                                                                     : 0.0f;
                }
                named_bar_sync(1, 256);
                if (valid) {
                    const int ly = r / kW, lx = r - ly * kW;       // position inside the 3 x 40 tile
                    float spn[2][9];
#pragma unroll
                    for (int st = 0; st < 2; ++st)
#pragma unroll
                        for (int t9 = 0; t9 < 9; ++t9)
                            spn[st][t9] = (st < S) ? Hs[st * 210 + (ly + t9 / 3) * 42 + lx + t9 % 3] : 0.0f;
                    const float *xg = a.xg + pix * kGateCols + n_tile * kBlockN + half * (kBlockN / 2);
                    float *cptr = a.c + pix * kE + ch0;
                    __half *hhi = a.h_out_hi + pix * kE + ch0, *hlo = a.h_out_lo + pix * kE + ch0;
#pragma unroll
                    for (int j0 = 0; j0 < 32; j0 += 4) {
                        const float4 xi = *reinterpret_cast<const float4 *>(xg + j0);
                        const float4 xf = *reinterpret_cast<const float4 *>(xg + 32 + j0);
                        const float4 xo = *reinterpret_cast<const float4 *>(xg + 64 + j0);
                        const float4 xm = *reinterpret_cast<const float4 *>(xg + 96 + j0);
                        const float4 cv = *reinterpret_cast<const float4 *>(cptr + j0);
                        const float xiv[4] = {xi.x, xi.y, xi.z, xi.w}, xfv[4] = {xf.x, xf.y, xf.z, xf.w};
                        const float xov[4] = {xo.x, xo.y, xo.z, xo.w}, xmv[4] = {xm.x, xm.y, xm.z, xm.w};
                        const float cold[4] = {cv.x, cv.y, cv.z, cv.w};
                        float cn[4];
                        __half hh[4], hl[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j0 + e;
                            float pi = xiv[e] + tot[j] * a.inv_scale;
                            float pf = xfv[e] + tot[32 + j] * a.inv_scale;
                            float po = xov[e] + tot[64 + j] * a.inv_scale;
                            const float pm = xmv[e] + tot[96 + j] * a.inv_scale;
                            for (int st = 0; st < S; ++st) {
                                const float *v = Vs + (st * 3 * 64 + half * 32 + j) * 9;
                                float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
#pragma unroll
                                for (int t9 = 0; t9 < 9; ++t9) {
                                    const float sv = spn[st][t9];
                                    r0 = fmaf(v[t9], sv, r0);
                                    r1 = fmaf(v[64 * 9 + t9], sv, r1);
                                    r2 = fmaf(v[2 * 64 * 9 + t9], sv, r2);
                                }
                                pi += r0; pf += r1; po += r2;
                            }
                            const float gi = sigmoidf_acc(pi), gf = sigmoidf_acc(pf), go = sigmoidf_acc(po);
                            const float gg = tanhf(pm);
                            cn[e] = gf * cold[e] + gi * gg;
                            const float hv = go * cn[e];
                            hh[e] = __float2half_rn(hv);
                            hl[e] = __float2half_rn((hv - __half2float(hh[e])) * kLoScale);
                        }
                        *reinterpret_cast<float4 *>(cptr + j0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                        *reinterpret_cast<uint2 *>(hhi + j0) =
                            make_uint2((uint32_t)__half_as_ushort(hh[0]) | ((uint32_t)__half_as_ushort(hh[1]) << 16),
                                       (uint32_t)__half_as_ushort(hh[2]) | ((uint32_t)__half_as_ushort(hh[3]) << 16));
                        *reinterpret_cast<uint2 *>(hlo + j0) =
                            make_uint2((uint32_t)__half_as_ushort(hl[0]) | ((uint32_t)__half_as_ushort(hl[1]) << 16),
                                       (uint32_t)__half_as_ushort(hl[2]) | ((uint32_t)__half_as_ushort(hl[3]) << 16));
                    }
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

static int make_map_a(CUtensorMap *m, const __half *ptr, int n_images) {
    const cuuint64_t dims[4] = {(cuuint64_t)kE, (cuuint64_t)kW, (cuuint64_t)kH, (cuuint64_t)n_images};
    const cuuint64_t strides[3] = {(cuuint64_t)kE * 2, (cuuint64_t)kW * kE * 2, (cuuint64_t)kHW * kE * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)kW, 3, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)ptr, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_map_b(CUtensorMap *m, const __half *ptr, int64_t rows, int64_t K) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockN};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace tc

int conv_gemm_tc(const ConvGemmArgs &a, cudaStream_t s) {
    using namespace tc;
    if (a.cols % kBlockN != 0 || (a.ldo % 4) != 0 || (a.ks != 3 && a.ks != 5)) {
        set_error("conv_gemm_tc: cols must be a multiple of %d, ldo of 4, ks 3 or 5", kBlockN);
        return SPB_ERR_ARG;
    }
    if (get_encode() == nullptr) {
        set_error("conv_gemm_tc: cuTensorMapEncodeTiled not available from the driver");
        return SPB_ERR_CUDA;
    }
    const int64_t K = (int64_t)a.ks * a.ks * kE;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc = make_map_a(&ma_hi, a.a_hi, a.n_images);
    if (!rc) rc = make_map_a(&ma_lo, a.a_lo, a.n_images);
    if (!rc) rc = make_map_b(&mb_hi, a.w_hi, a.w_rows, K);
    if (!rc) rc = make_map_b(&mb_lo, a.w_lo, a.w_rows, K);
    if (rc) {
        set_error("conv_gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d", rc);
        return SPB_ERR_CUDA;
    }
    const int ntn = a.cols / kBlockN;
    const int num_tiles = ntn * (kHW / kValidM) * a.n_images;
    const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;       // persistent: one CTA per SM
    if (a.mode == 1) {
        if (a.ks != 3 || a.cols != kGateCols || !a.xg || !a.c || !a.V || !a.sp_mem || !a.h_out_hi || !a.h_out_lo ||
            a.n_streams < 1 || a.n_streams > 2 || a.h_out_hi == a.a_hi) {
            set_error("conv_gemm_tc: bad arguments for the fused ConvLSTM-cell epilogue");
            return SPB_ERR_ARG;
        }
        SPB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        conv_gemm_tc_kernel<3, 1><<<grid, kThreads, kSmemBytes, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, a, num_tiles, ntn);
    } else if (a.ks == 3) {
        SPB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        conv_gemm_tc_kernel<3, 0><<<grid, kThreads, kSmemBytes, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, a, num_tiles, ntn);
    } else {
        SPB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        conv_gemm_tc_kernel<5, 0><<<grid, kThreads, kSmemBytes, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, a, num_tiles, ntn);
    }
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

}  // namespace spb
