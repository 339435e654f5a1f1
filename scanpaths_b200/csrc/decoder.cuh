// Internal declarations shared by the decoder sources (decode.cu, conv_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace spb {

constexpr int kE = 512;          // embedding channels
constexpr int kH = 30, kW = 40, kHW = 1200;
constexpr int kGateCols = 4 * kE;   // i, f, o, g
constexpr float kLoScale = 2048.0f; // lo half is stored as (x - hi) * 2^11
// The tensor core adds into its fp32 accumulator with truncation toward zero.  Over the 32 k-steps a
// "main" accumulator lives (one filter tap / one Winograd position) that shrinks the result by a factor
// measured on B200 as 1 - 5.5e-7 for mixed-sign operands (scratch/bias_probe.py: -5.36e-7 .. -5.52e-7 for
// every kernel of conv_tc.cu; up to -2.6e-6 if every product has the same sign -- pinned by
// tests/test_gpu_decoder.py::test_acc_trunc_fix_validity_range).  Unlike rounding noise this bias is coherent
// over all outputs and steps -- it made the decode error grow linearly with the step count -- so the drain
// warps multiply it back: x += x * fix.  The factor is a run-time value: the default below, re-measured on
// the device at decoder construction by a probe GEMM (models/baseline_attention.py::calibrate_acc_trunc_fix)
// so that another SKU / driver with a different accumulation cannot silently shift parity.
constexpr float kAccTruncFixDefault = 5.5e-7f;
float acc_trunc_fix();                 // current value (spb_set_acc_trunc_fix), accumulators of 32 k-steps
float acc_trunc_fix_fine();            // the same for the 8-k-step accumulators of wino_gemm_tc_kernel<4>

// One implicit-GEMM convolution:  out[(n*1200+p)*ldo + col] = inv_scale * conv(a, w)[p, col] (+ bias[col])
//   a  = a_hi + a_lo / 2^11   fp16 NHWC [N,30,40,512]
//   w  = w_hi + w_lo / 2^11   fp16 [rows, ks*ks*512], K index = (ky*ks+kx)*512 + ci, pre-multiplied by 1/inv_scale
//   row of w used for output column `col` of image n:  w_row_base[n] / w_row_div + col   (w_row_base NULL -> 0)
struct ConvGemmArgs {
    const __half *a_hi, *a_lo;
    const __half *w_hi, *w_lo;
    const int32_t *w_row_base;
    int64_t w_rows;               // total rows of w (for the tensor map)
    const float *bias;            // [cols] indexed like w rows (after w_row_base), or NULL
    float *out;
    int64_t ldo;
    int n_images, cols, ks;
    float inv_scale;
    int w_row_div = 1;            // w_row_base is given in units of w_row_div rows
    int rows_per_img = kHW;       // ks = 1 only: any multiple of 240 (the kernel as a plain batched GEMM)
    float trunc_fix = 0.0f;       // set by conv_gemm_tc from acc_trunc_fix()
    int cin = kE;                 // input channels: 512, or 2048 for the encoder's sal_conv (ks = 3, tensor-core route)
    int relu = 0;                 // epilogue: max(., 0)
    int nchw = 0;                 // epilogue: channel-major output out[(n*cols + col)*rows_per_img + p] (ldo unused)
};

// Column order of the 2048 gate columns: [64-channel block cb][32-channel half][gate i,f,o,g][32].
// (a 128-row weight tile of the tensor-core kernel = all four gates of 32 channels.)
__host__ __device__ inline int gate_col(int ch, int g) {
    return (ch >> 6) * 256 + ((ch >> 5) & 1) * 128 + g * 32 + (ch & 31);
}

int conv_gemm_simt(const ConvGemmArgs &a, cudaStream_t s);
int conv_gemm_tc(const ConvGemmArgs &a, cudaStream_t s);     // tcgen05 / TMEM / TMA (conv_tc.cu)
// the 24 Winograd F(2x4,3x3) per-position GEMMs + the row half of the output transform (conv_tc.cu):
// u [24][rows_pad][512], w [24*cols][512] fp16 pairs (position 4j+i) -> out [12][cols/128][rows_pad][128] fp32
// pos_j = 4: the 16 positions of F(2x2,3x3) (u [16][rows_pad][512], w [16*cols][512] -> out [8][...])
int wino_gemm_tc(const __half *u_hi, const __half *u_lo, const __half *w_hi, const __half *w_lo, float *out,
                 int64_t rows_pad, int cols, float inv_scale, cudaStream_t s, bool fine_drain = false, int pos_j = 6);

}  // namespace spb
