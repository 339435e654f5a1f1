// Internal declarations shared by the decoder sources (decode.cu, conv_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace spb {

constexpr int kE = 512;          // embedding channels
constexpr int kH = 30, kW = 40, kHW = 1200;
constexpr int kGateCols = 4 * kE;   // i, f, o, g
constexpr float kLoScale = 2048.0f; // lo half is stored as (x - hi) * 2^11

// One implicit-GEMM convolution:  out[(n*1200+p)*ldo + col] = inv_scale * conv(a, w)[p, col] (+ bias[col])
//   a  = a_hi + a_lo / 2^11   fp16 NHWC [N,30,40,512]
//   w  = w_hi + w_lo / 2^11   fp16 [rows, ks*ks*512], K index = (ky*ks+kx)*512 + ci, pre-multiplied by 1/inv_scale
//   row of w used for output column `col` of image n:  w_row_base[n] + col   (w_row_base NULL -> 0)
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
    // mode 1 (tensor-core path only): the ConvLSTM cell is fused into the epilogue of the 3x3 gate
    // convolution -- gates = out + xg + rank-1 memory term; c' = f c + i g; h' = o c' is written
    // as the fp16 pair the next convolutions consume.  `out` is unused in this mode.
    int mode = 0;
    const float *xg = nullptr;        // [N,1200,2048] x-convolution + biases, same column order as the weights
    float *c = nullptr;               // [N,1200,512] cell state, updated in place
    const float *V = nullptr;         // [N,S,3,512,9] rank-1 projections
    const float *sp_mem = nullptr;    // [N,S,1200] spatial memory
    int n_streams = 0;
    __half *h_out_hi = nullptr, *h_out_lo = nullptr;   // [N,1200,512] next hidden state (NOT the buffer being read)
};

// Column order of the 2048 gate columns: [64-channel block cb][32-channel half][gate i,f,o,g][32].
// One epilogue thread of the tensor-core kernel owns one pixel x one (cb, half): all four gates of
// 32 channels sit in its 128 accumulator columns.
__host__ __device__ inline int gate_col(int ch, int g) {
    return (ch >> 6) * 256 + ((ch >> 5) & 1) * 128 + g * 32 + (ch & 31);
}

int conv_gemm_simt(const ConvGemmArgs &a, cudaStream_t s);
int conv_gemm_tc(const ConvGemmArgs &a, cudaStream_t s);     // tcgen05 / TMEM / TMA (conv_tc.cu)

}  // namespace spb
