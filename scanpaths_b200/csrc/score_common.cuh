// Shared by the two scoring kernels (score_pairs.cu: one warp per pair, any length; score_pairs_g8.cu: four
// pairs per warp, the fast path for the drivers' sizes).
#pragma once
#include "common.cuh"

namespace spb {

struct PairLayout {                             // per-warp shared memory slice (byte offsets)
    int ax, ay, bx, by, D, W, arun, brun, ased, bsed, ar, ac, br, bc, awr, awc, bwr, bwc, bytes, pitch;
};

__host__ __device__ inline PairLayout make_layout(int LA, int LB) {
    PairLayout L;
    int o = 0;
    L.pitch = LA | 1;
    L.ax = o; o += 8 * LA;
    L.ay = o; o += 8 * LA;
    L.bx = o; o += 8 * LB;
    L.by = o; o += 8 * LB;
    L.D = o; o += 8 * LB * L.pitch;
    L.W = o; o += 8 * LB * L.pitch;
    L.arun = o; o += 4 * LA;
    L.brun = o; o += 4 * LB;
    L.ased = o; o += 4 * LA;
    L.bsed = o; o += 4 * LB;
    L.ar = o; o += LA;
    L.ac = o; o += LA;
    L.br = o; o += LB;
    L.bc = o; o += LB;
    L.awr = o; o += LA;
    L.awc = o; o += LA;
    L.bwr = o; o += LB;
    L.bwc = o; o += LB;
    L.bytes = (o + 15) & ~15;
    return L;
}

// Fast path (score_pairs_g8.cu).  Returns SPB_OK and sets *handled = 1 if the launch was taken: GapValue = 0
// and per-pair shared-memory slices small enough for >= 16 pairs per block; otherwise *handled = 0 and the
// caller runs the warp-per-pair kernel.
int score_pairs_g8(const spb_path_pack &A, const spb_path_pack &B, const int32_t *pair_h, const int32_t *pair_s,
                   int64_t n_pairs, const spb_score_cfg &cfg, double *scores, void *workspace, int64_t workspace_bytes,
                   int32_t *err, cudaStream_t stream, int *handled);

}  // namespace spb
