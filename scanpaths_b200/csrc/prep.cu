// K1 scanpath_prep: fixations -> symbol pack, one thread per fixation.
//
// Replaces, hoisted out of the per-pair loop (the reference recomputes them for
// every pair, OSIE/utils/evaluation.py:184-197):
//   ScanMatch.fixationToSequence   utils/evaltools/scanmatch.py:116-133
//   _scanpath_to_string            utils/evaltools/visual_attention_metrics.py:288-298
//   STDE coordinate rescaling      utils/evaltools/visual_attention_metrics.py:405-415
// HBM-bound, 24 B read + 21 B written per fixation; all arithmetic in f64/int to
// reproduce the reference's truncation and half-to-even rounding exactly.
#include "common.cuh"

namespace spb {

__device__ __forceinline__ int floordiv_i32(int a, int b) {
    int q = a / b;
    return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__global__ void __launch_bounds__(256)
prep_paths_kernel(const double *__restrict__ xyd, const int32_t *__restrict__ len, int64_t n_paths, int lmax,
                  spb_score_cfg cfg, uint8_t *__restrict__ sym, int32_t *__restrict__ run,
                  int32_t *__restrict__ sed, double *__restrict__ xyn) {
    const int64_t total = n_paths * (int64_t)lmax;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx / lmax;
        const int f = (int)(idx - p * lmax);
        if (f >= len[p]) {
            sym[idx] = 0; run[idx] = 0; sed[idx] = 0;
            xyn[2 * idx] = 0.0; xyn[2 * idx + 1] = 0.0;
            continue;
        }
        const double x0 = xyd[3 * idx], y0 = xyd[3 * idx + 1];
        const double t0 = xyd[3 * idx + 2] * cfg.dur_scale;             // evaluation.py:182 (s -> ms)
        // --- ScanMatch bin + temporal repeat count (scanmatch.py:117-130)
        double x = x0 - cfg.sm.OffsetX, y = y0 - cfg.sm.OffsetY, t = t0;
        x = x < 0 ? 0.0 : x;  y = y < 0 ? 0.0 : y;  t = t < 0 ? 0.0 : t;  // d[d < 0] = 0, all columns
        x = x >= cfg.sm.Xres ? (double)(cfg.sm.Xres - 1) : x;
        y = y >= cfg.sm.Yres ? (double)(cfg.sm.Yres - 1) : y;
        const int xi = (int)x, yi = (int)y;                               // int(): toward zero
        const long long ti = (long long)t;
        sym[idx] = cfg.d_mask ? cfg.d_mask[(int64_t)yi * cfg.sm.Xres + xi]       // maskFromArray (scanmatch.py:199)
                              : (uint8_t)(cfg.d_ylut[yi] * cfg.sm.Xbin + cfg.d_xlut[xi]);
        int r = 1;
        if (cfg.sm.TempBin != 0.0) {
            double q = rint((double)ti / cfg.sm.TempBin);                 // numpy.round: half to even
            r = q > 2147483647.0 ? 2147483647 : (int)q;
        }
        run[idx] = r;
        // --- SED grid symbol on the raw coordinates (visual_attention_metrics.py:289-295)
        const int hs = cfg.sed_height / cfg.sed_n, ws = cfg.sed_width / cfg.sed_n;
        sed[idx] = floordiv_i32((int)x0, ws) + floordiv_i32((int)y0, hs) * cfg.sed_n;
        // --- STDE rescaling (:409-415)
        xyn[2 * idx] = x0 / cfg.stde_max_dim;
        xyn[2 * idx + 1] = y0 / cfg.stde_max_dim;
    }
}

// total with-duration string length per path (sum of runs), one warp per path
__global__ void __launch_bounds__(256)
prep_nwd_kernel(const int32_t *__restrict__ run, const int32_t *__restrict__ len, int64_t n_paths, int lmax,
                int32_t *__restrict__ nwd) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_paths; p += nwarps) {
        const int L = len[p];
        long long s = 0;
        for (int f = lane; f < L; f += 32) s += run[p * lmax + f];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) nwd[p] = s > 2147483647LL ? 2147483647 : (int32_t)s;
    }
}

}  // namespace spb

extern "C" int spb_prep_paths(const double *d_xyd, const int32_t *d_len, int64_t n_paths, int32_t lmax,
                              const spb_score_cfg *cfg, uint8_t *d_sym, int32_t *d_run, int32_t *d_nwd,
                              int32_t *d_sed, double *d_xyn, spb_stream stream) {
    SPB_CHECK_ARG(cfg != nullptr, "cfg is null");
    SPB_CHECK_ARG(n_paths >= 0 && lmax > 0, "bad sizes");
    SPB_CHECK_ARG(cfg->d_xlut && cfg->d_ylut, "cfg tables missing");
    SPB_CHECK_ARG(cfg->sed_n > 0 && cfg->sed_height >= cfg->sed_n && cfg->sed_width >= cfg->sed_n, "bad SED grid");
    if (n_paths == 0) return SPB_OK;
    SPB_CHECK_ARG(d_xyd && d_len && d_sym && d_run && d_nwd && d_sed && d_xyn, "null device pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = n_paths * (int64_t)lmax;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)spb::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    spb::prep_paths_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_xyd, d_len, n_paths, lmax, *cfg, d_sym, d_run, d_sed,
                                                             d_xyn);
    SPB_LAUNCH_CHECK();
    int64_t wblocks = (n_paths + 7) / 8;
    if (wblocks > cap) wblocks = cap;
    spb::prep_nwd_kernel<<<(unsigned)wblocks, 256, 0, s>>>(d_run, d_len, n_paths, lmax, d_nwd);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}
