// Error plumbing + host-side ScanMatch tables (C ABI, see include/scanpaths_b200.h).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace spb {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace spb

extern "C" int spb_version(void) { return 100; }

extern "C" const char *spb_last_error(void) { return spb::g_err; }

// Mirrors ScanMatch.CreateSubMatrix / GridMask (utils/evaltools/scanmatch.py:88-114).
// SubMatrix[a, b] = |D - Dmax| - (Dmax - Threshold), D = Euclid distance between
// the (row, col) of bins a and b: a function of (|d row|, |d col|) only, so the
// device needs Ybin*Xbin doubles instead of nb*nb; the values are computed with
// the same IEEE operations as numpy and are bit-equal to the reference's table.
extern "C" int spb_scanmatch_tables(const spb_scanmatch_cfg *cfg, double *h_sub_delta, double *h_sub_full,
                                    uint8_t *h_xlut, uint8_t *h_ylut, double *h_max_sub) {
    SPB_CHECK_ARG(cfg != nullptr, "cfg is null");
    SPB_CHECK_ARG(cfg->Xbin > 0 && cfg->Ybin > 0 && cfg->Xres > 0 && cfg->Yres > 0, "non-positive size");
    SPB_CHECK_ARG(cfg->Xbin * cfg->Ybin <= 256 && cfg->Xbin <= 255 && cfg->Ybin <= 255,
                  "more than 256 bins (symbols are u8)");
    const int xb = cfg->Xbin, yb = cfg->Ybin, nb = xb * yb;
    const double dmax = sqrt((double)((xb - 1) * (xb - 1) + (yb - 1) * (yb - 1)));
    double mx = -INFINITY;
    for (int dr = 0; dr < yb; ++dr)
        for (int dc = 0; dc < xb; ++dc) {
            double v = fabs(sqrt((double)(dc * dc + dr * dr)) - dmax) - (dmax - cfg->Threshold);
            if (h_sub_delta) h_sub_delta[dr * xb + dc] = v;
            if (v > mx) mx = v;
        }
    if (h_max_sub) *h_max_sub = mx;
    if (h_sub_full)
        for (int a = 0; a < nb; ++a)
            for (int b = 0; b < nb; ++b) {
                int dc = abs(a % xb - b % xb), dr = abs(a / xb - b / xb);
                h_sub_full[a * nb + b] = fabs(sqrt((double)(dc * dc + dr * dr)) - dmax) - (dmax - cfg->Threshold);
            }
    // numpy.int32(numpy.arange(0, Xbin, Xbin/Xres))[i] == trunc(i * step)
    const double sx = (double)xb / cfg->Xres, sy = (double)yb / cfg->Yres;
    if (h_xlut)
        for (int i = 0; i < cfg->Xres; ++i) {
            int v = (int)(i * sx);
            h_xlut[i] = (uint8_t)(v < xb ? v : xb - 1);
        }
    if (h_ylut)
        for (int i = 0; i < cfg->Yres; ++i) {
            int v = (int)(i * sy);
            h_ylut[i] = (uint8_t)(v < yb ? v : yb - 1);
        }
    return SPB_OK;
}
