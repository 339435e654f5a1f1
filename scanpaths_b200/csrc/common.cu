// Error plumbing + host-side ScanMatch tables (C ABI, see include/scanpaths_b200.h).
#include <math.h>
#include <atomic>
#include <mutex>
#include <vector>
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

namespace spb {
int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;
    }
    cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct Prof {
    std::vector<cudaEvent_t> ev;     // 2 per pair
    std::vector<int> tags;
    int cap = 0, n = 0;
    bool open = false;
};
static Prof g_prof;
static std::mutex g_prof_mu;

void prof_begin(int tag, cudaStream_t s) {
    if (g_prof.cap == 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.n >= g_prof.cap || g_prof.open) return;
    g_prof.tags[g_prof.n] = tag;
    cudaEventRecord(g_prof.ev[2 * g_prof.n], s);
    g_prof.open = true;
}
void prof_end(cudaStream_t s) {
    if (g_prof.cap == 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.open) return;
    cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], s);
    g_prof.open = false;
    ++g_prof.n;
}
}  // namespace spb

extern "C" int spb_version(void) { return 100; }

extern "C" int64_t spb_kernel_launches(void) { return spb::g_launches.load(); }

extern "C" int spb_profile_enable(int32_t max_pairs) {
    std::lock_guard<std::mutex> lk(spb::g_prof_mu);
    for (cudaEvent_t e : spb::g_prof.ev) cudaEventDestroy(e);
    spb::g_prof.ev.clear(); spb::g_prof.tags.clear();
    spb::g_prof.cap = 0; spb::g_prof.n = 0; spb::g_prof.open = false;
    if (max_pairs <= 0) return SPB_OK;
    spb::g_prof.ev.resize(2 * (size_t)max_pairs);
    spb::g_prof.tags.resize(max_pairs);
    for (auto &e : spb::g_prof.ev) SPB_CUDA(cudaEventCreate(&e));
    spb::g_prof.cap = max_pairs;
    return SPB_OK;
}

extern "C" int spb_profile_collect(float *h_ms, int32_t *h_tags, int32_t cap, int32_t *n_out) {
    SPB_CHECK_ARG(h_ms && h_tags && n_out, "null pointer");
    std::lock_guard<std::mutex> lk(spb::g_prof_mu);
    int n = spb::g_prof.n < cap ? spb::g_prof.n : cap;
    for (int i = 0; i < n; ++i) {
        SPB_CUDA(cudaEventSynchronize(spb::g_prof.ev[2 * i + 1]));
        SPB_CUDA(cudaEventElapsedTime(&h_ms[i], spb::g_prof.ev[2 * i], spb::g_prof.ev[2 * i + 1]));
        h_tags[i] = spb::g_prof.tags[i];
    }
    *n_out = n;
    spb::g_prof.n = 0;
    return SPB_OK;
}

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
