// K2-K4 score_pairs: one warp per (human, simulated) scanpath pair.
//
// For each pair the warp stages both symbol packs in its shared-memory slice and
// runs, back to back:
//   * ScanMatch without duration: Needleman-Wunsch, f64      scanmatch.py:135-150,190-193
//   * SED: Levenshtein, int32 (bit-exact gate)               visual_attention_metrics.py:236-285
//   * ScanMatch with duration: the same NW on the run-length expanded strings
//   * STDE: pairwise distance tile + running window sums     visual_attention_metrics.py:332-441
// All three DPs use the same anti-diagonal wavefront: lane l owns a strip of C
// consecutive columns (simulated string) in registers and walks the rows (human
// string) one step behind lane l-1; one shuffle per step carries the strip's
// right edge to the next lane.  Strings longer than 32*8 columns are processed
// in 256-column panels whose right boundary column lives in a global workspace.
// The NW recurrences are evaluated in f64 in the reference's operation order, so
// the ScanMatch scores are bit-identical to numpy's; only STDE differs in the
// last ulps (summation order of the window means).
//
// Not HBM-bound (about 0.45 KB per 15 pairs): the limiter is
// instruction issue, so the roofline unit is DP cell-updates/s.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "score_common.cuh"

namespace spb {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxStrip = 8;                    // columns per lane in registers
constexpr int kPanelCols = 32 * kMaxStrip;      // 256

// ---------------------------------------------------------------------------
// Needleman-Wunsch over one panel of <= 32*C columns.
//   rows: human string, symbols (ar, ac) with run lengths arun (UNIT: every run is 1)
//   cols: simulated string, likewise.  F indices follow scanmatch.py: F is (n+1) x (m+1),
//   borders F[i][0] = gap*(i+1), F[0][j] = gap*(j+1).
// Returns F[n][col0 + pcols] broadcast to all lanes; `best` accumulates max(F) over
// the panel's interior cells when gap != 0.
// ---------------------------------------------------------------------------
template <int C, bool UNIT, bool GAP0>
__device__ __forceinline__ double nw_panel(const uint8_t *ar, const uint8_t *ac, const int *arun, int n,
                                           const uint8_t *br, const uint8_t *bc, const int *brun, int nb_runs,
                                           int col0, int pcols, const double *subd, int xbin, double gap,
                                           double *bnd, double &best, int lane) {
    const int nl = (pcols + C - 1) / C;          // active lanes
    const int j0 = col0 + lane * C;
    int brow[C], bcol[C];
    bool valid[C];
    {
        int acc = 0, r = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = j0 + c;
            valid[c] = (j < col0 + pcols);
            int si = 0;
            if (valid[c]) {
                if (UNIT) si = j;
                else {
                    while (r < nb_runs - 1 && acc + brun[r] <= j) { acc += brun[r]; ++r; }
                    si = r;
                }
            }
            brow[c] = br[si]; bcol[c] = bc[si];
        }
    }
    double prev[C];
#pragma unroll
    for (int c = 0; c < C; ++c) prev[c] = GAP0 ? 0.0 : gap * (double)(j0 + c + 2);
    double leftPrev = GAP0 ? 0.0 : gap * (double)(j0 + 1);
    double myLast = 0.0;
    int ri = 0, rem = UNIT ? 1 : arun[0];
    const int steps = n + nl - 1;
    for (int t = 0; t < steps; ++t) {
        const double recv = shfl_up_f64(myLast, 1);
        const int i = t - lane;
        if (i >= 0 && i < n && lane < nl) {
            double leftCur;
            if (lane == 0) leftCur = (col0 == 0) ? (GAP0 ? 0.0 : gap * (double)(i + 2)) : bnd[i + 1];
            else leftCur = recv;
            const int sa = UNIT ? i : ri;
            const int arow = ar[sa], acol = ac[sa];
            double d = leftPrev, l = leftCur;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (valid[c]) {
                    const double s = subd[abs(arow - brow[c]) * xbin + abs(acol - bcol[c])];
                    double v = d + s;                                   // match
                    const double ins = GAP0 ? l : l + gap;              // F[i][j-1] + gap
                    const double del = GAP0 ? prev[c] : prev[c] + gap;  // F[i-1][j] + gap
                    v = fmax(v, fmax(ins, del));
                    d = prev[c];
                    prev[c] = v;
                    l = v;
                    if (!GAP0) best = fmax(best, v);
                }
            }
            leftPrev = leftCur;
            myLast = l;
            if (bnd != nullptr && lane == nl - 1) bnd[i + 1] = l;      // right boundary column of this panel
            if (!UNIT) {
                if (--rem == 0) { ++ri; rem = (i + 1 < n) ? arun[ri] : 1; }
            }
        }
    }
    return shfl_idx_f64(myLast, nl - 1);
}

template <bool UNIT, bool GAP0>
__device__ __forceinline__ double nw_dispatch(const uint8_t *ar, const uint8_t *ac, const int *arun, int n,
                                              const uint8_t *br, const uint8_t *bc, const int *brun, int nb_runs,
                                              int col0, int pcols, const double *subd, int xbin, double gap,
                                              double *bnd, double &best, int lane) {
    if (pcols <= 32) return nw_panel<1, UNIT, GAP0>(ar, ac, arun, n, br, bc, brun, nb_runs, col0, pcols, subd, xbin, gap, bnd, best, lane);
    if (pcols <= 64) return nw_panel<2, UNIT, GAP0>(ar, ac, arun, n, br, bc, brun, nb_runs, col0, pcols, subd, xbin, gap, bnd, best, lane);
    if (pcols <= 128) return nw_panel<4, UNIT, GAP0>(ar, ac, arun, n, br, bc, brun, nb_runs, col0, pcols, subd, xbin, gap, bnd, best, lane);
    return nw_panel<8, UNIT, GAP0>(ar, ac, arun, n, br, bc, brun, nb_runs, col0, pcols, subd, xbin, gap, bnd, best, lane);
}

// Full NW score of strings with n rows / m columns (expanded lengths).
template <bool UNIT, bool GAP0>
__device__ double nw_score(const uint8_t *ar, const uint8_t *ac, const int *arun, int n, const uint8_t *br,
                           const uint8_t *bc, const int *brun, int nb_runs, int m, const double *subd, int xbin,
                           double gap, double max_sub, double *bnd, int64_t bnd_cap, int *err, int lane) {
    // max over the borders (scanmatch.py:139-143, :190): gap*(k+1), k = 0..max(n, m)
    double best = GAP0 ? 0.0 : fmax(gap, gap * (double)((n > m ? n : m) + 1));
    double corner = best;
    if (n > 0 && m > 0) {
        if (m > kPanelCols && (bnd == nullptr || bnd_cap < (int64_t)n + 1)) {
            if (lane == 0) atomicExch(err, 1);
            return nan("");
        }
        for (int col0 = 0; col0 < m; col0 += kPanelCols) {
            const int pcols = min(kPanelCols, m - col0);
            double *b = (m > kPanelCols) ? bnd : nullptr;
            corner = nw_dispatch<UNIT, GAP0>(ar, ac, arun, n, br, bc, brun, nb_runs, col0, pcols, subd, xbin, gap, b,
                                             best, lane);
            __syncwarp();
        }
        if (GAP0) best = corner;        // F is monotone for gap 0: max(F) = F[n][m]
        else {
            for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
    }
    return best / (max_sub * (double)(m > n ? m : n));      // 0/0 -> NaN like numpy
}

// ---------------------------------------------------------------------------
// Levenshtein distance, unit costs (visual_attention_metrics.py:236-285);
// rows a[0..n), columns b[0..m), m <= 32*C.
// ---------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ int lev_strip(const int *a, int n, const int *b, int m, int lane) {
    const int nl = (m + C - 1) / C;
    const int j0 = lane * C;
    int bs[C], prev[C];
    bool valid[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        valid[c] = (j0 + c < m);
        bs[c] = valid[c] ? b[j0 + c] : 0;
        prev[c] = j0 + c + 1;                       // D[0][j]
    }
    int leftPrev = j0;                              // D[0][j0]
    int myLast = 0;
    const int steps = n + nl - 1;
    for (int t = 0; t < steps; ++t) {
        const int recv = __shfl_up_sync(0xffffffffu, myLast, 1);
        const int i = t - lane;
        if (i >= 0 && i < n && lane < nl) {
            const int leftCur = (lane == 0) ? i + 1 : recv;      // D[i+1][0] = i+1
            const int sa = a[i];
            int d = leftPrev, l = leftCur;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (valid[c]) {
                    int v = min(min(prev[c] + 1, l + 1), d + (sa != bs[c] ? 1 : 0));
                    d = prev[c];
                    prev[c] = v;
                    l = v;
                }
            }
            leftPrev = leftCur;
            myLast = l;
        }
    }
    return __shfl_sync(0xffffffffu, myLast, nl - 1);
}

__device__ int lev_distance(const int *a, int n, const int *b, int m, int lane) {
    if (n == 0) return m;
    if (m == 0) return n;
    if (m <= 32) return lev_strip<1>(a, n, b, m, lane);
    if (m <= 64) return lev_strip<2>(a, n, b, m, lane);
    if (m <= 128) return lev_strip<4>(a, n, b, m, lane);
    return lev_strip<8>(a, n, b, m, lane);
}

// ---------------------------------------------------------------------------
// STDE (visual_attention_metrics.py:393-441): lane i owns simulated window start i.
// W_k[i][j] = sum_{t<k} D[i+t][j+t] is carried from k-1 (256 distances + running
// sums instead of re-summing every window).
// ---------------------------------------------------------------------------
__device__ double stde_similarity(const double *ax, const double *ay, int Lh, const double *bx, const double *by,
                                  int Ls, double *D, double *W, int pitch, int lane) {
    const int kmax = Lh < Ls ? Lh : Ls;
    if (kmax == 0) return nan("");
    for (int idx = lane; idx < Ls * Lh; idx += 32) {
        const int i = idx / Lh, j = idx - i * Lh;
        const double dx = bx[i] - ax[j], dy = by[i] - ay[j];
        D[i * pitch + j] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        W[i * pitch + j] = 0.0;
    }
    __syncwarp();
    double total = 0.0;
    for (int k = 1; k <= kmax; ++k) {
        const int nw = Ls - k + 1, nh = Lh - k + 1;
        double acc = 0.0;
        for (int i = lane; i < nw; i += 32) {
            double best = INFINITY;
            const double *dp = D + (i + k - 1) * pitch + (k - 1);
            double *wp = W + i * pitch;
            for (int j = 0; j < nh; ++j) {
                const double w = wp[j] + dp[j];
                wp[j] = w;
                best = fmin(best, w);
            }
            acc += best / (double)k;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        total += exp(-(acc / (double)nw));
    }
    return total / (double)kmax;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
score_pairs_kernel(spb_path_pack A, spb_path_pack B, const int32_t *__restrict__ pair_h,
                   const int32_t *__restrict__ pair_s, int64_t n_pairs, spb_score_cfg cfg,
                   double *__restrict__ scores, double *workspace, int64_t ws_per_warp, int *err) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int ntab = cfg.sm.Xbin * cfg.sm.Ybin;
    double *subd = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) subd[i] = cfg.d_sub_delta[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const PairLayout L = make_layout(A.lmax, B.lmax);
    unsigned char *base = smem + ((ntab * 8 + 15) & ~15) + (size_t)wib * L.bytes;
    double *ax = (double *)(base + L.ax), *ay = (double *)(base + L.ay);
    double *bx = (double *)(base + L.bx), *by = (double *)(base + L.by);
    double *D = (double *)(base + L.D), *W = (double *)(base + L.W);
    int *arun = (int *)(base + L.arun), *brun = (int *)(base + L.brun);
    int *ased = (int *)(base + L.ased), *bsed = (int *)(base + L.bsed);
    uint8_t *ar = base + L.ar, *ac = base + L.ac, *br = base + L.br, *bc = base + L.bc;
    uint8_t *awr = base + L.awr, *awc = base + L.awc, *bwr = base + L.bwr, *bwc = base + L.bwc;

    const int wpb = blockDim.x >> 5;                  // 8 unless long scanpaths need more shared memory per warp
    const int64_t gwarp = (int64_t)blockIdx.x * wpb + wib;
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    double *bnd = workspace ? workspace + gwarp * ws_per_warp : nullptr;
    const int xbin = cfg.sm.Xbin;
    const double gap = cfg.sm.GapValue;

    for (int64_t p = gwarp; p < n_pairs; p += nwarps) {
        const int64_t ia = pair_h[p], ib = pair_s[p];
        if (ia < 0 || ia >= A.n_paths || ib < 0 || ib >= B.n_paths) {      // stale / foreign pair map: fail loudly
            if (lane == 0) {
                atomicExch(err, 2);
                double *o = scores + 4 * p;
                o[0] = o[1] = o[2] = o[3] = nan("");
            }
            continue;
        }
        const int La = A.d_len[ia], Lb = B.d_len[ib];
        const int n_wd = A.d_nwd[ia], m_wd = B.d_nwd[ib];
        int na_runs = 0, nb_runs = 0;
        // ---- stage both packs; compact the zero-length runs out of the wd strings
        for (int f0 = 0; f0 < La; f0 += 32) {
            const int f = f0 + lane;
            int r = 0, s = 0;
            if (f < La) {
                const int64_t g = ia * A.lmax + f;
                s = A.d_sym[g]; r = A.d_run[g];
                ar[f] = (uint8_t)(s / xbin); ac[f] = (uint8_t)(s % xbin);
                ased[f] = A.d_sed[g];
                ax[f] = A.d_xyn[2 * g]; ay[f] = A.d_xyn[2 * g + 1];
            }
            const unsigned m = __ballot_sync(0xffffffffu, r > 0);
            if (r > 0) {
                const int pos = na_runs + __popc(m & ((1u << lane) - 1));
                awr[pos] = (uint8_t)(s / xbin); awc[pos] = (uint8_t)(s % xbin); arun[pos] = r;
            }
            na_runs += __popc(m);
        }
        for (int f0 = 0; f0 < Lb; f0 += 32) {
            const int f = f0 + lane;
            int r = 0, s = 0;
            if (f < Lb) {
                const int64_t g = ib * B.lmax + f;
                s = B.d_sym[g]; r = B.d_run[g];
                br[f] = (uint8_t)(s / xbin); bc[f] = (uint8_t)(s % xbin);
                bsed[f] = B.d_sed[g];
                bx[f] = B.d_xyn[2 * g]; by[f] = B.d_xyn[2 * g + 1];
            }
            const unsigned m = __ballot_sync(0xffffffffu, r > 0);
            if (r > 0) {
                const int pos = nb_runs + __popc(m & ((1u << lane) - 1));
                bwr[pos] = (uint8_t)(s / xbin); bwc[pos] = (uint8_t)(s % xbin); brun[pos] = r;
            }
            nb_runs += __popc(m);
        }
        __syncwarp();

        double wd, wod;
        if (gap == 0.0) {
            wod = nw_score<true, true>(ar, ac, nullptr, La, br, bc, nullptr, Lb, Lb, subd, xbin, gap, cfg.max_sub,
                                       nullptr, 0, err, lane);
            wd = nw_score<false, true>(awr, awc, arun, n_wd, bwr, bwc, brun, nb_runs, m_wd, subd, xbin, gap,
                                       cfg.max_sub, bnd, ws_per_warp, err, lane);
        } else {
            wod = nw_score<true, false>(ar, ac, nullptr, La, br, bc, nullptr, Lb, Lb, subd, xbin, gap, cfg.max_sub,
                                        nullptr, 0, err, lane);
            wd = nw_score<false, false>(awr, awc, arun, n_wd, bwr, bwc, brun, nb_runs, m_wd, subd, xbin, gap,
                                        cfg.max_sub, bnd, ws_per_warp, err, lane);
        }
        const int sed = lev_distance(ased, La, bsed, Lb, lane);
        const double stde = stde_similarity(ax, ay, La, bx, by, Lb, D, W, L.pitch, lane);
        if (lane == 0) {
            double *o = scores + 4 * p;
            o[0] = wd; o[1] = wod; o[2] = (double)sed; o[3] = stde;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// a7: per-group reduction of the score table (pairs_eval aggregation, OSIE/utils/evaluation.py:325-338) and the
// global sums behind `evaluation`'s mean / std / best (:211-237), in one pass over the table.
//   group g = one simulated path against the (padded) subject list of image g % n_images;
//   real pair      : s < count[image]                       (subjects padded in by pack_subject_lists are skipped)
//   surviving row  : real, both paths >= min_len_valid fixations (the MultiMatch NaN rule), no NaN score
//   table row      : sums of the surviving rows / count[image]  (the reference divides by len(gt), :329)
//   accumulators   : over ALL real pairs (evaluation() does not eliminate rows): sum, sum of squares of the four
//                    scores, of the per-group SED min / STDE max, pair and group counts.  Deterministic: per-block
//                    partials, summed in block order by the last block to finish (no floating-point atomics).
// ---------------------------------------------------------------------------
constexpr int kAccSlots = 16;     // 0-3 sum, 4-7 sumsq, 8-9 best sums, 10-11 best sumsq, 12 real pairs, 13 groups, 14 surviving groups

__global__ void __launch_bounds__(256)
reduce_pairs_kernel(spb_reduce_args a) {
    __shared__ double sh[8][kAccSlots];
    __shared__ bool last;
    double loc[kAccSlots];
#pragma unroll
    for (int i = 0; i < kAccSlots; ++i) loc[i] = 0.0;
    const int gs = a.group_size;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < a.n_groups;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int img = a.n_images > 0 ? (int)(g % a.n_images) : 0;
        const int cnt = a.d_group_count ? min(max(a.d_group_count[img], 0), gs) : gs;
        double s_wod = 0, s_wd = 0, s_sed = 0, s_stde = 0, best_sed = INFINITY, best_stde = -INFINITY;
        double all_best_sed = INFINITY, all_best_stde = -INFINITY;
        int kept = 0;
        for (int s = 0; s < cnt; ++s) {
            const int64_t p = g * gs + s;
            const double *r = a.d_scores + 4 * p;
            const double r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
            loc[0] += r0; loc[1] += r1; loc[2] += r2; loc[3] += r3;
            loc[4] += r0 * r0; loc[5] += r1 * r1; loc[6] += r2 * r2; loc[7] += r3 * r3;
            loc[12] += 1.0;
            all_best_sed = fmin(all_best_sed, r2); all_best_stde = fmax(all_best_stde, r3);
            bool ok = (a.d_valid == nullptr || a.d_valid[p]) && !(isnan(r0) || isnan(r1) || isnan(r2) || isnan(r3));
            if (ok && a.min_len_valid > 0 && a.d_pair_h != nullptr)
                ok = a.d_len_h[a.d_pair_h[p]] >= a.min_len_valid && a.d_len_s[a.d_pair_s[p]] >= a.min_len_valid;
            if (!ok) continue;
            ++kept;
            s_wd += r0; s_wod += r1; s_sed += r2; s_stde += r3;
            best_sed = fmin(best_sed, r2); best_stde = fmax(best_stde, r3);
        }
        if (cnt > 0) {
            loc[8] += all_best_sed; loc[9] += all_best_stde;
            loc[10] += all_best_sed * all_best_sed; loc[11] += all_best_stde * all_best_stde;
            loc[13] += 1.0;
        }
        const float qnan = __int_as_float(0x7fc00000);
        const int div = a.mean_over_kept ? kept : cnt;
        double rw = nan("");
        if (a.d_out) {
            float *o = a.d_out + 11 * g;
            if (kept > 0) {
                // slots 0..4 are MultiMatch's (external package, out of scope): a finite placeholder, so that the
                // reference's `np.any(np.isnan(metrics_reward))` trial rejection (train.py:237) sees NaN exactly
                // where the reference does -- when no row of the image survives
                o[0] = o[1] = o[2] = o[3] = o[4] = 0.0f;
                o[5] = (float)(s_wod / div); o[6] = (float)(s_wd / div);
                o[7] = (float)(s_sed / div); o[8] = (float)(s_stde / div);
                o[9] = (float)best_sed; o[10] = (float)best_stde;
            } else {
                for (int i = 0; i < 11; ++i) o[i] = qnan;
            }
        }
        if (kept > 0) {
            // the reward is computed from the float32 table (train.py:241,252: scipy.stats.hmean of slots 5, 6)
            const double x = (double)(float)(s_wod / div), y = (double)(float)(s_wd / div);
            rw = (x > 0.0 && y > 0.0) ? 2.0 / (1.0 / x + 1.0 / y) : 0.0;
            loc[14] += 1.0;
        }
        if (a.d_reward) a.d_reward[g] = rw;
        if (a.d_group_valid) a.d_group_valid[g] = kept > 0 ? 1 : 0;
    }
    if (a.d_acc == nullptr) return;
    // block partials in a fixed order: lanes -> warps -> blocks
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kAccSlots; ++i) {
        double v = loc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[w][i] = v;
    }
    __syncthreads();
    double *part = a.d_acc + kAccSlots;                 // [gridDim.x][kAccSlots], then the arrival counter
    unsigned int *counter = reinterpret_cast<unsigned int *>(a.d_acc + kAccSlots + (int64_t)a.acc_blocks * kAccSlots);
    if (threadIdx.x < kAccSlots) {
        double v = 0.0;
        for (int i = 0; i < 8; ++i) v += sh[i][threadIdx.x];
        part[(int64_t)blockIdx.x * kAccSlots + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < kAccSlots) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += part[(int64_t)b * kAccSlots + threadIdx.x];
        a.d_acc[threadIdx.x] += v;
    }
    if (threadIdx.x == 0) *counter = 0;                 // ready for the next launch
}

}  // namespace spb

extern "C" int64_t spb_score_workspace_bytes(int64_t max_human_nwd) {
    if (max_human_nwd < 0) max_human_nwd = 0;
    const int64_t per_warp = (max_human_nwd + 2 + 1) & ~(int64_t)1;
    return per_warp * 8 * 64 * spb::num_sms();               // <= 64 pairs in flight per SM (score_pairs_g8.cu)
}

extern "C" int spb_score_pairs(const spb_path_pack *human, const spb_path_pack *sim, const int32_t *d_pair_h,
                               const int32_t *d_pair_s, int64_t n_pairs, const spb_score_cfg *cfg, double *d_scores,
                               void *d_workspace, int64_t workspace_bytes, int32_t *d_err, spb_stream stream) {
    SPB_CHECK_ARG(human && sim && cfg, "null struct pointer");
    SPB_CHECK_ARG(n_pairs >= 0, "negative pair count");
    if (n_pairs == 0) return SPB_OK;
    SPB_CHECK_ARG(d_pair_h && d_pair_s && d_scores && d_err, "null device pointer");
    SPB_CHECK_ARG(cfg->d_sub_delta != nullptr, "cfg tables missing");
    SPB_CHECK_ARG(human->lmax > 0 && sim->lmax > 0, "bad lmax");
    if (sim->lmax > spb::kPanelCols) {
        spb::set_error("spb_score_pairs: simulated scanpaths longer than %d fixations are not supported", spb::kPanelCols);
        return SPB_ERR_UNSUPPORTED;
    }
    {   // fast path: four pairs per warp (score_pairs_g8.cu); declines long scanpaths and GapValue != 0
        int handled = 0;
        const int rc = spb::score_pairs_g8(*human, *sim, d_pair_h, d_pair_s, n_pairs, *cfg, d_scores, d_workspace,
                                           workspace_bytes, d_err, (cudaStream_t)stream, &handled);
        if (rc != SPB_OK || handled) return rc;
    }
    const spb::PairLayout L = spb::make_layout(human->lmax, sim->lmax);
    const int ntab_bytes = (cfg->sm.Xbin * cfg->sm.Ybin * 8 + 15) & ~15;
    // 8 warps per block unless the per-warp slice (the STDE tiles grow with lmax_h x lmax_s) needs more room:
    // long human scanpaths (the AiR / COCO evaluation sets do not truncate them) run with 4, 2 or 1 warps
    int wpb = spb::kWarpsPerBlock;
    while (wpb > 1 && (size_t)ntab_bytes + (size_t)L.bytes * wpb > 227 * 1024) wpb >>= 1;
    const size_t smem = (size_t)ntab_bytes + (size_t)L.bytes * wpb;
    if (smem > 227 * 1024) {
        spb::set_error("spb_score_pairs: lmax %d x %d needs %zu B of shared memory per warp (> 227 KB)", human->lmax,
                       sim->lmax, smem);
        return SPB_ERR_UNSUPPORTED;
    }
    SPB_CUDA(cudaFuncSetAttribute(spb::score_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spb::score_pairs_kernel, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * wpb > 32) per_sm = 32 / wpb;
    int64_t blocks = (int64_t)spb::num_sms() * per_sm;          // persistent grid, whole waves
    const int64_t need = (n_pairs + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    int64_t ws_per_warp = 0;
    if (d_workspace != nullptr && workspace_bytes > 0) {
        ws_per_warp = workspace_bytes / 8 / (blocks * wpb);
        ws_per_warp &= ~(int64_t)1;
    }
    spb::prof_begin(spb::kTagScore, (cudaStream_t)stream);
    spb::score_pairs_kernel<<<(unsigned)blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        *human, *sim, d_pair_h, d_pair_s, n_pairs, *cfg, d_scores, ws_per_warp > 0 ? (double *)d_workspace : nullptr,
        ws_per_warp, d_err);
    SPB_LAUNCH_CHECK();
    spb::prof_end((cudaStream_t)stream);
    return SPB_OK;
}

// ---------------------------------------------------------------------------
// ScanMatch.match's F matrix of ONE pair (scanmatch.py:138-150), for the single-pair API's alignment
// (utils/evaltools/scanmatch.py::ScanMatch.match): F [(n+1), (m+1)] row-major, borders GapValue * (index + 1),
// F[i][j] = max(F[i-1][j-1] + Sub[a[i-1]][b[j-1]], F[i][j-1] + gap, F[i-1][j] + gap) in f64, the reference's
// operations -- bit-identical.  One block sweeps the anti-diagonals (every cell of a diagonal depends on the two
// before it only); the O(n + m) traceback runs on the host from this matrix.
// ---------------------------------------------------------------------------
namespace spb {
__global__ void __launch_bounds__(256)
scanmatch_matrix_kernel(const int32_t *__restrict__ a, int n, const int32_t *__restrict__ b, int m,
                        const double *__restrict__ sub_delta, int xbin, int nbins, double gap, double *F) {
    const int ld = m + 1;
    for (int i = threadIdx.x; i <= n; i += blockDim.x) F[(int64_t)i * ld] = gap * (double)(i + 1);
    for (int j = threadIdx.x; j <= m; j += blockDim.x) F[j] = gap * (double)(j + 1);
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
        const int i_lo = max(1, d - m), i_hi = min(n, d - 1);
        for (int i = i_lo + threadIdx.x; i <= i_hi; i += blockDim.x) {
            const int j = d - i;
            const int sa = a[i - 1], sb = b[j - 1];
            // a symbol outside the bins (the reference raises IndexError; the Python mirror checks first) poisons its cells
            const bool ok = (unsigned)sa < (unsigned)nbins && (unsigned)sb < (unsigned)nbins;
            const double s = ok ? sub_delta[abs(sa / xbin - sb / xbin) * xbin + abs(sa % xbin - sb % xbin)] : nan("");
            const double mt = F[(int64_t)(i - 1) * ld + j - 1] + s;
            const double ins = F[(int64_t)i * ld + j - 1] + gap;
            const double del = F[(int64_t)(i - 1) * ld + j] + gap;
            F[(int64_t)i * ld + j] = fmax(fmax(mt, ins), del);
        }
        __syncthreads();
    }
}
}  // namespace spb

extern "C" int spb_scanmatch_matrix(const int32_t *d_a, int32_t n, const int32_t *d_b, int32_t m,
                                    const spb_score_cfg *cfg, double *d_F, spb_stream stream) {
    SPB_CHECK_ARG(cfg != nullptr && cfg->d_sub_delta != nullptr, "cfg tables missing");
    SPB_CHECK_ARG(n >= 0 && m >= 0 && d_F != nullptr, "bad sizes");
    SPB_CHECK_ARG((n == 0 || d_a) && (m == 0 || d_b), "null device pointer");
    spb::scanmatch_matrix_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_a, n, d_b, m, cfg->d_sub_delta, cfg->sm.Xbin,
                                                                      cfg->sm.Xbin * cfg->sm.Ybin, cfg->sm.GapValue, d_F);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

// ---------------------------------------------------------------------------
// The time-delay-embedding distances of ONE pair for every window length k (visual_attention_metrics.py:332-390),
// for the single-pair API (time_delay_embedding_distance, scaled_time_delay_embedding_distance, euclidean_distance):
//   out[k-1] = { mean over the simulated k-windows of (distance to the nearest human k-window) / k,   'Mean'
//                the maximum of the same                                                             'Hausdorff'
//                sum_{t<k} |s_t - h_t|  (euclidean_distance of the first k points) }
// One block; D [Ls, Lh] point distances and the running window sums W_k = W_{k-1} + D shifted (the reference's
// order of additions up to numpy's pairwise summation for k >= 8) live in the caller's work buffer.
// ---------------------------------------------------------------------------
namespace spb {
__global__ void __launch_bounds__(256)
tde_distances_kernel(const double *__restrict__ h_xy, int Lh, const double *__restrict__ s_xy, int Ls, double *work,
                     double *__restrict__ out) {
    double *D = work, *W = work + (int64_t)Ls * Lh, *best = W + (int64_t)Ls * Lh;
    const int kmax = Lh < Ls ? Lh : Ls;
    for (int idx = threadIdx.x; idx < Ls * Lh; idx += blockDim.x) {
        const int i = idx / Lh, j = idx - i * Lh;
        const double dx = s_xy[2 * i] - h_xy[2 * j], dy = s_xy[2 * i + 1] - h_xy[2 * j + 1];
        D[idx] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        W[idx] = 0.0;
    }
    __syncthreads();
    for (int k = 1; k <= kmax; ++k) {
        const int nw = Ls - k + 1, nh = Lh - k + 1;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            double b = INFINITY;
            for (int j = 0; j < nh; ++j) {
                const double w = W[(int64_t)i * Lh + j] + D[(int64_t)(i + k - 1) * Lh + (j + k - 1)];
                W[(int64_t)i * Lh + j] = w;
                b = fmin(b, fabs(w));
            }
            best[i] = b / (double)k;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0, mx = -INFINITY;
            for (int i = 0; i < nw; ++i) { sum += best[i]; mx = fmax(mx, best[i]); }   // Python's sum(): in order
            out[3 * (k - 1)] = sum / (double)nw;
            out[3 * (k - 1) + 1] = mx;
            out[3 * (k - 1) + 2] = W[0];
        }
        __syncthreads();
    }
}
}  // namespace spb

extern "C" int64_t spb_tde_work_bytes(int32_t Lh, int32_t Ls) {
    if (Lh <= 0 || Ls <= 0) return 8;
    return ((int64_t)2 * Lh * Ls + Ls) * 8;
}

extern "C" int spb_tde_distances(const double *d_h_xy, int32_t Lh, const double *d_s_xy, int32_t Ls, double *d_work,
                                 int64_t work_bytes, double *d_out, spb_stream stream) {
    SPB_CHECK_ARG(Lh >= 0 && Ls >= 0, "bad sizes");
    if (Lh == 0 || Ls == 0) return SPB_OK;
    SPB_CHECK_ARG(d_h_xy && d_s_xy && d_work && d_out, "null device pointer");
    if (work_bytes < spb_tde_work_bytes(Lh, Ls)) {
        spb::set_error("spb_tde_distances: work buffer too small (%lld < %lld bytes)", (long long)work_bytes,
                       (long long)spb_tde_work_bytes(Lh, Ls));
        return SPB_ERR_WORKSPACE;
    }
    spb::tde_distances_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_h_xy, Lh, d_s_xy, Ls, d_work, d_out);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

static int reduce_blocks(int64_t n_groups) {
    int64_t blocks = (n_groups + 255) / 256;
    const int64_t cap = (int64_t)spb::num_sms() * 4;
    return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int64_t spb_reduce_acc_bytes(void) {
    // 16 result slots + per-block partials of the largest grid + the arrival counter
    return (int64_t)(spb::kAccSlots + (int64_t)spb::num_sms() * 4 * spb::kAccSlots + 2) * 8;
}

extern "C" int spb_reduce_pairs(const spb_reduce_args *args, spb_stream stream) {
    SPB_CHECK_ARG(args != nullptr, "null struct pointer");
    SPB_CHECK_ARG(args->n_groups >= 0 && args->group_size > 0, "bad sizes");
    if (args->n_groups == 0) return SPB_OK;
    SPB_CHECK_ARG(args->d_scores != nullptr, "null device pointer");
    SPB_CHECK_ARG(args->d_group_count == nullptr || args->n_images > 0, "group counts need n_images");
    SPB_CHECK_ARG(args->min_len_valid <= 0 || args->d_pair_h == nullptr ||
                      (args->d_pair_s && args->d_len_h && args->d_len_s),
                  "the length rule needs pair maps and both length arrays");
    spb_reduce_args a = *args;
    const int blocks = reduce_blocks(a.n_groups);
    a.acc_blocks = (int32_t)((int64_t)spb::num_sms() * 4);
    if (a.d_acc) SPB_CHECK_ARG(a.acc_bytes >= spb_reduce_acc_bytes(), "accumulator buffer smaller than spb_reduce_acc_bytes()");
    spb::reduce_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    SPB_LAUNCH_CHECK();
    return SPB_OK;
}

extern "C" int spb_reduce_pairs_eval(const double *d_scores, const uint8_t *d_valid, int64_t n_groups,
                                     int32_t group_size, float *d_out, double *d_reward, spb_stream stream) {
    SPB_CHECK_ARG(n_groups >= 0 && group_size > 0, "bad sizes");
    if (n_groups == 0) return SPB_OK;
    SPB_CHECK_ARG(d_scores && d_out, "null device pointer");
    spb_reduce_args a;
    memset(&a, 0, sizeof(a));
    a.d_scores = d_scores; a.d_valid = d_valid; a.n_groups = n_groups; a.group_size = group_size;
    a.d_out = d_out; a.d_reward = d_reward;
    return spb_reduce_pairs(&a, stream);
}
