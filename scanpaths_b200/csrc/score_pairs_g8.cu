// K2-K4 fast path: FOUR (human, simulated) pairs per warp, eight lanes per pair.
//
// Same four scores and the same arithmetic as score_pairs.cu (ScanMatch with / without duration:
// scanmatch.py:135-150,190-193 in f64 in the reference's operation order -> bit-identical; SED:
// visual_attention_metrics.py:236-285, int32; STDE: :332-441, f64), for the sizes the drivers produce
// (simulated scanpaths <= 32 fixations, GapValue = 0).  Why another mapping: with one warp per pair the
// wavefront of a ~50-symbol string keeps 16 of 32 lanes busy and spends ~4 warp instructions per DP cell.
// Here lane l of an 8-lane group owns C consecutive columns (C = 8 for the with-duration strings: 64-column
// panels; C = 2 or 4 for the fixation strings) and walks the rows one step behind lane l-1, so a warp
// instruction updates 4 pairs x 8 lanes cells, the per-step overhead (one shuffle, the row symbol, the run
// counter) is amortised over C cells per lane, and the ramp of the wavefront is 7 steps instead of 25.
// Columns past the end of a string are computed like any other (their values never feed a valid cell:
// dependencies only run left / up), which removes the per-cell validity predicates.
// All loops are warp-uniform (trip counts are the maximum over the warp's four pairs, lanes predicate
// themselves off), so the four groups never diverge.
// Strings longer than a panel carry the panel's last column in a per-group global workspace.
#include <math.h>
#include <stdlib.h>

#include "score_common.cuh"

namespace spb {

constexpr int kG = 8;                     // lanes per pair
constexpr int kPairsPerWarp = 32 / kG;
constexpr int kG8Warps = 4;               // warps per block, at most (16 pairs per block); SPB_SCORE_WARPS=1|2|4 picks the launch size
constexpr int kWdC = 16;                  // with-duration columns per lane, at most (the warp uses ceil(m / 8), rounded up to 4/8/12/16)
constexpr int kWdPanel = kG * kWdC;       // 128

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }

// max(a, b) for finite doubles of which at most one is negative (every F value of the gap-0 recurrence is >= 0,
// only F[i-1][j-1] + s can dip below): IEEE order equals the order of the bit patterns as signed 64-bit integers,
// so the maximum is two integer compares and two selects on the ALU pipe instead of an FP64-pipe DSETP plus
// selects plus a NaN fix-up.  Exact: it returns one of its operands.
__device__ __forceinline__ double max_f64_bits(double a, double b) {
    const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    return __longlong_as_double(ia > ib ? ia : ib);
}
// min of two non-negative doubles (or +inf), same argument as above
__device__ __forceinline__ double min_f64_bits(double a, double b) {
    const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    return __longlong_as_double(ia < ib ? ia : ib);
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// One panel of the gap-0 Needleman-Wunsch recurrence for the group's pair; returns F[n][col0 + pcols]
// (group-uniform).  `active` false: the group idles through the warp's steps.
template <int C, bool UNIT>
__device__ __forceinline__ double nw_panel_g8(const uint8_t *ar, const uint8_t *ac, const int *arun, int n,
                                              const uint8_t *br, const uint8_t *bc, const int *brun, int nb_runs,
                                              int col0, int pcols, bool active, bool more_panels,
                                              const double *subd, int xbin, double *bnd, int gl) {
    const int nl = active ? (pcols + C - 1) / C : 0;          // lanes of the group that own columns
    const int j0 = col0 + gl * C;
    int brow[C], bcol[C];
    if (gl < nl) {
        int acc = 0, r = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = min(j0 + c, col0 + pcols - 1);       // columns past the end repeat the last symbol
            int si;
            if (UNIT) si = j;
            else {
                while (r < nb_runs - 1 && acc + brun[r] <= j) { acc += brun[r]; ++r; }
                si = r;
            }
            brow[c] = br[si]; bcol[c] = bc[si];
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) { brow[c] = 0; bcol[c] = 0; }
    }
    const int steps_w = warp_max(nl > 0 ? n + nl - 1 : 0);
    double prev[C];
#pragma unroll
    for (int c = 0; c < C; ++c) prev[c] = 0.0;
    double leftPrev = 0.0, myLast = 0.0;
    int ri = 0, rem = (UNIT || n == 0) ? 1 : arun[0];
#pragma unroll 1
    for (int t = 0; t < steps_w; ++t) {
        const double recv = __shfl_up_sync(0xffffffffu, myLast, 1, kG);
        const int i = t - gl;
        if (i >= 0 && i < n && gl < nl) {
            const double leftCur = (gl == 0) ? (col0 == 0 ? 0.0 : bnd[i + 1]) : recv;
            const int sa = UNIT ? i : ri;
            const int arow = ar[sa], acol = ac[sa];
            double d = leftPrev, l = leftCur;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double s = subd[abs(arow - brow[c]) * xbin + abs(acol - bcol[c])];
                const double v = max_f64_bits(d + s, max_f64_bits(l, prev[c]));   // match | F[i][j-1] | F[i-1][j]  (gap 0)
                d = prev[c];
                prev[c] = v;
                l = v;
            }
            leftPrev = leftCur;
            myLast = l;
            if (more_panels && gl == kG - 1) bnd[i + 1] = l;        // a full panel: its last column feeds the next one
            if (!UNIT) {
                if (--rem == 0) { ++ri; rem = (i + 1 < n) ? arun[ri] : 1; }
            }
        }
    }
    // F[n][col0 + pcols]: strip position (pcols - 1) % C of the last lane that owns columns
    double res = 0.0;
    const int cm = nl > 0 ? (pcols - 1) - (nl - 1) * C : 0;
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (c == cm) res = prev[c];
    return __shfl_sync(0xffffffffu, res, nl > 0 ? nl - 1 : 0, kG);
}

// The with-duration strings (run-length coded: ~10 runs of ~5 symbols) -- >80 % of all cell updates.
//   * the substitution score of a cell depends only on (human run, simulated run): a per-pair table
//     S[u][v] in shared memory (it aliases the STDE tile W, which is initialised later) turns the lookup
//     into one LDS at (row base + a per-column constant);
//   * every lane owns CW consecutive columns, CW = ceil(longest string of the warp's four pairs / 8) rounded up
//     to 4, 8, 12 or 16 -- a compile-time strip (four instantiations behind a warp-uniform switch: straight-line
//     code the scheduler can interleave, yet small enough for the instruction cache -- eight instantiations
//     made the kernel 145 KB and 60 % of its stalls instruction fetches): a 50-symbol string costs 8 cells per
//     step and an 80-symbol one 12 -- not two 64-column panels;
//   * the recurrence is evaluated in two phases: t = F[i-1][j-1] + s and m = max(t, F[i-1][j]) only read the
//     previous row (independent across the strip), then v = max(m, F[i][j-1]) is a running maximum -- the same
//     three operands per cell as scanmatch.py:146-148, bit-identical, with a dependent chain of one max per cell;
//   * the run counter advances branch-free.
template <int CW>
__device__ __noinline__ double nw_wd_g8(const double *S, int spitch, const int *arun, int n, const int *brun,
                                        int nb_runs, int col0, int pcols, bool active, bool more_panels,
                                        double *bnd, int gl) {
    const int nl = active ? (pcols + CW - 1) / CW : 0;        // lanes of the group that own columns
    const int j0 = col0 + gl * CW;
    uint32_t voff[CW];                                        // byte offset of each of my columns' run in a row of S
    {
        int acc = 0, r = 0;                                   // run r covers columns [acc, acc + brun[r])
        if (gl < nl)
            while (r < nb_runs - 1 && acc + brun[r] <= j0) { acc += brun[r]; ++r; }
        int end = (gl < nl) ? acc + brun[r] : 0;
#pragma unroll
        for (int c = 0; c < CW; ++c) {
            // runs are >= 1 symbol long (empty ones were compacted out): at most one step per column;
            // columns past the end of the string stay on the last run
            if (gl < nl && j0 + c >= end && r < nb_runs - 1) { ++r; end += brun[r]; }
            voff[c] = (uint32_t)r * 8u;
        }
    }
    const int steps_w = warp_max(nl > 0 ? n + nl - 1 : 0);
    double prev[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) prev[c] = 0.0;
    double leftPrev = 0.0, myLast = 0.0;
    int ri = 0, rem = n == 0 ? 1 : arun[0];
    uint32_t row = (uint32_t)__cvta_generic_to_shared(S);     // shared-space address of S[human run ri][0]
    const uint32_t row_step = (uint32_t)spitch * 8u;
#pragma unroll 1
    for (int t = 0; t < steps_w; ++t) {
        const double recv = __shfl_up_sync(0xffffffffu, myLast, 1, kG);
        const int i = t - gl;
        if (i >= 0 && i < n && gl < nl) {
            const double leftCur = (gl == 0) ? (col0 == 0 ? 0.0 : bnd[i + 1]) : recv;
            double mm[CW];
            double d = leftPrev;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const double tt = d + lds_f64(row + voff[c]);               // F[i-1][j-1] + s
                d = prev[c];
                mm[c] = max_f64_bits(tt, d);                                // | F[i-1][j]
            }
            double l = leftCur;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                l = max_f64_bits(mm[c], l);                                 // | F[i][j-1]
                prev[c] = l;
            }
            leftPrev = leftCur;
            myLast = l;
            if (more_panels && gl == kG - 1) bnd[i + 1] = l;                // a full panel: its last column feeds the next one
            --rem;                                                          // next row: advance the human run, branch-free
            const bool adv = rem == 0;
            ri += adv ? 1 : 0;
            row += adv ? row_step : 0u;
            const int nxt = arun[min(ri, n - 1)];                           // (runs <= symbols: always inside the slice)
            rem = adv ? ((i + 1 < n) ? nxt : 1) : rem;
        }
    }
    // F[n][col0 + pcols]: strip position of the last column in the last lane that owns columns
    double res = 0.0;
    const int cm = nl > 0 ? (pcols - 1) - (nl - 1) * CW : 0;
#pragma unroll
    for (int c = 0; c < CW; ++c)
        if (c == cm) res = prev[c];
    return __shfl_sync(0xffffffffu, res, nl > 0 ? nl - 1 : 0, kG);
}

// Levenshtein distance (unit costs), columns b[0..m) with m <= kG * C.
template <int C>
__device__ __forceinline__ int lev_g8(const int *a, int n, const int *b, int m, int gl) {
    const bool active = n > 0 && m > 0;
    const int nl = active ? (m + C - 1) / C : 0;
    const int j0 = gl * C;
    int bs[C], prev[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        bs[c] = (gl < nl) ? b[min(j0 + c, m - 1)] : 0;
        prev[c] = j0 + c + 1;                        // D[0][j]
    }
    int leftPrev = j0, myLast = 0;
    const int steps_w = warp_max(nl > 0 ? n + nl - 1 : 0);
#pragma unroll 1
    for (int t = 0; t < steps_w; ++t) {
        const int recv = __shfl_up_sync(0xffffffffu, myLast, 1, kG);
        const int i = t - gl;
        if (i >= 0 && i < n && gl < nl) {
            const int leftCur = (gl == 0) ? i + 1 : recv;
            const int sa = a[i];
            int d = leftPrev, l = leftCur;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int v = min(min(prev[c] + 1, l + 1), d + (sa != bs[c] ? 1 : 0));
                d = prev[c];
                prev[c] = v;
                l = v;
            }
            leftPrev = leftCur;
            myLast = l;
        }
    }
    int res = 0;
    const int cm = nl > 0 ? (m - 1) - (nl - 1) * C : 0;
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (c == cm) res = prev[c];
    res = __shfl_sync(0xffffffffu, res, nl > 0 ? nl - 1 : 0, kG);
    return n == 0 ? m : (m == 0 ? n : res);
}

// STDE (visual_attention_metrics.py:393-441): the group's lanes share the simulated window starts;
// W_k[i][j] = sum_{t<k} D[i+t][j+t] is carried from k-1 to k.
__device__ __forceinline__ double stde_g8(const double *ax, const double *ay, int Lh, const double *bx,
                                          const double *by, int Ls, double *D, double *W, int pitch, int gl) {
    const int kmax = Lh < Ls ? Lh : Ls;
    const int tot = Ls * Lh, tot_w = warp_max(tot);
    for (int idx0 = 0; idx0 < tot_w; idx0 += kG) {
        const int idx = idx0 + gl;
        if (idx < tot) {
            const int i = idx / Lh, j = idx - i * Lh;
            const double dx = bx[i] - ax[j], dy = by[i] - ay[j];
            D[i * pitch + j] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            W[i * pitch + j] = 0.0;
        }
    }
    __syncwarp();
    const int kmax_w = warp_max(kmax);
    double total = 0.0;
#pragma unroll 1
    for (int k = 1; k <= kmax_w; ++k) {
        const bool kon = k <= kmax;
        const int nw = kon ? Ls - k + 1 : 0, nh = kon ? Lh - k + 1 : 0;
        const int nw_w = warp_max(nw), nh_w = warp_max(nh);
        double acc = 0.0;
#pragma unroll 1
        for (int i0 = 0; i0 < nw_w; i0 += kG) {
            const int i = i0 + gl;
            const bool on = i < nw;
            double best = INFINITY;
            const double *dp = D + (on ? (i + k - 1) * pitch + (k - 1) : 0);
            double *wp = W + (on ? i * pitch : 0);
#pragma unroll 2
            for (int j = 0; j < nh_w; ++j) {
                if (on && j < nh) {
                    const double w = wp[j] + dp[j];
                    wp[j] = w;
                    best = min_f64_bits(best, w);       // distances are >= 0
                }
            }
            if (on) acc += best / (double)k;
        }
        for (int o = kG / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, kG);
        if (kon) total += exp(-(acc / (double)nw));
    }
    return kmax == 0 ? nan("") : total / (double)kmax;
}

template <int LC>
__global__ void __launch_bounds__(kG8Warps * 32)
score_pairs_g8_kernel(spb_path_pack A, spb_path_pack B, const int32_t *__restrict__ pair_h,
                      const int32_t *__restrict__ pair_s, int64_t n_pairs, spb_score_cfg cfg,
                      double *__restrict__ scores, double *workspace, int64_t ws_per_group, int *err) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int ntab = cfg.sm.Xbin * cfg.sm.Ybin;
    double *subd = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) subd[i] = cfg.d_sub_delta[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / kG, gl = lane % kG;
    const PairLayout L = make_layout(A.lmax, B.lmax);
    unsigned char *base = smem + ((ntab * 8 + 15) & ~15) + (size_t)(wib * kPairsPerWarp + sub) * L.bytes;
    double *ax = (double *)(base + L.ax), *ay = (double *)(base + L.ay);
    double *bx = (double *)(base + L.bx), *by = (double *)(base + L.by);
    double *D = (double *)(base + L.D), *W = (double *)(base + L.W);
    int *arun = (int *)(base + L.arun), *brun = (int *)(base + L.brun);
    int *ased = (int *)(base + L.ased), *bsed = (int *)(base + L.bsed);
    uint8_t *ar = base + L.ar, *ac = base + L.ac, *br = base + L.br, *bc = base + L.bc;
    uint8_t *awr = base + L.awr, *awc = base + L.awc, *bwr = base + L.bwr, *bwc = base + L.bwc;

    const int wpb = blockDim.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * wpb + wib;
    const int64_t ngroups = (int64_t)gridDim.x * wpb * kPairsPerWarp;
    double *bnd = workspace ? workspace + (gwarp * kPairsPerWarp + sub) * ws_per_group : nullptr;
    const int xbin = cfg.sm.Xbin;
    const unsigned gshift = (unsigned)(sub * kG);

    for (int64_t p0 = gwarp * kPairsPerWarp; p0 < n_pairs; p0 += ngroups) {       // warp-uniform
        const int64_t p = p0 + sub;
        bool has = p < n_pairs;
        int64_t ia = 0, ib = 0;
        if (has) {
            ia = pair_h[p]; ib = pair_s[p];
            if (ia < 0 || ia >= A.n_paths || ib < 0 || ib >= B.n_paths) {          // stale / foreign pair map: fail loudly
                if (gl == 0) {
                    atomicExch(err, 2);
                    double *o = scores + 4 * p;
                    o[0] = o[1] = o[2] = o[3] = nan("");
                }
                has = false;
            }
        }
        const int La = has ? A.d_len[ia] : 0, Lb = has ? B.d_len[ib] : 0;
        const int n_wd = has ? A.d_nwd[ia] : 0, m_wd = has ? B.d_nwd[ib] : 0;
        // ---- stage both packs in the group's slice; compact the zero-length runs out of the wd strings
        int na_runs = 0, nb_runs = 0;
        const int la_w = warp_max(La), lb_w = warp_max(Lb);
        for (int f0 = 0; f0 < la_w; f0 += kG) {
            const int f = f0 + gl;
            int r = 0, s = 0;
            if (f < La) {
                const int64_t g = ia * A.lmax + f;
                s = A.d_sym[g]; r = A.d_run[g];
                ar[f] = (uint8_t)(s / xbin); ac[f] = (uint8_t)(s % xbin);
                ased[f] = A.d_sed[g];
                ax[f] = A.d_xyn[2 * g]; ay[f] = A.d_xyn[2 * g + 1];
            }
            const unsigned m = (__ballot_sync(0xffffffffu, r > 0) >> gshift) & 0xffu;
            if (r > 0) {
                const int pos = na_runs + __popc(m & ((1u << gl) - 1));
                awr[pos] = (uint8_t)(s / xbin); awc[pos] = (uint8_t)(s % xbin); arun[pos] = r;
            }
            na_runs += __popc(m);
        }
        for (int f0 = 0; f0 < lb_w; f0 += kG) {
            const int f = f0 + gl;
            int r = 0, s = 0;
            if (f < Lb) {
                const int64_t g = ib * B.lmax + f;
                s = B.d_sym[g]; r = B.d_run[g];
                br[f] = (uint8_t)(s / xbin); bc[f] = (uint8_t)(s % xbin);
                bsed[f] = B.d_sed[g];
                bx[f] = B.d_xyn[2 * g]; by[f] = B.d_xyn[2 * g + 1];
            }
            const unsigned m = (__ballot_sync(0xffffffffu, r > 0) >> gshift) & 0xffu;
            if (r > 0) {
                const int pos = nb_runs + __popc(m & ((1u << gl) - 1));
                bwr[pos] = (uint8_t)(s / xbin); bwc[pos] = (uint8_t)(s % xbin); brun[pos] = r;
            }
            nb_runs += __popc(m);
        }
        __syncwarp();

        // ---- ScanMatch without duration: fixation strings, one strip of <= 8 * LC columns
        double wod;
        {
            const bool act = La > 0 && Lb > 0;
            const double corner = nw_panel_g8<LC, true>(ar, ac, nullptr, La, br, bc, nullptr, Lb, 0, Lb, act, false, subd,
                                                        xbin, nullptr, gl);
            wod = (act ? corner : 0.0) / (cfg.max_sub * (double)(Lb > La ? Lb : La));      // 0/0 -> NaN like numpy
        }
        // ---- SED
        const int sed = lev_g8<LC>(ased, La, bsed, Lb, gl);
        // ---- ScanMatch with duration: run-length strings
        double wd = nan("");
        {
            bool act = n_wd > 0 && m_wd > 0;
            if (act && m_wd > kWdPanel && (bnd == nullptr || ws_per_group < (int64_t)n_wd + 1)) {
                if (gl == 0) atomicExch(err, 1);
                act = false;
            }
            // substitution table of the pair: S[human run][simulated run] (aliases W, pitch = simulated lmax)
            double *S = W;
            const int spitch = B.lmax;
            const int ntab_w = warp_max(act ? na_runs * nb_runs : 0);
            for (int e0 = 0; e0 < ntab_w; e0 += kG) {
                const int e = e0 + gl;
                if (act && e < na_runs * nb_runs) {
                    const int u = e / nb_runs, v = e - u * nb_runs;
                    S[u * spitch + v] = subd[abs((int)awr[u] - (int)bwr[v]) * xbin + abs((int)awc[u] - (int)bwc[v])];
                }
            }
            __syncwarp();
            const int m_w = warp_max(act ? min(m_wd, kWdPanel) : 0);
            const int cw = (m_w + kG - 1) / kG;                      // columns per lane, warp-uniform, <= 16
            const int npan = act ? (m_wd + kWdPanel - 1) / kWdPanel : 0;
            const int npan_w = warp_max(npan);
            double corner = 0.0;
#pragma unroll 1
            for (int pn = 0; pn < npan_w; ++pn) {
                const bool on = pn < npan;
                const int col0 = pn * kWdPanel;
                const int pcols = on ? min(kWdPanel, m_wd - col0) : 0;
                double c;
                switch ((cw + 3) >> 2) {                               // strip width rounded up to 4, 8, 12, 16
                case 1: c = nw_wd_g8<4>(S, spitch, arun, n_wd, brun, nb_runs, col0, pcols, on, pn + 1 < npan, bnd, gl); break;
                case 2: c = nw_wd_g8<8>(S, spitch, arun, n_wd, brun, nb_runs, col0, pcols, on, pn + 1 < npan, bnd, gl); break;
                case 3: c = nw_wd_g8<12>(S, spitch, arun, n_wd, brun, nb_runs, col0, pcols, on, pn + 1 < npan, bnd, gl); break;
                default: c = nw_wd_g8<16>(S, spitch, arun, n_wd, brun, nb_runs, col0, pcols, on, pn + 1 < npan, bnd, gl); break;
                }
                if (on) corner = c;
                __syncwarp();
            }
            if (act || !(n_wd > 0 && m_wd > 0))
                wd = corner / (cfg.max_sub * (double)(m_wd > n_wd ? m_wd : n_wd));
        }
        // ---- STDE
        const double stde = stde_g8(ax, ay, La, bx, by, Lb, D, W, L.pitch, gl);
        if (gl == 0 && has) {
            double *o = scores + 4 * p;
            o[0] = wd; o[1] = wod; o[2] = (double)sed; o[3] = stde;
        }
        __syncwarp();
    }
}

static int g_score_warps = kG8Warps;

int score_pairs_g8(const spb_path_pack &A, const spb_path_pack &B, const int32_t *pair_h, const int32_t *pair_s,
                   int64_t n_pairs, const spb_score_cfg &cfg, double *scores, void *workspace, int64_t workspace_bytes,
                   int32_t *err, cudaStream_t stream, int *handled) {
    *handled = 0;
    if (cfg.sm.GapValue != 0.0 || B.lmax > kG * 4) return SPB_OK;
    const PairLayout L = make_layout(A.lmax, B.lmax);
    const int ntab_bytes = (cfg.sm.Xbin * cfg.sm.Ybin * 8 + 15) & ~15;
    // warps per block: 4 by default; SPB_SCORE_WARPS=1|2|4 picks the launch size.  (One-warp blocks -- 21 KB of shared
    // memory, 5 K registers -- were meant to run beside a resident gate-GEMM CTA of the next wave; they cannot: the
    // register file is partitioned per SM sub-partition and the GEMM's 10 warps x 168 registers fill two of the
    // four.  Measured: no gain)
    static int wpb_env = -1;
    if (wpb_env < 0) {
        const char *e = getenv("SPB_SCORE_WARPS");
        wpb_env = e ? atoi(e) : 0;
        if (wpb_env != 1 && wpb_env != 2 && wpb_env != 4) wpb_env = 0;
    }
    int wpb = wpb_env ? wpb_env : g_score_warps;
    if ((size_t)ntab_bytes + (size_t)L.bytes * wpb * kPairsPerWarp > 227 * 1024) return SPB_OK;
    const size_t smem = (size_t)ntab_bytes + (size_t)L.bytes * wpb * kPairsPerWarp;
    const int lc = B.lmax <= kG * 2 ? 2 : 4;
    auto kern = lc == 2 ? score_pairs_g8_kernel<2> : score_pairs_g8_kernel<4>;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * wpb > 16) per_sm = 16 / wpb;                     // <= 64 groups per SM (workspace sizing)
    int64_t blocks = (int64_t)num_sms() * per_sm;                  // persistent grid, whole waves
    const int64_t per_block = wpb * kPairsPerWarp;
    const int64_t need = (n_pairs + per_block - 1) / per_block;
    if (blocks > need) blocks = need;
    int64_t ws_per_group = 0;
    if (workspace != nullptr && workspace_bytes > 0) {
        ws_per_group = workspace_bytes / 8 / (blocks * per_block);
        ws_per_group &= ~(int64_t)1;
    }
    prof_begin(kTagScore, stream);
    kern<<<(unsigned)blocks, wpb * 32, smem, stream>>>(A, B, pair_h, pair_s, n_pairs, cfg, scores,
                                                            ws_per_group > 0 ? (double *)workspace : nullptr, ws_per_group,
                                                            err);
    SPB_LAUNCH_CHECK();
    prof_end(stream);
    *handled = 1;
    return SPB_OK;
}

}  // namespace spb
