/* CPU oracle in plain C: ScanMatch / SED / STDE scoring of (human, prediction)
 * pairs.  TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path at sizes
 * the Python oracle cannot reach, never shipped, never on the product path.
 *
 * Restates (all paths under /root/reference, identical in the three task trees):
 *   OSIE/utils/evaltools/scanmatch.py:88-114      substitution matrix + grid mask
 *   OSIE/utils/evaltools/scanmatch.py:116-133     fixationToSequence
 *   OSIE/utils/evaltools/scanmatch.py:135-150,190-193  Needleman-Wunsch score
 *   OSIE/utils/evaltools/visual_attention_metrics.py:236-298  SED
 *   OSIE/utils/evaltools/visual_attention_metrics.py:205-218,332-441  STDE
 * Pinned against the reference's outputs in tests/golden/scoring_*.npz
 * (tests/test_oracle_c.py) and against the Python oracle on random inputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int Xres, Yres, Xbin, Ybin;
    double Threshold, GapValue, TempBin, OffsetX, OffsetY;
} spo_cfg;

/* scanmatch.py:88-103 */
void spo_sub_matrix(const spo_cfg *c, double *sub /* [nb*nb] */) {
    int nb = c->Xbin * c->Ybin;
    double mx = 0.0;
    for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) {
            int dc = a % c->Xbin - b % c->Xbin, dr = a / c->Xbin - b / c->Xbin;
            double d = sqrt((double)(dc * dc + dr * dr));
            sub[a * nb + b] = d;
            if (d > mx) mx = d;
        }
    for (int i = 0; i < nb * nb; ++i) sub[i] = fabs(sub[i] - mx) - (mx - c->Threshold);
}

/* scanmatch.py:105-114: xi = int32(arange(0, Xbin, Xbin/Xres)) i.e. trunc(i*step) */
static int spo_symbol(const spo_cfg *c, long xi, long yi) {
    double mx = (double)c->Xbin / c->Xres, my = (double)c->Ybin / c->Yres;
    int col = (int)(xi * mx), row = (int)(yi * my);
    return row * c->Xbin + col;
}

/* scanmatch.py:116-133.  Returns the sequence length; writes at most cap symbols. */
long spo_fixation_to_sequence(const spo_cfg *c, const double *xyd, int L, int32_t *out, long cap) {
    long n = 0;
    for (int f = 0; f < L; ++f) {
        double x = xyd[3 * f] - c->OffsetX, y = xyd[3 * f + 1] - c->OffsetY, t = xyd[3 * f + 2];
        if (x < 0) x = 0;
        if (y < 0) y = 0;
        if (t < 0) t = 0;
        if (x >= c->Xres) x = c->Xres - 1;
        if (y >= c->Yres) y = c->Yres - 1;
        long xi = (long)x, yi = (long)y, ti = (long)t;            /* int(): toward zero */
        int sym = spo_symbol(c, xi, yi);
        long reps = 1;
        if (c->TempBin != 0) reps = (long)nearbyint((double)ti / c->TempBin); /* half-to-even */
        for (long r = 0; r < reps; ++r) {
            if (n < cap) out[n] = sym;
            ++n;
        }
    }
    return n;
}

/* scanmatch.py:135-150, 190-193 */
double spo_nw_score(const spo_cfg *c, const double *sub, const int32_t *A, long n, const int32_t *B, long m) {
    int nb = c->Xbin * c->Ybin;
    double gap = c->GapValue, best, maxsub = sub[0];
    double *prev = (double *)malloc(sizeof(double) * (m + 1)), *cur = (double *)malloc(sizeof(double) * (m + 1));
    for (int i = 0; i < nb * nb; ++i)
        if (sub[i] > maxsub) maxsub = sub[i];
    best = gap;                                  /* F[0][0] = gap * (0 + 1) */
    for (long j = 0; j <= m; ++j) {
        prev[j] = gap * (j + 1);
        if (prev[j] > best) best = prev[j];
    }
    for (long i = 1; i <= n; ++i) {
        cur[0] = gap * (i + 1);
        if (cur[0] > best) best = cur[0];
        for (long j = 1; j <= m; ++j) {
            double d = prev[j - 1] + sub[(long)A[i - 1] * nb + B[j - 1]];
            double u = prev[j] + gap, l = cur[j - 1] + gap;
            double v = d > l ? d : l;
            v = v > u ? v : u;
            cur[j] = v;
            if (v > best) best = v;
        }
        double *t = prev; prev = cur; cur = t;
    }
    free(prev); free(cur);
    return best / (maxsub * (double)(m > n ? m : n));           /* 0/0 -> NaN as in numpy */
}

static int32_t floordiv(int32_t a, int32_t b) {
    int32_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

/* visual_attention_metrics.py:288-298 */
void spo_sed_symbols(const double *xyd, int L, int height, int width, int n, int32_t *out) {
    int hs = height / n, ws = width / n;
    for (int i = 0; i < L; ++i) {
        int32_t x = (int32_t)xyd[3 * i], y = (int32_t)xyd[3 * i + 1];
        out[i] = floordiv(x, ws) + floordiv(y, hs) * n;
    }
}

/* visual_attention_metrics.py:236-285 */
int spo_levenshtein(const int32_t *a, int la, const int32_t *b, int lb) {
    int *prev = (int *)malloc(sizeof(int) * (lb + 1)), *cur = (int *)malloc(sizeof(int) * (lb + 1));
    for (int j = 0; j <= lb; ++j) prev[j] = j;
    for (int i = 1; i <= la; ++i) {
        cur[0] = i;
        for (int j = 1; j <= lb; ++j) {
            int ins = prev[j] + 1, del = cur[j - 1] + 1, sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
            int v = ins < del ? ins : del;
            cur[j] = v < sub ? v : sub;
        }
        int *t = prev; prev = cur; cur = t;
    }
    int r = prev[lb];
    free(prev); free(cur);
    return r;
}

/* visual_attention_metrics.py:393-441 (human = first argument).  NaN when a
 * scanpath is empty (the reference returns None). */
double spo_stde(const double *human, int Lh, const double *sim, int Ls, double max_dim) {
    int kmax = Lh < Ls ? Lh : Ls;
    if (kmax == 0) return NAN;
    double total = 0.0;
    for (int k = 1; k <= kmax; ++k) {
        double acc = 0.0;
        int nw = Ls - k + 1;
        for (int i = 0; i < nw; ++i) {
            double best = INFINITY;
            for (int j = 0; j + k <= Lh; ++j) {
                double s = 0.0;
                for (int t = 0; t < k; ++t) {
                    double dx = sim[3 * (i + t)] / max_dim - human[3 * (j + t)] / max_dim;
                    double dy = sim[3 * (i + t) + 1] / max_dim - human[3 * (j + t) + 1] / max_dim;
                    s += sqrt(dx * dx + dy * dy);
                }
                if (s < best) best = s;
            }
            acc += best / k;
        }
        total += exp(-(acc / nw));
    }
    return total / kmax;
}

/* One (human, prediction) pair in the evaluation configuration
 * (OSIE/utils/evaluation.py:177-204): out = (SM with duration, SM w/o duration, SED, STDE).
 * xyd arrays are [L,3] (x, y, ms). */
void spo_score_pair(const spo_cfg *cfg_wd, const double *sub, const double *gt, int Lg, const double *pr, int Lp,
                    int height, int width, double *out4) {
    spo_cfg wod = *cfg_wd;
    wod.TempBin = 0;
    long cap_g = spo_fixation_to_sequence(cfg_wd, gt, Lg, NULL, 0);
    long cap_p = spo_fixation_to_sequence(cfg_wd, pr, Lp, NULL, 0);
    int32_t *sg = (int32_t *)malloc(sizeof(int32_t) * (cap_g + Lg + 1));
    int32_t *sp = (int32_t *)malloc(sizeof(int32_t) * (cap_p + Lp + 1));
    spo_fixation_to_sequence(cfg_wd, gt, Lg, sg, cap_g);
    spo_fixation_to_sequence(cfg_wd, pr, Lp, sp, cap_p);
    out4[0] = spo_nw_score(cfg_wd, sub, sg, cap_g, sp, cap_p);
    spo_fixation_to_sequence(&wod, gt, Lg, sg, Lg);
    spo_fixation_to_sequence(&wod, pr, Lp, sp, Lp);
    out4[1] = spo_nw_score(&wod, sub, sg, Lg, sp, Lp);
    spo_sed_symbols(gt, Lg, height, width, 5, sg);
    spo_sed_symbols(pr, Lp, height, width, 5, sp);
    out4[2] = (double)spo_levenshtein(sg, Lg, sp, Lp);
    int md = height > width ? height : width;
    out4[3] = spo_stde(gt, Lg, pr, Lp, (double)(md > 3 ? md : 3));
    free(sg); free(sp);
}

/* Batch: pair p scores human[gi[p]] against pred[pi[p]]; paths are padded
 * [*, Lmax, 3] with lengths.  Single-threaded; callers split the pair list over threads
 * (ctypes releases the GIL). */
void spo_score_pairs(const spo_cfg *cfg_wd, const double *human, const int32_t *hlen, int hLmax,
                     const double *pred, const int32_t *plen, int pLmax, const int64_t *gi, const int64_t *pi,
                     int64_t npairs, int height, int width, double *out /* [npairs,4] */) {
    int nb = cfg_wd->Xbin * cfg_wd->Ybin;
    double *sub = (double *)malloc(sizeof(double) * nb * nb);
    spo_sub_matrix(cfg_wd, sub);
    for (int64_t p = 0; p < npairs; ++p)
        spo_score_pair(cfg_wd, sub, human + gi[p] * hLmax * 3, hlen[gi[p]], pred + pi[p] * pLmax * 3, plen[pi[p]],
                       height, width, out + 4 * p);
    free(sub);
}
