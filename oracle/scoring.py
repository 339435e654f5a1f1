"""CPU oracle: ScanMatch / SED / STDE scoring (TEST INFRASTRUCTURE ONLY).

Independent numpy restatement of the reference's metric code.  The loops keep
the reference's algorithmic structure (naive per-cell / per-window Python loops)
so that timing this oracle is a fair stand-in for timing the reference.

Reference (all paths relative to /root/reference, identical in OSIE/, AiR/,
COCO_Search18/):
  * ScanMatch tables   OSIE/utils/evaltools/scanmatch.py:88-114
  * fixationToSequence OSIE/utils/evaltools/scanmatch.py:116-133
  * match (NW DP)      OSIE/utils/evaltools/scanmatch.py:135-150, 190-193
  * _scanpath_to_string / Levenshtein
                       OSIE/utils/evaltools/visual_attention_metrics.py:236-317
  * STDE               OSIE/utils/evaltools/visual_attention_metrics.py:205-218, 332-441
Pinned by tests/golden/scoring_*.npz (outputs of the reference itself).
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------
# ScanMatch
# --------------------------------------------------------------------------
class ScanMatchOracle:
    """Restates ``ScanMatch`` (scanmatch.py:39-197): ``match_score`` is what the drivers use, ``match`` also
    returns the alignment and the (transposed) F matrix like the reference's."""

    _KEYS = ("Xres", "Yres", "Xbin", "Ybin", "Threshold", "GapValue", "TempBin", "Offset")

    def __init__(self, **kw):
        self.Xres, self.Yres, self.Xbin, self.Ybin = 1024, 768, 8, 6
        self.Threshold, self.GapValue, self.TempBin, self.Offset = 3.5, 0.0, 0.0, (0, 0)
        for k, v in kw.items():
            if k not in self._KEYS:                      # scanmatch.py:80-81
                raise ValueError("Unknown parameter: %s." % k)
            setattr(self, k, v)
        self.SubMatrix = self._sub_matrix()
        self.mask = self._grid_mask()

    def _sub_matrix(self):
        # scanmatch.py:88-103. entry [r, c]: Euclid distance between the bin
        # centres of symbols r and c (symbol = row * Xbin + col), then
        # |D - Dmax| - (Dmax - Threshold).
        nb = self.Xbin * self.Ybin
        rows = np.arange(nb) // self.Xbin
        cols = np.arange(nb) % self.Xbin
        d2 = (cols[:, None] - cols[None, :]) ** 2 + (rows[:, None] - rows[None, :]) ** 2
        mat = np.sqrt(d2.astype(np.float64))
        mx = np.max(mat)
        return np.abs(mat - mx) - (mx - self.Threshold)

    def _grid_mask(self):
        # scanmatch.py:105-114: float-step aranges truncated to int32 give the
        # per-pixel bin column / row.
        a = np.arange(self.Xbin * self.Ybin).reshape(self.Ybin, self.Xbin)
        xi = np.int32(np.arange(0, self.Xbin, float(self.Xbin) / self.Xres))
        yi = np.int32(np.arange(0, self.Ybin, float(self.Ybin) / self.Yres))
        mask = np.zeros((self.Yres, self.Xres))
        for y in range(self.Yres):
            mask[y, :] = a[yi[y], xi]
        return mask

    def fixationToSequence(self, data):
        # scanmatch.py:116-133
        d = np.array(data, dtype=np.float64, copy=True)
        d[:, :2] -= self.Offset
        d[d < 0] = 0                                    # all columns, duration too
        d[d[:, 0] >= self.Xres, 0] = self.Xres - 1
        d[d[:, 1] >= self.Yres, 1] = self.Yres - 1
        di = np.trunc(d).astype(np.int64)               # int(): toward zero
        seq = self.mask[di[:, 1], di[:, 0]]
        if self.TempBin != 0:
            reps = np.round(di[:, 2] / float(self.TempBin))   # half-to-even
            out = []
            for f in range(di.shape[0]):
                out.extend([seq[f]] * int(reps[f]))
            seq = np.array(out)
        return seq

    def _fill(self, A, B):
        # scanmatch.py:135-150: borders gap * (index + 1), then the three-way maximum
        n, m = len(A), len(B)
        gap = self.GapValue
        F = np.zeros((n + 1, m + 1))
        for i in range(n + 1):
            F[i, 0] = gap * (i + 1)
        for j in range(m + 1):
            F[0, j] = gap * (j + 1)
        sub = self.SubMatrix
        for i in range(1, n + 1):
            for j in range(1, m + 1):
                diag = F[i - 1, j - 1] + sub[A[i - 1], B[j - 1]]
                up = F[i - 1, j] + gap
                left = F[i, j - 1] + gap
                F[i, j] = max(diag, left, up)
        return F

    def match_score(self, A, B):
        # scanmatch.py:135-150, 190-193 (the callers drop align and F).
        A = np.asarray(A).astype(np.int64)
        B = np.asarray(B).astype(np.int64)
        F = self._fill(A, B)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.float64(np.max(F)) / np.float64(np.max(self.SubMatrix) * max(len(B), len(A)))

    @staticmethod
    def traceback(F, A, B, sub, gap):
        """scanmatch.py:152-185, 195: walk back from F[n, m] -- a diagonal step when the cell equals the diagonal
        neighbour plus the substitution score, else a step in A when it equals F[i-1, j] + gap, else a step in
        B; the rest of either string is appended; -1 marks a gap.  Returns [steps, 2] in forward order."""
        i, j = len(A), len(B)
        ra, rb = [], []
        while i > 0 and j > 0:
            if F[i, j] == F[i - 1, j - 1] + sub[A[i - 1], B[j - 1]]:
                ra.append(A[i - 1]); rb.append(B[j - 1]); i -= 1; j -= 1
            elif F[i, j] == F[i - 1, j] + gap:
                ra.append(A[i - 1]); rb.append(-1); i -= 1
            else:
                ra.append(-1); rb.append(B[j - 1]); j -= 1
        while i > 0:
            ra.append(A[i - 1]); rb.append(-1); i -= 1
        while j > 0:
            ra.append(-1); rb.append(B[j - 1]); j -= 1
        return np.array([ra[::-1], rb[::-1]], dtype=np.float64).T.reshape(-1, 2)

    def match(self, A, B):
        A = np.asarray(A).astype(np.int64)
        B = np.asarray(B).astype(np.int64)
        F = self._fill(A, B)
        with np.errstate(invalid="ignore", divide="ignore"):
            score = np.float64(np.max(F)) / np.float64(np.max(self.SubMatrix) * max(len(B), len(A)))
        return score, self.traceback(F, A, B, self.SubMatrix, self.GapValue), F.transpose()


# --------------------------------------------------------------------------
# SED (string edit distance)
# --------------------------------------------------------------------------
def sed_symbols(scanpath, height, width, n=5):
    """visual_attention_metrics.py:288-298 as integer symbols (chr(97+sq) there)."""
    hs, ws = height // n, width // n
    sp = np.asarray(scanpath)
    out = []
    for i in range(sp.shape[0]):
        f = sp[i].astype(np.int32)
        out.append(int(f[0] // ws + (f[1] // hs) * n))
    return out


def levenshtein(s1, s2):
    """visual_attention_metrics.py:236-285, unit costs."""
    l1, l2 = len(s1), len(s2)
    D = [[0] * (l2 + 1) for _ in range(l1 + 1)]
    for i in range(l1 + 1):
        D[i][0] = i
    for j in range(l2 + 1):
        D[0][j] = j
    for i in range(1, l1 + 1):
        for j in range(1, l2 + 1):
            D[i][j] = min(D[i - 1][j] + 1, D[i][j - 1] + 1,
                          D[i - 1][j - 1] + (s1[i - 1] != s2[j - 1]))
    return D[l1][l2]


def string_edit_distance(stimulus, human_scanpath, simulated_scanpath, n=5):
    """visual_attention_metrics.py:301-317."""
    height, width = np.shape(stimulus)[0:2]
    return levenshtein(sed_symbols(human_scanpath, height, width, n),
                       sed_symbols(simulated_scanpath, height, width, n))


# --------------------------------------------------------------------------
# STDE (scaled time-delay embedding similarity)
# --------------------------------------------------------------------------
def _window_distance(hs, ss):
    # euclidean_distance (:205-218): sum over the window of point distances.
    dist = np.zeros(len(hs))
    for i in range(len(hs)):
        dist[i] = np.sqrt((hs[i][0] - ss[i][0]) ** 2 + (hs[i][1] - ss[i][1]) ** 2)
    return dist.sum()


def euclidean_distance(human_scanpath, simulated_scanpath):
    """visual_attention_metrics.py:205-218: sum of the point distances of two equally long scanpaths, else False."""
    if len(human_scanpath) != len(simulated_scanpath):
        return False
    return _window_distance(human_scanpath, simulated_scanpath)


def time_delay_embedding_distance(human, simulated, k=3, distance_mode="Mean"):
    """visual_attention_metrics.py:332-390: every k-window of `simulated` looks for its nearest k-window of
    `human` (distance / k); 'Mean' or 'Hausdorff' (max) over the simulated windows; False when k is longer than
    a scanpath or the mode is unknown."""
    if len(human) < k or len(simulated) < k:
        return False
    per_window = []
    for i in range(len(simulated) - k + 1):
        best = None
        for j in range(len(human) - k + 1):
            # reference calls euclidean_distance(s_k_vec, h_k_vec): symmetric
            d = abs(_window_distance(simulated[i:i + k], human[j:j + k]))
            best = d if best is None or d < best else best
        per_window.append(best / k)
    if distance_mode == "Mean":
        return sum(per_window) / len(per_window)
    if distance_mode == "Hausdorff":
        return max(per_window)
    return False


def scaled_time_delay_embedding_distance(human_scanpath, simulated_scanpath, image):
    """visual_attention_metrics.py:444-492: mean over k of the 'Mean' distances of the rescaled scanpaths."""
    H = np.array(human_scanpath, dtype=np.float64, copy=True)
    S = np.array(simulated_scanpath, dtype=np.float64, copy=True)
    max_dim = float(max(np.shape(image)))
    H[:, :2] /= max_dim
    S[:, :2] /= max_dim
    d = [time_delay_embedding_distance(H, S, k) for k in range(1, min(len(H), len(S)) + 1)]
    return sum(d) / len(d) if d else None


def scaled_time_delay_embedding_similarity(human_scanpath, simulated_scanpath, image):
    """visual_attention_metrics.py:393-441.  Asymmetric: windows of `simulated`
    look for their nearest `human` window."""
    H = np.array(human_scanpath, dtype=np.float64, copy=True)
    S = np.array(simulated_scanpath, dtype=np.float64, copy=True)
    max_dim = float(max(np.shape(image)))
    H[:, 0] /= max_dim
    H[:, 1] /= max_dim
    S[:, 0] /= max_dim
    S[:, 1] /= max_dim
    sims = []
    for k in range(1, min(len(H), len(S)) + 1):
        sims.append(np.exp(-time_delay_embedding_distance(H, S, k)))
    if len(sims) == 0:
        return None
    return sum(sims) / len(sims)


# --------------------------------------------------------------------------
# pair scoring used by the evaluation drivers
# --------------------------------------------------------------------------
EVAL_CFG = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
EVAL_TEMPBIN = 50
STIMULUS_SHAPE = (240, 320, 3)

_cached = {}


def eval_scanmatch_objects():
    if "wd" not in _cached:
        _cached["wd"] = ScanMatchOracle(TempBin=EVAL_TEMPBIN, **EVAL_CFG)
        _cached["wod"] = ScanMatchOracle(**EVAL_CFG)
    return _cached["wd"], _cached["wod"]


def structured_to_array(fix_vector):
    """``np.array([list(_) for _ in list(fix_vector)])`` (OSIE/utils/evaluation.py:180)
    followed by the s -> ms scaling (:182)."""
    if len(fix_vector) == 0:
        return np.zeros((0, 3))
    arr = np.array([[float(v) for v in row] for row in list(fix_vector)], dtype=np.float64)
    arr[:, -1] *= 1000
    return arr


def score_pair(gt_xyms, pred_xyms, sm_wd=None, sm_wod=None):
    """The four in-scope numbers of one (human, prediction) pair in the order
    the drivers compute them (OSIE/utils/evaluation.py:177-204):
    (ScanMatch-wd, ScanMatch-wod, SED, STDE).  Inputs are [L,3] (x, y, ms)."""
    if sm_wd is None:
        sm_wd, sm_wod = eval_scanmatch_objects()
    stimulus = np.zeros(STIMULUS_SHAPE, dtype=np.float32)
    s1 = sm_wd.fixationToSequence(gt_xyms).astype(np.int32)
    s2 = sm_wd.fixationToSequence(pred_xyms).astype(np.int32)
    wd = sm_wd.match_score(s1, s2)
    s1 = sm_wod.fixationToSequence(gt_xyms).astype(np.int32)
    s2 = sm_wod.fixationToSequence(pred_xyms).astype(np.int32)
    wod = sm_wod.match_score(s1, s2)
    sed = string_edit_distance(stimulus, gt_xyms, pred_xyms)
    stde = scaled_time_delay_embedding_similarity(gt_xyms, pred_xyms, stimulus)
    return wd, wod, sed, stde


def multimatch_valid(gt_len, pred_len, min_len_valid=3):
    """MultiMatch (multimatch-gaze 0.1.2, un-vendored, parity unpinned) yields
    NaN iff either scanpath has fewer than 3 fixations; only this rule touches
    in-scope numbers (OSIE/utils/evaluation.py:296-299, 326-327)."""
    return gt_len >= min_len_valid and pred_len >= min_len_valid


def evaluation(gt_fix_vectors, predict_fix_vectors):
    """OSIE/utils/evaluation.py:151-282 minus the MultiMatch numbers.

    gt_fix_vectors[i]: list of S structured arrays (start_x, start_y, duration[s]);
    predict_fix_vectors[i]: one structured array.  Returns (metrics, metrics_std,
    per_image) with per_image rows = mean over S of (wd, wod, sed, stde)."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    wd, wod, sed, stde, per_image = [], [], [], [], []
    last_S = 0
    for gts, pred in zip(gt_fix_vectors, predict_fix_vectors):
        p = structured_to_array(pred)
        rows = []
        for g in gts:
            r = score_pair(structured_to_array(g), p, sm_wd, sm_wod)
            wd.append(r[0]); wod.append(r[1]); sed.append(r[2]); stde.append(r[3])
            rows.append(r)
        per_image.append(list(np.array(rows, dtype=np.float64).mean(axis=0)))
        last_S = len(gts)
    sed_t = np.array(sed).reshape(-1, last_S)            # :224 (assumes constant S)
    stde_t = np.array(stde).reshape(-1, last_S)
    m = {"ScanMatch": {"w/o duration": np.mean(wod), "with duration": np.mean(wd)},
         "VAME": {"SED": sed_t.mean(), "STDE": stde_t.mean(),
                  "SED_best": sed_t.min(-1).mean(), "STDE_best": stde_t.max(-1).mean()}}
    s = {"ScanMatch": {"w/o duration": np.std(wod), "with duration": np.std(wd)},
         "VAME": {"SED": sed_t.std(), "STDE": stde_t.std(),
                  "SED_best": sed_t.min(-1).std(), "STDE_best": stde_t.max(-1).std()}}
    return m, s, per_image


def human_evaluation(batches_fix_vectors, per_image_best=False):
    """OSIE/utils/evaluation.py:11-148 minus MultiMatch: ordered pairs (i, j!=i)
    inside each image; first scanpath plays 'human', second 'simulated'.
    `batches_fix_vectors`: list over images of list over subjects.
    per_image_best=True is the COCO-Search18 variant (COCO_Search18/utils/evaluation.py:88-125):
    the number of subjects may differ per image and SED_best / STDE_best are the min / max over
    ALL ordered pairs of an image (OSIE: per first subject, with a constant subject count)."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    wd, wod, sed, stde, per_image, best = [], [], [], [], [], []
    last_S = 0
    for fvs in batches_fix_vectors:
        arrs = [structured_to_array(f) for f in fvs]
        rows = []
        for i in range(len(arrs)):
            for j in range(len(arrs)):
                if i == j:
                    continue
                r = score_pair(arrs[i], arrs[j], sm_wd, sm_wod)
                wd.append(r[0]); wod.append(r[1]); sed.append(r[2]); stde.append(r[3])
                rows.append(r)
        rows = np.array(rows, dtype=np.float64)
        per_image.append(list(rows.mean(axis=0)))
        best.append((rows[:, 2].min(), rows[:, 3].max()))
        last_S = len(fvs)
    if per_image_best:
        sed_best, stde_best = np.array([b[0] for b in best]), np.array([b[1] for b in best])
    else:
        sed_best = np.array(sed).reshape(-1, last_S - 1).min(-1)
        stde_best = np.array(stde).reshape(-1, last_S - 1).max(-1)
    m = {"ScanMatch": {"w/o duration": np.mean(wod), "with duration": np.mean(wd)},
         "VAME": {"SED": np.mean(sed), "STDE": np.mean(stde), "SED_best": sed_best.mean(), "STDE_best": stde_best.mean()}}
    s = {"ScanMatch": {"w/o duration": np.std(wod), "with duration": np.std(wd)},
         "VAME": {"SED": np.std(sed), "STDE": np.std(stde), "SED_best": sed_best.std(), "STDE_best": stde_best.std()}}
    return m, s, per_image


def pairs_eval(gt_fix_vectors, predict_fix_vectors, is_eliminating_nan=True, min_len_valid=3):
    """OSIE/utils/evaluation.py:284-340.  Returns [N, 11] float; the five
    MultiMatch slots [0:5] are NaN (out of scope) but MultiMatch's NaN rule
    still drops rows.  Slots: 5 = SM w/o duration, 6 = SM with duration,
    7 = SED mean, 8 = STDE mean, 9 = SED best(min), 10 = STDE best(max);
    means divide by len(gt) even when rows were dropped (:329)."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    out = []
    for gts, pred in zip(gt_fix_vectors, predict_fix_vectors):
        p = structured_to_array(pred)
        rows = []
        for g in gts:
            if not multimatch_valid(len(g), len(pred), min_len_valid):
                if not is_eliminating_nan:
                    rows.append([np.nan] * 4)
                continue
            wd, wod, sed, stde = score_pair(structured_to_array(g), p, sm_wd, sm_wod)
            rows.append([wod, wd, sed, stde])
        rows = np.array(rows, dtype=np.float64).reshape(-1, 4)
        v = np.full((11,), np.nan, dtype=np.float64)
        if rows.shape[0] != 0:
            mean = rows.sum(axis=0) / len(gts)
            v[5:9] = mean
            v[9] = rows[:, 2].min()
            v[10] = rows[:, 3].max()
            v = v.astype(np.float32).astype(np.float64)   # metric_value is float32 (:330)
        out.append(v)
    return np.array(out)


def pairs_eval_scanmatch(gt_fix_vectors, predict_fix_vectors):
    """COCO_Search18/utils/evaluation.py:313-352: two numbers per image
    (SM w/o duration, SM with duration), mean over subjects, no MultiMatch."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    out = []
    for gts, pred in zip(gt_fix_vectors, predict_fix_vectors):
        p = structured_to_array(pred)
        rows = []
        for g in gts:
            a = structured_to_array(g)
            wd = sm_wd.match_score(sm_wd.fixationToSequence(a).astype(np.int32),
                                   sm_wd.fixationToSequence(p).astype(np.int32))
            wod = sm_wod.match_score(sm_wod.fixationToSequence(a).astype(np.int32),
                                     sm_wod.fixationToSequence(p).astype(np.int32))
            rows.append([wod, wd])
        rows = np.array(rows, dtype=np.float64).reshape(-1, 2)
        rows = rows[~np.isnan(rows.sum(axis=1))]               # NaN rows (both wd strings empty) dropped
        out.append(rows.sum(axis=0) / len(gts) if rows.shape[0] else np.array([np.nan] * 2))
    return np.array(out)


# --------------------------------------------------------------------------
# AiR: performance-related drivers (AiR/utils/evaluation.py:188-577)
# --------------------------------------------------------------------------
def _sm_pair(a, b, sm_wd, sm_wod):
    wd = sm_wd.match_score(sm_wd.fixationToSequence(a).astype(np.int32), sm_wd.fixationToSequence(b).astype(np.int32))
    wod = sm_wod.match_score(sm_wod.fixationToSequence(a).astype(np.int32),
                             sm_wod.fixationToSequence(b).astype(np.int32))
    return [wod, wd]


def _mean_rows(rows, drop_nan=True):
    rows = np.array(rows, dtype=np.float64).reshape(-1, 2)
    flag = True
    if drop_nan and rows.shape[0] != 0:
        rows = rows[~np.isnan(rows.sum(axis=1))]
        flag = rows.shape[0] != 0
    return (rows.sum(axis=0) / rows.shape[0] if rows.shape[0] else np.array([np.nan] * 2)), flag


def pairs_eval_scanmatch_performance_related(gt_fix_vectors, predict_fix_vectors, performance, given_performance,
                                             is_eliminating_nan=True):
    """:361-420 -> (same [N,2], diff [N,2], accept_flag): subjects whose performance equals
    `given_performance` vs the others; columns (SM w/o duration, SM with duration)."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    same, diff, accept = [], [], True
    for i, (gts, pred) in enumerate(zip(gt_fix_vectors, predict_fix_vectors)):
        p = structured_to_array(pred)
        s_rows, d_rows = [], []
        for j, g in enumerate(gts):
            (s_rows if performance[i][j] == given_performance else d_rows).append(
                _sm_pair(structured_to_array(g), p, sm_wd, sm_wod))
        m, f1 = _mean_rows(s_rows, is_eliminating_nan)
        same.append(m)
        m, f2 = _mean_rows(d_rows, is_eliminating_nan)
        diff.append(m)
        accept = accept and f1 and f2
    return np.array(same), np.array(diff), accept


def gtpairs_eval_scanmatch_performance_related(gt_fix_vectors, performance, is_eliminating_nan=True):
    """:423-577 -> (good-vs-good, poor-vs-poor, good-vs-poor) [N,2] human-human ScanMatch means."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    good_m, poor_m, gp_m = [], [], []
    for gts, perf in zip(gt_fix_vectors, performance):
        arrs = [structured_to_array(g) for g in gts]
        good = [a for a, p in zip(arrs, perf) if p == True]     # noqa: E712 (the reference compares with ==)
        poor = [a for a, p in zip(arrs, perf) if not (p == True)]  # noqa: E712
        for grp, dst in ((good, good_m), (poor, poor_m)):
            rows = []
            if len(grp) > 1:
                for a in range(len(grp)):
                    for b in range(a + 1, len(grp)):
                        rows.append(_sm_pair(grp[a], grp[b], sm_wd, sm_wod))
            dst.append(_mean_rows(rows, is_eliminating_nan)[0])
        rows = []
        if len(good) > 1 and len(poor) > 1:
            for a in good:
                for b in poor:
                    rows.append(_sm_pair(a, b, sm_wd, sm_wod))
        gp_m.append(_mean_rows(rows, is_eliminating_nan)[0])
    return np.array(good_m), np.array(poor_m), np.array(gp_m)


def evaluation_performance_related(gt_fix_vectors, predict_fix_vectors, all_performances,
                                   all_allocated_performances, min_len_valid=3):
    """:188-359 minus MultiMatch.  Rows are stored as float32 there, so means / stds are float32.
    Returns (mean [3,6], std [3,6], per_image [N,4]) for the categories (all, right_answer,
    wrong_answer); the six columns are specific_mean[5:11] = (SM WITH duration, SM w/o duration,
    SED, STDE, SED_best, STDE_best) -- which AiR then labels 'w/o duration' / 'with duration'
    the other way round (:321-322)."""
    sm_wd, sm_wod = eval_scanmatch_objects()
    cats = [[], [], []]
    per_image = []
    for i, (gts, pred) in enumerate(zip(gt_fix_vectors, predict_fix_vectors)):
        p = structured_to_array(pred)
        rows = [[], [], []]
        for j, g in enumerate(gts):
            if not multimatch_valid(len(g), len(pred), min_len_valid):
                continue
            wd, wod, sed, stde = score_pair(structured_to_array(g), p, sm_wd, sm_wod)
            r = [wd, wod, sed, stde]
            rows[0].append(r)
            if all_performances[i][j] == True and all_allocated_performances[i] == True:      # noqa: E712
                rows[1].append(r)
            elif all_performances[i][j] == False and all_allocated_performances[i] == False:  # noqa: E712
                rows[2].append(r)
        for c in range(3):
            cats[c].append(np.array(rows[c], dtype=np.float32).reshape(-1, 4))
        own = rows[1] if all_allocated_performances[i] == True else rows[2]                     # noqa: E712
        per_image.append(list(np.array(own).mean(axis=0)) if own else [0.0] * 4)
    mean, std = [], []
    for c in range(3):
        groups = [g for g in cats[c] if len(g) != 0]
        allrows = np.concatenate(groups, axis=0)
        best = np.array([[g[:, 2].min(), g[:, 3].max()] for g in groups], dtype=np.float32)
        mean.append(np.concatenate([allrows.mean(0), best.mean(0)]))
        std.append(np.concatenate([allrows.std(0), best.std(0)]))
    return np.array(mean), np.array(std), np.array(per_image, dtype=np.float64)
