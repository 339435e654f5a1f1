"""Drop-in for the reference's evaluation drivers on the GPU.

Mirrors (names, arguments, return types):
  evaluation, human_evaluation, pairs_eval   OSIE/utils/evaluation.py:11-340
  pairs_eval_scanmatch                       COCO_Search18/utils/evaluation.py:313-352
Every (human, prediction) pair of the call is scored by ONE launch of
csrc/score_pairs.cu instead of the reference's triple Python loop.  MultiMatch
(external multimatch-gaze 0.1.2, not part of this path) is not computed: its
five slots are NaN, but its NaN rule (either scanpath shorter than 3 fixations)
still drops rows exactly where the reference's aggregation drops them.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import scoring as S

MIN_LEN_VALID = 3   # multimatch_gaze.docomparison returns NaN below this

_cfg = {}


def _eval_cfg(device=None):
    key = str(device)
    if key not in _cfg:
        _cfg[key] = S.ScoreConfig.evaluation(device=device, dur_scale=1000.0)
    return _cfg[key]


def _nan5():
    return {"vector": np.nan, "direction": np.nan, "length": np.nan, "position": np.nan, "duration": np.nan}


def _pack_humans(gt_fix_vectors, cfg):
    """Packs each distinct subject list once (test.py:125 re-appends the same list
    objects for every trial).  Returns (pack, first human index per entry, S per entry)."""
    seen, paths, base, sizes = {}, [], [], []
    for gts in gt_fix_vectors:
        key = id(gts)
        if key not in seen:
            seen[key] = len(paths)
            paths.extend(S.structured_to_xyd(g) for g in gts)
        base.append(seen[key]); sizes.append(len(gts))
    return S.pack_paths(paths, cfg), np.array(base, dtype=np.int64), np.array(sizes, dtype=np.int64)


class _Scored:
    """One scored call: device score table + both packs + the pair map (host and device copies)."""

    def __init__(self, scores, hpack, ppack, pair_h, pair_s, sizes, d_pair_h, d_pair_s):
        self.scores, self.hpack, self.ppack = scores, hpack, ppack
        self.pair_h, self.pair_s, self.sizes = pair_h, pair_s, sizes
        self.d_pair_h, self.d_pair_s = d_pair_h, d_pair_s

    def __iter__(self):          # (scores, hpack, ppack, pair_h, pair_s, sizes)
        return iter((self.scores, self.hpack, self.ppack, self.pair_h, self.pair_s, self.sizes))


def _score_lists(gt_fix_vectors, predict_fix_vectors, device=None):
    cfg = _eval_cfg(device)
    preds = []
    for p in predict_fix_vectors:
        a = S.structured_to_xyd(p)
        if len(a) == 0:
            raise IndexError("too many indices for array: empty predicted scanpath")   # evaluation.py:181-183
        preds.append(a)
    hpack, base, sizes = _pack_humans(gt_fix_vectors, cfg)
    ppack = S.pack_paths(preds, cfg)
    pair_s = np.repeat(np.arange(len(preds), dtype=np.int64), sizes)
    offs = np.concatenate([np.arange(s) for s in sizes]) if len(sizes) else np.zeros(0, np.int64)
    pair_h = np.repeat(base, sizes) + offs
    dev = cfg.device
    d_ph = torch.from_numpy(pair_h.astype(np.int32)).to(dev)
    d_ps = torch.from_numpy(pair_s.astype(np.int32)).to(dev)
    scores = S.score_pairs(hpack, ppack, d_ph, d_ps, cfg)
    return _Scored(scores, hpack, ppack, pair_h, pair_s, sizes, d_ph, d_ps)


def _metric_dicts(scores, last_group):
    """Aggregation of evaluation.py:211-280 from the [P,4] table (device, f64)."""
    wd, wod = scores[:, 0], scores[:, 1]
    sed = scores[:, 2].reshape(-1, last_group)
    stde = scores[:, 3].reshape(-1, last_group)
    f = lambda t: float(t.item())
    std = lambda t: f(t.std(unbiased=False))
    m = {"MultiMatch": _nan5(),
         "ScanMatch": {"w/o duration": f(wod.mean()), "with duration": f(wd.mean())},
         "VAME": {"SED": f(sed.mean()), "STDE": f(stde.mean()),
                  "SED_best": f(sed.min(-1)[0].mean()), "STDE_best": f(stde.max(-1)[0].mean())}}
    s = {"MultiMatch": _nan5(),
         "ScanMatch": {"w/o duration": std(wod), "with duration": std(wd)},
         "VAME": {"SED": std(sed), "STDE": std(stde),
                  "SED_best": std(sed.min(-1)[0]), "STDE_best": std(stde.max(-1)[0])}}
    return m, s


def _per_group_means(scores, sizes):
    sc = scores.cpu().numpy()
    out, o = [], 0
    for s in sizes:
        out.append([np.nan] * 5 + list(sc[o:o + s].mean(axis=0)))
        o += s
    return out


def evaluation(gt_fix_vectors, predict_fix_vectors, is_eliminating_nan=True):
    """OSIE/utils/evaluation.py:151-282."""
    scores, _, _, _, _, sizes = _score_lists(gt_fix_vectors, predict_fix_vectors)
    m, s = _metric_dicts(scores, int(sizes[-1]))
    return m, s, _per_group_means(scores, sizes)


def human_evaluation(dataloader, per_image_best=False):
    """OSIE/utils/evaluation.py:11-148: all ordered pairs (i, j != i) of each image's
    subjects; i plays the human, j the simulated scanpath.  per_image_best=True is the
    COCO-Search18 variant (COCO_Search18/utils/evaluation.py:88-125): subject counts may differ per
    image and SED_best / STDE_best are taken over all ordered pairs of an image."""
    cfg = _eval_cfg()
    paths, pair_h, pair_s, sizes, names = [], [], [], [], []
    last_S = 0
    for batch in dataloader:
        names.extend(batch["img_names"])
        for fix_vectors in batch["fix_vectors"]:
            b = len(paths)
            paths.extend(S.structured_to_xyd(f) for f in fix_vectors)
            n = len(fix_vectors)
            for i in range(n):
                for j in range(n):
                    if i != j:
                        pair_h.append(b + i); pair_s.append(b + j)
            sizes.append(n * (n - 1)); last_S = n
    pack = S.pack_paths(paths, cfg)
    dev = cfg.device
    scores = S.score_pairs(pack, pack, torch.tensor(pair_h, dtype=torch.int32, device=dev),
                           torch.tensor(pair_s, dtype=torch.int32, device=dev), cfg)
    if per_image_best:
        sc = scores.cpu().numpy()
        off = np.concatenate([[0], np.cumsum(sizes)])
        sed_b = np.array([sc[off[i]:off[i + 1], 2].min() for i in range(len(sizes))])
        stde_b = np.array([sc[off[i]:off[i + 1], 3].max() for i in range(len(sizes))])
        m = {"MultiMatch": _nan5(),
             "ScanMatch": {"w/o duration": sc[:, 1].mean(), "with duration": sc[:, 0].mean()},
             "VAME": {"SED": sc[:, 2].mean(), "STDE": sc[:, 3].mean(), "SED_best": sed_b.mean(),
                      "STDE_best": stde_b.mean()}}
        s = {"MultiMatch": _nan5(),
             "ScanMatch": {"w/o duration": sc[:, 1].std(), "with duration": sc[:, 0].std()},
             "VAME": {"SED": sc[:, 2].std(), "STDE": sc[:, 3].std(), "SED_best": sed_b.std(), "STDE_best": stde_b.std()}}
    else:
        m, s = _metric_dicts(scores, last_S - 1)
    per = _per_group_means(scores, sizes)
    return m, s, {name: sc_ for name, sc_ in zip(names, per)}


def human_evaluation_packed(xyd, lens, n_subjects=None, per_image_best=False, img_names=None):
    """human_evaluation on the packed layout (scanpaths_b200.dataset / scoring.pack_subject_lists):
    xyd [N,S,Lmax,3] f64 (seconds), lens [N,S] i32, n_subjects [N] or None (= S everywhere).  Same pair
    order, same aggregates and the same return value as ``human_evaluation`` -- without building Python lists
    of structured arrays or a Python-level pair loop (the pair map is one vectorised numpy expression)."""
    cfg = _eval_cfg()
    dev = cfg.device
    xyd = torch.as_tensor(xyd, dtype=torch.float64)
    lens = torch.as_tensor(lens, dtype=torch.int32)
    N, Smax, L, _ = xyd.shape
    nsub = np.full(N, Smax, dtype=np.int64) if n_subjects is None else np.asarray(torch.as_tensor(n_subjects).cpu(), dtype=np.int64)
    i, j = np.nonzero(~np.eye(Smax, dtype=bool))                    # ordered pairs (i, j != i), i-major like :30-33
    keep = (i[None, :] < nsub[:, None]) & (j[None, :] < nsub[:, None])
    base = (np.arange(N, dtype=np.int64) * Smax)[:, None]
    pair_h = (base + i[None, :])[keep].astype(np.int32)
    pair_s = (base + j[None, :])[keep].astype(np.int32)
    sizes = nsub * (nsub - 1)
    pack = S.prep_paths(xyd.reshape(N * Smax, L, 3).to(dev), lens.reshape(-1).to(dev), cfg)
    scores = S.score_pairs(pack, pack, torch.from_numpy(pair_h).to(dev), torch.from_numpy(pair_s).to(dev), cfg)
    if per_image_best or not np.all(nsub == nsub[-1]):
        sc = scores.cpu().numpy()
        off = np.concatenate([[0], np.cumsum(sizes)])
        sed_b = np.array([sc[off[k]:off[k + 1], 2].min() for k in range(N) if sizes[k]])
        stde_b = np.array([sc[off[k]:off[k + 1], 3].max() for k in range(N) if sizes[k]])
        m = {"MultiMatch": _nan5(),
             "ScanMatch": {"w/o duration": sc[:, 1].mean(), "with duration": sc[:, 0].mean()},
             "VAME": {"SED": sc[:, 2].mean(), "STDE": sc[:, 3].mean(), "SED_best": sed_b.mean(),
                      "STDE_best": stde_b.mean()}}
        s = {"MultiMatch": _nan5(),
             "ScanMatch": {"w/o duration": sc[:, 1].std(), "with duration": sc[:, 0].std()},
             "VAME": {"SED": sc[:, 2].std(), "STDE": sc[:, 3].std(), "SED_best": sed_b.std(), "STDE_best": stde_b.std()}}
    else:
        m, s = _metric_dicts(scores, int(nsub[-1]) - 1)
    per = _per_group_means(scores, sizes)
    names = img_names if img_names is not None else list(range(N))
    return m, s, {name: sc_ for name, sc_ in zip(names, per)}


def pairs_eval(gt_fix_vectors, predict_fix_vectors, ScanMatchwithDuration=None, ScanMatchwithoutDuration=None,
               is_eliminating_nan=True):
    """OSIE/utils/evaluation.py:284-340 -> [N, 11] (float32 values).  The two ScanMatch
    arguments are accepted for signature compatibility; the drivers' fixed
    configuration (train.py:201-203) is used.  Slots 0..4 (MultiMatch, out of scope) hold the
    placeholder 0.0 wherever the reference returns numbers, and the row is all-NaN exactly where the
    reference's is (no pair of the image survives the < 3 fixations rule) -- so train.py:237's
    `np.any(np.isnan(metrics_reward))` rejects the same trials as with the reference."""
    scored = _score_lists(gt_fix_vectors, predict_fix_vectors)
    scores, hpack, ppack, pair_h, pair_s, sizes = scored
    S_ = int(sizes[0]) if len(sizes) else 0
    if len(sizes) and not np.all(sizes == S_):
        return _pairs_eval_ragged(scores, hpack, ppack, pair_h, pair_s, sizes, is_eliminating_nan)
    table, _, _ = S.reduce_pairs(scores, S_, pair_h=scored.d_pair_h, pair_s=scored.d_pair_s, len_h=hpack.len,
                                 len_s=ppack.len,
                                 min_len_valid=MIN_LEN_VALID)
    out = table.cpu().numpy().astype(np.float64)
    if not is_eliminating_nan:      # without elimination one NaN row poisons the image's sums (:326-335)
        hl, pl = hpack.len.cpu().numpy(), ppack.len.cpu().numpy()
        bad = ((hl[pair_h] < MIN_LEN_VALID) | (pl[pair_s] < MIN_LEN_VALID)).reshape(-1, S_).any(1)
        out[bad] = np.nan
    return out


def _pairs_eval_ragged(scores, hpack, ppack, pair_h, pair_s, sizes, is_eliminating_nan):
    sc = scores.cpu().numpy()
    hl, pl = hpack.len.cpu().numpy(), ppack.len.cpu().numpy()
    out, o = [], 0
    for s in sizes:
        rows = sc[o:o + s]
        ok = (hl[pair_h[o:o + s]] >= MIN_LEN_VALID) & (pl[pair_s[o:o + s]] >= MIN_LEN_VALID) & ~np.isnan(rows.sum(1))
        v = np.full(11, np.nan)
        if ok.any() and (is_eliminating_nan or ok.all()):
            r = rows[ok]
            v[:5] = 0.0
            v[5], v[6], v[7], v[8] = (r[:, 1].sum() / s, r[:, 0].sum() / s, r[:, 2].sum() / s, r[:, 3].sum() / s)
            v[9], v[10] = r[:, 2].min(), r[:, 3].max()
            v = v.astype(np.float32).astype(np.float64)
        out.append(v); o += s
    return np.array(out)


def pairs_eval_scanmatch(gt_fix_vectors, predict_fix_vectors, ScanMatchwithDuration=None,
                         ScanMatchwithoutDuration=None, is_eliminating_nan=True):
    """COCO_Search18/utils/evaluation.py:313-352 -> [N, 2] = (SM w/o duration, SM with duration)."""
    scores, _, _, _, _, sizes = _score_lists(gt_fix_vectors, predict_fix_vectors)
    sc = scores[:, :2].cpu().numpy()
    out, o = [], 0
    for s in sizes:
        rows = sc[o:o + s][:, ::-1]                       # (wod, wd)
        if is_eliminating_nan:
            rows = rows[~np.isnan(rows.sum(axis=1))]
        out.append(rows.sum(axis=0) / s if rows.shape[0] else np.array([np.nan] * 2))
        o += s
    return np.array(out)


# ---------------------------------------------------------------------------------------
# AiR: performance-related drivers (AiR/utils/evaluation.py:188-577).  Same kernels, other
# pair maps / groupings; the grouping itself is O(pairs) host bookkeeping on the score table.
# ---------------------------------------------------------------------------------------
def _score_index_pairs(paths_h, paths_s, pair_h, pair_s, device=None):
    """Scores explicit (human index, simulated index) pairs of two lists of [L,3] arrays (seconds)."""
    cfg = _eval_cfg(device)
    if len(pair_h) == 0:
        return np.zeros((0, 4))
    hp = S.pack_paths(paths_h, cfg)
    sp = hp if paths_s is paths_h else S.pack_paths(paths_s, cfg)
    dev = cfg.device
    sc = S.score_pairs(hp, sp, torch.tensor(pair_h, dtype=torch.int32, device=dev),
                       torch.tensor(pair_s, dtype=torch.int32, device=dev), cfg)
    return sc.cpu().numpy()


def _mean2(rows, drop_nan):
    """rows [n,2] (wod, wd) -> (mean or NaN, group still non-empty after NaN elimination)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 2)
    ok = True
    if drop_nan and rows.shape[0] != 0:
        rows = rows[~np.isnan(rows.sum(axis=1))]
        ok = rows.shape[0] != 0
    return (rows.sum(axis=0) / rows.shape[0] if rows.shape[0] else np.array([np.nan] * 2)), ok


def pairs_eval_scanmatch_performance_related(gt_fix_vectors, predict_fix_vectors, ScanMatchwithDuration=None,
                                             ScanMatchwithoutDuration=None, performance=None, given_performance=True,
                                             is_eliminating_nan=True):
    """AiR/utils/evaluation.py:361-420 -> (same [N,2], diff [N,2], accept_flag)."""
    scores, _, _, _, _, sizes = _score_lists(gt_fix_vectors, predict_fix_vectors)
    sc = scores[:, :2].cpu().numpy()[:, ::-1]              # (wod, wd)
    same, diff, accept, o = [], [], True, 0
    for i, s in enumerate(sizes):
        rows = sc[o:o + s]
        mask = np.array([performance[i][j] == given_performance for j in range(s)], dtype=bool)
        m, f1 = _mean2(rows[mask], is_eliminating_nan)
        same.append(m)
        m, f2 = _mean2(rows[~mask], is_eliminating_nan)
        diff.append(m)
        accept = accept and f1 and f2
        o += s
    return np.array(same), np.array(diff), accept


def gtpairs_eval_scanmatch_performance_related(gt_fix_vectors, ScanMatchwithDuration=None,
                                               ScanMatchwithoutDuration=None, performance=None,
                                               is_eliminating_nan=True):
    """AiR/utils/evaluation.py:423-577 -> (good v good, poor v poor, good v poor) [N,2]: human-human
    ScanMatch means inside each image, grouped by answer correctness.  One launch for all images."""
    paths, pair_h, pair_s, spans = [], [], [], []
    for gts, perf in zip(gt_fix_vectors, performance):
        b = len(paths)
        paths.extend(S.structured_to_xyd(g) for g in gts)
        good = [b + j for j in range(len(gts)) if perf[j] == True]        # noqa: E712
        poor = [b + j for j in range(len(gts)) if not (perf[j] == True)]  # noqa: E712
        span = []
        for grp in (good, poor):
            lo = len(pair_h)
            if len(grp) > 1:
                for a in range(len(grp)):
                    for c in range(a + 1, len(grp)):
                        pair_h.append(grp[a]); pair_s.append(grp[c])
            span.append((lo, len(pair_h)))
        lo = len(pair_h)
        if len(good) > 1 and len(poor) > 1:
            for a in good:
                for c in poor:
                    pair_h.append(a); pair_s.append(c)
        span.append((lo, len(pair_h)))
        spans.append(span)
    sc = _score_index_pairs(paths, paths, pair_h, pair_s)[:, :2][:, ::-1]
    out = [[], [], []]
    for span in spans:
        for k, (lo, hi) in enumerate(span):
            out[k].append(_mean2(sc[lo:hi], is_eliminating_nan)[0])
    return np.array(out[0]), np.array(out[1]), np.array(out[2])


def evaluation_performance_related(gt_fix_vectors, predict_fix_vectors, all_performances, all_allocated_performances):
    """AiR/utils/evaluation.py:188-359.  Returns (metrics, metrics_std, per_image) with the categories
    all / right_answer / wrong_answer.  Quirks kept: rows are aggregated in float32; AiR stores
    (SM with duration, SM w/o duration) in slots 5, 6 but LABELS slot 5 'w/o duration' and slot 6
    'with duration' (:231-236 vs :321-322) -- the labels below are the reference's."""
    scores, hpack, ppack, pair_h, pair_s, sizes = _score_lists(gt_fix_vectors, predict_fix_vectors)
    sc = scores.cpu().numpy()
    hl, pl = hpack.len.cpu().numpy(), ppack.len.cpu().numpy()
    cats, per_image, o = [[], [], []], [], 0
    for i, s in enumerate(sizes):
        rows = [[], [], []]
        for j in range(s):
            if hl[pair_h[o + j]] < MIN_LEN_VALID or pl[pair_s[o + j]] < MIN_LEN_VALID:
                continue                                   # MultiMatch NaN -> pair skipped (:218-219)
            r = sc[o + j]
            rows[0].append(r)
            if all_performances[i][j] == True and all_allocated_performances[i] == True:      # noqa: E712
                rows[1].append(r)
            elif all_performances[i][j] == False and all_allocated_performances[i] == False:  # noqa: E712
                rows[2].append(r)
        for c in range(3):
            cats[c].append(np.array(rows[c], dtype=np.float32).reshape(-1, 4))
        own = rows[1] if all_allocated_performances[i] == True else rows[2]                     # noqa: E712
        per_image.append([np.nan] * 5 + list(np.array(own).mean(axis=0)) if own else list(np.zeros(9)))
        o += s
    m_out, s_out = {}, {}
    for c, name in enumerate(("all", "right_answer", "wrong_answer")):
        groups = [g for g in cats[c] if len(g) != 0]
        allrows = np.concatenate(groups, axis=0)
        best = np.array([[g[:, 2].min(), g[:, 3].max()] for g in groups], dtype=np.float32)
        for dst, fn in ((m_out, np.mean), (s_out, np.std)):
            a, b = fn(allrows, axis=0), fn(best, axis=0)
            dst[name] = {"MultiMatch": _nan5(),
                         "ScanMatch": {"w/o duration": a[0], "with duration": a[1]},
                         "VAME": {"SED": a[2], "STDE": a[3], "SED_best": b[0], "STDE_best": b[1]}}
    return m_out, s_out, per_image
