"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference through tests/refload.py).

Run in the authoring container only:   python tests/golden/make_goldens.py [scoring|eval|sampling|decoder|all]

The fixtures are small .npz files holding seeded inputs and the reference's
outputs; the GPU box (which has no /root/reference) checks the oracle and the
CUDA path against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refload  # noqa: E402

FIX_DTYPE = np.dtype({"names": ("start_x", "start_y", "duration"), "formats": ("f8", "f8", "f8")})


def to_struct(arr):
    out = np.zeros(len(arr), dtype=FIX_DTYPE)
    if len(arr):
        out["start_x"], out["start_y"], out["duration"] = arr[:, 0], arr[:, 1], arr[:, 2]
    return out


def pad(list_of_arr, lmax=None):
    lmax = lmax or max(1, max(len(a) for a in list_of_arr))
    out = np.zeros((len(list_of_arr), lmax, 3), dtype=np.float64)
    lens = np.zeros(len(list_of_arr), dtype=np.int32)
    for i, a in enumerate(list_of_arr):
        out[i, :len(a)] = a
        lens[i] = len(a)
    return out, lens


def human_paths(rng, n, lo=6, hi=14):
    """SURVEY.md section 8d synthetic human scanpaths: [L,3] (x, y, seconds)."""
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        x = rng.uniform(0, 320, L)
        y = rng.uniform(0, 240, L)
        d = np.exp(rng.normal(np.log(0.25), 0.4, L))
        out.append(np.stack([x, y, d], 1))
    return out


def pred_paths(rng, n, lo=1, hi=16, dur_sigma=0.4):
    """Prediction-like scanpaths: grid-cell centres (sampling.py:64-68) and
    float32 durations widened to f64 (sampling.py:73-74)."""
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        cell = rng.integers(0, 1200, L)
        x = (cell % 40) * 8.0 + 4.0
        y = (cell // 40) * 8.0 + 4.0
        d = np.exp(rng.normal(np.log(0.25), dur_sigma, L)).astype(np.float32).astype(np.float64)
        out.append(np.stack([x, y, d], 1))
    return out


def gen_scoring():
    import scipy.io as sio
    ns = refload.load_reference("OSIE")
    SM = ns.scanmatch.ScanMatch
    mat = sio.loadmat(os.path.join(ns.tree, "utils/evaltools/ScanMatch_DataExample.mat"))
    data = [np.asarray(mat["data%d" % i], dtype=np.float64) for i in (1, 2, 3)]

    def score4(wd_obj, wod_obj, a, b, stim):
        s1 = wd_obj.fixationToSequence(a).astype(np.int32)
        s2 = wd_obj.fixationToSequence(b).astype(np.int32)
        with np.errstate(all="ignore"):
            wd = wd_obj.match(s1, s2)[0]
        t1 = wod_obj.fixationToSequence(a).astype(np.int32)
        t2 = wod_obj.fixationToSequence(b).astype(np.int32)
        with np.errstate(all="ignore"):
            wod = wod_obj.match(t1, t2)[0]
        sed = ns.vame.string_edit_distance(stim, a, b)
        stde = ns.vame.scaled_time_delay_embedding_similarity(a, b, stim)
        return wd, wod, sed, np.nan if stde is None else stde, len(s1), len(s2)

    # --- the reference's only fixture, both configurations
    out = {"data1": data[0], "data2": data[1], "data3": data[2]}
    cfgs = {
        "main": (dict(Xres=1024, Yres=768, Xbin=12, Ybin=8, Offset=(0, 0), Threshold=3.5), 100, (768, 1024, 3), 1.0),
        "eval": (dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5), 50, (240, 320, 3), 0.3125),
    }
    for name, (cfg, tb, shp, sc) in cfgs.items():
        wd_obj, wod_obj = SM(TempBin=tb, **cfg), SM(**cfg)
        stim = np.zeros(shp, dtype=np.float32)
        rows = []
        for i, j in [(0, 1), (0, 2), (1, 2), (1, 0), (2, 2)]:
            a, b = data[i] * [sc, sc, 1.0], data[j] * [sc, sc, 1.0]
            rows.append((i, j) + score4(wd_obj, wod_obj, a, b, stim))
        out["mat_" + name] = np.array(rows, dtype=np.float64)
        out["sub_" + name] = wd_obj.SubMatrix
        out["mask_" + name] = wd_obj.mask.astype(np.int32)
    # temporal-binning micro golden
    micro = np.array([(10, 10, 125), (30, 10, 75), (50, 10, 24.9), (70, 10, 25.0), (90, 10, 25.1),
                      (400, 300, 49.99), (-5, -5, 100)], dtype=np.float64)
    cfg = cfgs["eval"][0]
    out["micro_in"] = micro
    out["micro_wd"] = SM(TempBin=50, **cfg).fixationToSequence(micro).astype(np.int32)
    out["micro_wod"] = SM(**cfg).fixationToSequence(micro).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "scoring_mat.npz"), **out)

    # --- random (human, prediction) pairs in the eval configuration, ms units
    rng = np.random.default_rng(0)
    wd_obj, wod_obj = SM(TempBin=50, **cfg), SM(**cfg)
    stim = np.zeros((240, 320, 3), dtype=np.float32)
    gts, prs = [], []
    H = human_paths(rng, 96)
    P = pred_paths(rng, 96)
    for h, p in zip(H, P):
        gts.append(h * [1, 1, 1000.0]); prs.append(p * [1, 1, 1000.0])
    # edge cases: out-of-range / negative coordinates, zero and sub-bin durations,
    # length-1 paths, long durations (long wd strings), both wd strings empty (NaN)
    e = []
    e.append((np.array([[-10., 5., 120.], [330., 250., 260.], [319.99, 239.99, 75.], [0., 0., 25.]]),
              np.array([[4., 4., 100.], [316., 236., 300.]])))
    e.append((np.array([[100., 100., 10.]]), np.array([[100., 100., 20.]])))          # both wd empty -> NaN
    e.append((np.array([[100., 100., 10.]]), np.array([[100., 100., 200.]])))         # one empty -> 0
    e.append((np.array([[12., 200., 251.]]), np.array([[300., 20., 249.]])))          # length 1 v 1
    e.append((human_paths(rng, 1, 20, 20)[0] * [1, 1, 1000.], pred_paths(rng, 1, 16, 16)[0] * [1, 1, 1000.]))
    e.append((human_paths(rng, 1, 30, 30)[0] * [1, 1, 1000.], pred_paths(rng, 1, 2, 2)[0] * [1, 1, 1000.]))
    lp = pred_paths(rng, 1, 16, 16)[0] * [1, 1, 1000.]
    lp[:, 2] = np.exp(rng.normal(np.log(4500.), 1.0, 16)).astype(np.float32)       # random-init-like: long strings
    e.append((human_paths(rng, 1, 10, 10)[0] * [1, 1, 1000.], lp))
    lh = human_paths(rng, 1, 14, 14)[0] * [1, 1, 1000.]
    lh[:, 2] *= 12
    e.append((lh, lp.copy()))
    e.append((np.array([[64., 48., 75.], [63.999, 47.999, 125.], [128.5, 96.5, 175.]]),
              np.array([[60., 44., 75.], [68., 52., 125.]])))                        # SED/grid boundaries, .5 ties
    e.append((np.array([[500., 400., 100.], [-3., 700., 100.]]), np.array([[4., 4., 100.]])))  # SED symbols > 24
    for a, b in e:
        gts.append(np.asarray(a, dtype=np.float64)); prs.append(np.asarray(b, dtype=np.float64))
    res = np.array([score4(wd_obj, wod_obj, a, b, stim) for a, b in zip(gts, prs)], dtype=np.float64)
    g, gl = pad(gts); p, pl = pad(prs)
    np.savez_compressed(os.path.join(HERE, "scoring_random.npz"), gt=g, gt_len=gl, pred=p, pred_len=pl,
                        wd=res[:, 0], wod=res[:, 1], sed=res[:, 2].astype(np.int64), stde=res[:, 3],
                        n_wd_gt=res[:, 4].astype(np.int64), n_wd_pred=res[:, 5].astype(np.int64))
    print("scoring goldens:", len(gts), "pairs; NaN wd rows:", int(np.isnan(res[:, 0]).sum()))


def gen_eval():
    """evaluation / human_evaluation / pairs_eval / pairs_eval_scanmatch drivers."""
    rng = np.random.default_rng(1)
    N, K, S = 6, 3, 5
    humans = [human_paths(rng, S, 2, 12) for _ in range(N)]
    humans[2][1] = humans[2][1][:2]          # a 2-fixation subject -> MultiMatch-NaN rule fires
    preds = [pred_paths(rng, K, 1, 16) for _ in range(N)]
    preds[4][0] = preds[4][0][:2]            # a 2-fixation prediction
    out = {}
    hp, hl = pad([h for img in humans for h in img])
    pp, pl = pad([p for img in preds for p in img])
    out.update(human=hp.reshape(N, S, -1, 3), human_len=hl.reshape(N, S),
               pred=pp.reshape(N, K, -1, 3), pred_len=pl.reshape(N, K))

    ns = refload.load_reference("OSIE")
    ev = ns.evaluation
    SM = ns.scanmatch.ScanMatch
    cfg = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
    wd_obj, wod_obj = SM(TempBin=50, **cfg), SM(**cfg)
    gt_struct = [[to_struct(h) for h in img] for img in humans]
    # evaluation(): test.py:124-149 layout = for trial: extend(all images)
    all_gt, all_pred = [], []
    for k in range(K):
        for i in range(N):
            all_gt.append(gt_struct[i]); all_pred.append(to_struct(preds[i][k]))
    m, s, per = ev.evaluation(all_gt, all_pred)
    out["evaluation_mean"] = np.array([m["ScanMatch"]["w/o duration"], m["ScanMatch"]["with duration"],
                                       m["VAME"]["SED"], m["VAME"]["STDE"], m["VAME"]["SED_best"], m["VAME"]["STDE_best"]])
    out["evaluation_std"] = np.array([s["ScanMatch"]["w/o duration"], s["ScanMatch"]["with duration"],
                                      s["VAME"]["SED"], s["VAME"]["STDE"], s["VAME"]["SED_best"], s["VAME"]["STDE_best"]])
    out["evaluation_per_image"] = np.array(per, dtype=np.float64)[:, 5:]      # drop MultiMatch stub slots

    class Loader(list):
        pass
    loader = Loader([{"fix_vectors": gt_struct[:3], "img_names": ["a", "b", "c"]},
                     {"fix_vectors": gt_struct[3:], "img_names": ["d", "e", "f"]}])
    m, s, per = ev.human_evaluation(loader)
    out["human_mean"] = np.array([m["ScanMatch"]["w/o duration"], m["ScanMatch"]["with duration"],
                                  m["VAME"]["SED"], m["VAME"]["STDE"], m["VAME"]["SED_best"], m["VAME"]["STDE_best"]])
    out["human_std"] = np.array([s["ScanMatch"]["w/o duration"], s["ScanMatch"]["with duration"],
                                 s["VAME"]["SED"], s["VAME"]["STDE"], s["VAME"]["SED_best"], s["VAME"]["STDE_best"]])
    out["human_per_image"] = np.array([per[k] for k in "abcdef"], dtype=np.float64)[:, 5:]

    pe = []
    for k in range(K):
        pe.append(ev.pairs_eval(gt_struct, [to_struct(preds[i][k]) for i in range(N)], wd_obj, wod_obj))
    out["pairs_eval"] = np.array(pe, dtype=np.float64)          # [K, N, 11], MM slots are stub values

    ns = refload.load_reference("COCO_Search18")
    SM = ns.scanmatch.ScanMatch
    wd_obj, wod_obj = SM(TempBin=50, **cfg), SM(**cfg)
    pe = []
    for k in range(K):
        pe.append(ns.evaluation.pairs_eval_scanmatch(gt_struct, [to_struct(preds[i][k]) for i in range(N)],
                                                     wd_obj, wod_obj))
    out["pairs_eval_scanmatch"] = np.array(pe, dtype=np.float64)  # [K, N, 2]
    # COCO human_evaluation: variable number of subjects per image, per-image best (:88-125)
    sizes = [3, 5, 2, 4, 5, 3]
    ragged = [gt_struct[i][:sizes[i]] for i in range(N)]
    loader = Loader([{"fix_vectors": ragged[:4], "img_names": ["a", "b", "c", "d"]},
                     {"fix_vectors": ragged[4:], "img_names": ["e", "f"]}])
    m, s, per = ns.evaluation.human_evaluation(loader)
    out["coco_human_sizes"] = np.array(sizes)
    out["coco_human_mean"] = np.array([m["ScanMatch"]["w/o duration"], m["ScanMatch"]["with duration"],
                                       m["VAME"]["SED"], m["VAME"]["STDE"], m["VAME"]["SED_best"], m["VAME"]["STDE_best"]])
    out["coco_human_std"] = np.array([s["ScanMatch"]["w/o duration"], s["ScanMatch"]["with duration"],
                                      s["VAME"]["SED"], s["VAME"]["STDE"], s["VAME"]["SED_best"], s["VAME"]["STDE_best"]])
    out["coco_human_per_image"] = np.array([per[k] for k in "abcdef"], dtype=np.float64)[:, 5:]
    np.savez_compressed(os.path.join(HERE, "eval_drivers.npz"), **out)
    print("eval goldens written")


def gen_sampling():
    """random_sample + generate_scanpath under a seeded CPU generator, with the
    draws (q ~ Exp(1), z ~ N(0,1)) the reference consumed recorded for injection
    (multinomial == argmax((p/sum p)/q))."""
    import torch
    ns = refload.load_reference("OSIE")
    out = {}
    N, T, A = 6, 16, 1201
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(N, T, A, generator=g) * 2.0
    logits[:, :, 0] += 3.0                              # make stops reasonably likely
    probs = torch.softmax(logits, -1)
    probs[1, 3] = 0.0; probs[1, 3, 0] = 1.0             # certain stop at step 3
    probs[2, :, 0] = 0.0                                # never stops: length 16
    mu = torch.randn(N, T, generator=g) * 0.3 + np.log(0.25)
    sigma2 = torch.exp(torch.randn(N, T, generator=g) * 0.2 + np.log(0.15))
    out.update(probs=probs.numpy(), mu=mu.numpy(), sigma2=sigma2.numpy())
    for min_len in (1, 2):
        sampler = ns.sampling.Sampling(convLSTM_length=T, min_length=min_len)
        for trial, seed in enumerate((11, 12, 13)):
            torch.manual_seed(seed)
            q = torch.empty(N * T, A).exponential_(1)
            z = torch.randn(N, T)
            torch.manual_seed(seed)
            s = sampler.random_sample(probs, mu, sigma2)
            images = torch.zeros(N, 3, 4, 4)
            fv, am, dm = sampler.generate_scanpath(images, s["selected_actions_probs"], s["durations"],
                                                   s["selected_actions"])
            tag = "m%d_t%d_" % (min_len, trial)
            fx, fl = pad([np.stack([f["start_x"], f["start_y"], f["duration"]], 1) if len(f) else np.zeros((0, 3))
                          for f in fv], 16)
            out.update({tag + "q": q.numpy().reshape(N, T, A), tag + "z": z.numpy(),
                        tag + "actions": s["selected_actions"].numpy(), tag + "sel_prob": s["selected_actions_probs"].numpy(),
                        tag + "dur": s["durations"].numpy(), tag + "length": s["scanpath_length"].numpy(),
                        tag + "action_mask": am.numpy(), tag + "duration_mask": dm.numpy(),
                        tag + "fix": fx, tag + "fix_len": fl})
            # log-likelihoods used by SCST (loss.py:34-45)
            out[tag + "log_action"] = ns.loss.LogAction(s["selected_actions_probs"], am).numpy()
            out[tag + "log_duration"] = ns.loss.LogDuration(s["durations"].clone(), mu, sigma2, dm).numpy()
    # supervised losses on a random target (loss.py:10-32)
    gt_onehot = torch.zeros(N, T, A)
    idx = torch.randint(0, A, (N, T), generator=g)
    gt_onehot.scatter_(2, idx.unsqueeze(-1), 1.0)
    mask = (torch.rand(N, T, generator=g) > 0.3).float()
    gt_dur = torch.exp(torch.randn(N, T, generator=g) * 0.4 + np.log(0.25))
    out.update(loss_logits=logits.numpy(), loss_gt_idx=idx.numpy(), loss_mask=mask.numpy(), loss_gt_dur=gt_dur.numpy(),
               loss_ce=ns.loss.CrossEntropyLoss(logits, gt_onehot, mask).numpy(),
               loss_lognormal=ns.loss.MLPLogNormalDistribution(mu, sigma2, gt_dur, mask).numpy())
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **out)
    print("sampling goldens written")


def gen_air_eval():
    """AiR's performance-related drivers (AiR/utils/evaluation.py:188-577)."""
    rng = np.random.default_rng(3)
    N, S = 5, 6
    humans = [human_paths(rng, S, 2, 12) for _ in range(N)]
    humans[1][2] = humans[1][2][:2]                      # MultiMatch-NaN rule
    perf = [[bool(b) for b in rng.integers(0, 2, S)] for _ in range(N)]
    perf[3] = [True] * S                                 # no wrong answerers for image 3
    preds = pred_paths(rng, N, 3, 16)
    alloc = [bool(b) for b in rng.integers(0, 2, N)]
    hp, hl = pad([h for img in humans for h in img])
    pp, pl = pad(preds)
    out = dict(human=hp.reshape(N, S, -1, 3), human_len=hl.reshape(N, S), pred=pp, pred_len=pl,
               perf=np.array(perf), alloc=np.array(alloc))
    ns = refload.load_reference("AiR")
    ev = ns.evaluation
    SM = ns.scanmatch.ScanMatch
    cfg = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
    wd_obj, wod_obj = SM(TempBin=50, **cfg), SM(**cfg)
    gt_struct = [[to_struct(h) for h in img] for img in humans]
    pr_struct = [to_struct(p) for p in preds]
    for given in (True, False):
        same, diff, flag = ev.pairs_eval_scanmatch_performance_related(gt_struct, pr_struct, wd_obj, wod_obj, perf, given)
        out["pesm_same_%d" % given], out["pesm_diff_%d" % given], out["pesm_flag_%d" % given] = same, diff, flag
    g, p, gp = ev.gtpairs_eval_scanmatch_performance_related(gt_struct, wd_obj, wod_obj, perf)
    out["gtp_good"], out["gtp_poor"], out["gtp_good_vs_poor"] = g, p, gp
    # evaluation_performance_related filters its lists with `_ != []` (AiR/utils/evaluation.py:277-279),
    # which numpy >= 2 rejects (shape mismatch raises instead of returning a scalar).  To pin the
    # aggregation anyway the function is re-executed IN MEMORY with those three comparisons spelt
    # `len(_) != 0` (what numpy 1.x evaluated them to); nothing else is touched, nothing is written.
    import inspect
    src = inspect.getsource(ev.evaluation_performance_related).replace("if _ != []]", "if len(_) != 0]")
    assert src.count("if len(_) != 0]") == 3
    scope = dict(ev.__dict__)
    exec(compile(src, "<evaluation_performance_related, numpy-2 shim>", "exec"), scope)
    m, s, per = scope["evaluation_performance_related"](gt_struct, pr_struct, perf, alloc)
    cats = ["all", "right_answer", "wrong_answer"]
    flat = lambda d: np.array([[d[c]["ScanMatch"]["w/o duration"], d[c]["ScanMatch"]["with duration"],
                                d[c]["VAME"]["SED"], d[c]["VAME"]["STDE"], d[c]["VAME"]["SED_best"],
                                d[c]["VAME"]["STDE_best"]] for c in cats])
    out["epr_mean"], out["epr_std"] = flat(m), flat(s)
    out["epr_per_image"] = np.array(per, dtype=np.float64)[:, 5:]
    np.savez_compressed(os.path.join(HERE, "eval_air.npz"), **out)
    print("air eval goldens written")


def gen_align():
    """ScanMatch.match's full return value (score, alignment, transposed F matrix; scanmatch.py:135-197) for the
    .mat fixture in the evaluation configuration (with- and without-duration strings), a GapValue != 0 case and
    short strings -- every reference caller drops align and F, the mirror returns them all the same."""
    import scipy.io as sio
    ns = refload.load_reference("OSIE")
    SM = ns.scanmatch.ScanMatch
    mat = sio.loadmat(os.path.join(ns.tree, "utils/evaltools/ScanMatch_DataExample.mat"))
    data = [np.asarray(mat["data%d" % i], dtype=np.float64) * [0.3125, 0.3125, 1.0] for i in (1, 2, 3)]
    cfg = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
    cases = []
    for gap, tb in ((0.0, 50), (0.0, 0), (-0.5, 0), (-1.25, 50), (0.75, 0)):
        kw = dict(cfg, GapValue=gap)
        if tb:
            kw["TempBin"] = tb
        obj = SM(**kw)
        for i, j in ((0, 1), (1, 2), (2, 2)):
            cases.append((obj, gap, tb, obj.fixationToSequence(data[i]).astype(np.int32),
                          obj.fixationToSequence(data[j]).astype(np.int32)))
    obj = SM(**cfg)
    rng = np.random.default_rng(5)
    cases.append((obj, 0.0, 0, np.array([17], np.int32), np.array([17, 3, 190], np.int32)))
    cases.append((obj, 0.0, 0, rng.integers(0, 192, 7).astype(np.int32), np.array([5], np.int32)))
    cases.append((obj, 0.0, 0, rng.integers(0, 192, 40).astype(np.int32), rng.integers(0, 192, 33).astype(np.int32)))
    out = {"n_cases": np.int64(len(cases))}
    for k, (o, gap, tb, A, B) in enumerate(cases):
        score, align, F = o.match(A, B)
        out["c%d_A" % k], out["c%d_B" % k] = A, B
        out["c%d_gap" % k], out["c%d_tempbin" % k] = np.float64(gap), np.int64(tb)
        out["c%d_score" % k], out["c%d_align" % k], out["c%d_F" % k] = np.float64(score), np.asarray(align), np.asarray(F)
    np.savez_compressed(os.path.join(HERE, "scoring_align.npz"), **out)
    print("alignment goldens:", len(cases), "cases; F shapes", [out["c%d_F" % k].shape for k in range(len(cases))][:6])


def gen_tde():
    """The helpers of the STDE family that the drivers reach only through scaled_time_delay_embedding_similarity
    (visual_attention_metrics.py:205-218, 332-390, 444-492): euclidean_distance, time_delay_embedding_distance for
    every k in both distance modes, scaled_time_delay_embedding_distance."""
    ns = refload.load_reference("OSIE")
    V = ns.vame
    rng = np.random.default_rng(11)
    pairs = [(human_paths(rng, 1, lh, lh)[0], pred_paths(rng, 1, ls, ls)[0])
             for lh, ls in ((9, 12), (14, 16), (6, 3), (1, 5), (12, 12), (20, 9), (3, 3))]
    out = {"n_cases": np.int64(len(pairs))}
    stim = np.zeros((240, 320, 3), dtype=np.float32)
    for c, (h, s) in enumerate(pairs):
        kmax = min(len(h), len(s))
        out["c%d_h" % c], out["c%d_s" % c] = h, s
        out["c%d_mean" % c] = np.array([V.time_delay_embedding_distance(h, s, k=k, distance_mode="Mean") for k in range(1, kmax + 1)])
        out["c%d_haus" % c] = np.array([V.time_delay_embedding_distance(h, s, k=k, distance_mode="Hausdorff") for k in range(1, kmax + 1)])
        out["c%d_scaled" % c] = np.float64(V.scaled_time_delay_embedding_distance(h, s, stim))
        e = V.euclidean_distance(h, s)
        out["c%d_euclid" % c] = np.float64(np.nan if e is False else e)
        assert V.time_delay_embedding_distance(h, s, k=kmax + 1) is False
        assert V.time_delay_embedding_distance(h, s, k=1, distance_mode="nope") is False
    np.savez_compressed(os.path.join(HERE, "vame_tde.npz"), **out)
    print("TDE goldens:", len(pairs), "pairs; euclid defined for", int(sum(np.isfinite(out["c%d_euclid" % c]) for c in range(len(pairs)))))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("tde", "all"):
        gen_tde()
    if what in ("align", "all"):
        gen_align()
    if what in ("scoring", "all"):
        gen_scoring()
    if what in ("eval", "all"):
        gen_eval()
    if what in ("sampling", "all"):
        gen_sampling()
    if what in ("air", "all"):
        gen_air_eval()
    if what in ("decoder", "all"):
        from make_decoder_goldens import gen_decoder
        gen_decoder()

