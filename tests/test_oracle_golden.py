"""The CPU oracle against the golden vectors recorded from the reference
(tests/golden/make_goldens.py).  Runs everywhere (no GPU, no /root/reference)."""
import os

import numpy as np
import pytest

from oracle import sampling as OS
from oracle import scoring as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _unpad(arr, lens):
    return [arr[i, :lens[i]].copy() for i in range(len(lens))]


CFG = {"main": (dict(Xres=1024, Yres=768, Xbin=12, Ybin=8, Offset=(0, 0), Threshold=3.5), 100, (768, 1024, 3), 1.0),
       "eval": (dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5), 50, (240, 320, 3), 0.3125)}


@pytest.mark.parametrize("name", ["main", "eval"])
def test_mat_fixture(golden_dir, name):
    g = _load(golden_dir, "scoring_mat.npz")
    cfg, tb, shp, sc = CFG[name]
    wd, wod = O.ScanMatchOracle(TempBin=tb, **cfg), O.ScanMatchOracle(**cfg)
    assert np.array_equal(wd.SubMatrix, g["sub_" + name])          # bit-exact table
    assert np.array_equal(wd.mask.astype(np.int32), g["mask_" + name])
    data = [g["data1"], g["data2"], g["data3"]]
    stim = np.zeros(shp, np.float32)
    for row in g["mat_" + name]:
        a, b = data[int(row[0])] * [sc, sc, 1.0], data[int(row[1])] * [sc, sc, 1.0]
        s1, s2 = wd.fixationToSequence(a), wd.fixationToSequence(b)
        assert (len(s1), len(s2)) == (int(row[6]), int(row[7]))
        assert wd.match_score(s1, s2) == row[2]
        assert wod.match_score(wod.fixationToSequence(a), wod.fixationToSequence(b)) == row[3]
        assert O.string_edit_distance(stim, a, b) == int(row[4])
        assert O.scaled_time_delay_embedding_similarity(a, b, stim) == pytest.approx(row[5], rel=1e-14)


def test_survey_known_answers(golden_dir):
    """The numbers SURVEY.md section 8c lists, as literal constants."""
    g = _load(golden_dir, "scoring_mat.npz")
    m = g["mat_main"]
    assert m[0, 2] == 0.6725138474550876 and m[0, 3] == 0.6178313750019084 and m[0, 4] == 9
    assert m[0, 5] == 0.9064806433533912 and m[3, 5] == 0.8540590287740126
    e = g["mat_eval"]
    assert e[0, 2] == 0.6535157780932709 and e[0, 3] == 0.6054726619924844 and e[0, 4] == 10
    assert e[1, 2] == 0.14615608621524606 and e[2, 4] == 19
    assert list(g["micro_wd"]) == [0, 0, 1, 1, 191, 0, 0]
    assert list(g["micro_wod"]) == [0, 1, 2, 3, 4, 191, 0]


def test_temporal_binning_micro(golden_dir):
    g = _load(golden_dir, "scoring_mat.npz")
    wd, wod = O.eval_scanmatch_objects()
    assert np.array_equal(wd.fixationToSequence(g["micro_in"]).astype(np.int32), g["micro_wd"])
    assert np.array_equal(wod.fixationToSequence(g["micro_in"]).astype(np.int32), g["micro_wod"])


def test_random_pairs(golden_dir):
    g = _load(golden_dir, "scoring_random.npz")
    gts, prs = _unpad(g["gt"], g["gt_len"]), _unpad(g["pred"], g["pred_len"])
    for i, (a, b) in enumerate(zip(gts, prs)):
        wd, wod, sed, stde = O.score_pair(a, b)
        assert np.array_equal(np.float64(wd), g["wd"][i], equal_nan=True), i
        assert wod == g["wod"][i], i
        assert sed == g["sed"][i], i
        assert stde == pytest.approx(g["stde"][i], rel=1e-13), i


def _struct_lists(g):
    from golden.make_goldens import to_struct  # noqa: F401  (dtype helper only; does not touch the reference)
    N, S = g["human_len"].shape
    K = g["pred_len"].shape[1]
    humans = [[to_struct(g["human"][i, s, :g["human_len"][i, s]]) for s in range(S)] for i in range(N)]
    preds = [[to_struct(g["pred"][i, k, :g["pred_len"][i, k]]) for k in range(K)] for i in range(N)]
    return humans, preds, N, K, S


def _flat(m):
    return np.array([m["ScanMatch"]["w/o duration"], m["ScanMatch"]["with duration"], m["VAME"]["SED"],
                     m["VAME"]["STDE"], m["VAME"]["SED_best"], m["VAME"]["STDE_best"]])


ALIGN_CFG = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Threshold=3.5)


def test_match_alignment_and_matrix(golden_dir):
    """ScanMatch.match's whole return value (score, alignment, transposed F; scanmatch.py:135-197) for the 18
    recorded cases: with / without duration strings of the .mat fixture, three non-zero gap values, short strings."""
    g = np.load(os.path.join(golden_dir, "scoring_align.npz"))
    for k in range(int(g["n_cases"])):
        o = O.ScanMatchOracle(GapValue=float(g["c%d_gap" % k]), **ALIGN_CFG)
        score, align, F = o.match(g["c%d_A" % k], g["c%d_B" % k])
        assert score == g["c%d_score" % k], k
        assert np.array_equal(align, g["c%d_align" % k]), k
        assert np.array_equal(F, g["c%d_F" % k]), k


def test_tde_family(golden_dir):
    """euclidean_distance, time_delay_embedding_distance (every k, both modes), scaled_time_delay_embedding_distance
    (visual_attention_metrics.py:205-218, 332-390, 444-492) against the recorded reference outputs."""
    g = np.load(os.path.join(golden_dir, "vame_tde.npz"))
    stim = np.zeros((240, 320, 3), dtype=np.float32)
    for c in range(int(g["n_cases"])):
        h, s = g["c%d_h" % c], g["c%d_s" % c]
        kmax = min(len(h), len(s))
        for k in range(1, kmax + 1):
            assert O.time_delay_embedding_distance(h, s, k, "Mean") == pytest.approx(g["c%d_mean" % c][k - 1], rel=1e-13)
            assert O.time_delay_embedding_distance(h, s, k, "Hausdorff") == pytest.approx(g["c%d_haus" % c][k - 1], rel=1e-13)
        assert O.time_delay_embedding_distance(h, s, kmax + 1) is False
        assert O.time_delay_embedding_distance(h, s, 1, "nope") is False
        assert O.scaled_time_delay_embedding_distance(h, s, stim) == pytest.approx(float(g["c%d_scaled" % c]), rel=1e-13)
        e = O.euclidean_distance(h, s)
        if np.isnan(g["c%d_euclid" % c]):
            assert e is False
        else:
            assert e == pytest.approx(float(g["c%d_euclid" % c]), rel=1e-13)


def test_eval_drivers(golden_dir):
    g = _load(golden_dir, "eval_drivers.npz")
    humans, preds, N, K, S = _struct_lists(g)
    all_gt, all_pred = [], []
    for k in range(K):
        for i in range(N):
            all_gt.append(humans[i]); all_pred.append(preds[i][k])
    m, s, per = O.evaluation(all_gt, all_pred)
    np.testing.assert_allclose(_flat(m), g["evaluation_mean"], rtol=1e-13)
    np.testing.assert_allclose(_flat(s), g["evaluation_std"], rtol=1e-12)
    np.testing.assert_allclose(np.array(per), g["evaluation_per_image"], rtol=1e-13)
    m, s, per = O.human_evaluation(humans)
    np.testing.assert_allclose(_flat(m), g["human_mean"], rtol=1e-13)
    np.testing.assert_allclose(_flat(s), g["human_std"], rtol=1e-12)
    np.testing.assert_allclose(np.array(per), g["human_per_image"], rtol=1e-13)
    ragged = [humans[i][:int(g["coco_human_sizes"][i])] for i in range(N)]
    m, s, per = O.human_evaluation(ragged, per_image_best=True)
    np.testing.assert_allclose(_flat(m), g["coco_human_mean"], rtol=1e-13)
    np.testing.assert_allclose(_flat(s), g["coco_human_std"], rtol=1e-12)
    np.testing.assert_allclose(np.array(per), g["coco_human_per_image"], rtol=1e-13)
    for k in range(K):
        pe = O.pairs_eval(humans, [preds[i][k] for i in range(N)])
        np.testing.assert_allclose(pe[:, 5:], g["pairs_eval"][k][:, 5:], rtol=1e-6, equal_nan=True)
        assert np.array_equal(np.isnan(pe[:, 5]), np.isnan(g["pairs_eval"][k][:, 5]))
        ps = O.pairs_eval_scanmatch(humans, [preds[i][k] for i in range(N)])
        np.testing.assert_allclose(ps, g["pairs_eval_scanmatch"][k], rtol=1e-13)


@pytest.mark.parametrize("min_len", [1, 2])
def test_sampling(golden_dir, min_len):
    g = _load(golden_dir, "sampling.npz")
    for trial in range(3):
        tag = "m%d_t%d_" % (min_len, trial)
        s = OS.random_sample(g["probs"], g["mu"], g["sigma2"], g[tag + "q"], g[tag + "z"], min_len)
        assert np.array_equal(s["selected_actions"], g[tag + "actions"])
        assert np.array_equal(s["selected_actions_probs"], g[tag + "sel_prob"])
        np.testing.assert_allclose(s["durations"], g[tag + "dur"], rtol=2e-7)
        assert np.array_equal(s["scanpath_length"], g[tag + "length"].reshape(-1, 1))
        fix, am, dm = OS.generate_scanpath(g[tag + "actions"], g[tag + "dur"])
        assert np.array_equal(am, g[tag + "action_mask"]) and np.array_equal(dm, g[tag + "duration_mask"])
        for n, f in enumerate(fix):
            assert len(f) == g[tag + "fix_len"][n]
            assert np.array_equal(f, g[tag + "fix"][n, :len(f)])
        np.testing.assert_allclose(OS.log_action(g[tag + "sel_prob"], am), g[tag + "log_action"], rtol=1e-5)
        np.testing.assert_allclose(OS.log_duration(g[tag + "dur"], g["mu"], g["sigma2"], dm),
                                   g[tag + "log_duration"], rtol=1e-5)


def test_supervised_losses(golden_dir):
    g = _load(golden_dir, "sampling.npz")
    assert OS.cross_entropy_loss(g["loss_logits"], g["loss_gt_idx"], g["loss_mask"]) == pytest.approx(
        float(g["loss_ce"]), rel=1e-5)
    assert OS.lognormal_nll(g["mu"], g["sigma2"], g["loss_gt_dur"], g["loss_mask"]) == pytest.approx(
        float(g["loss_lognormal"]), rel=1e-5)


def _air_lists(g):
    from golden.make_goldens import to_struct
    N, S = g["human_len"].shape
    humans = [[to_struct(g["human"][i, s, :g["human_len"][i, s]]) for s in range(S)] for i in range(N)]
    preds = [to_struct(g["pred"][i, :g["pred_len"][i]]) for i in range(N)]
    perf = [[bool(v) for v in row] for row in g["perf"]]
    alloc = [bool(v) for v in g["alloc"]]
    return humans, preds, perf, alloc


def test_air_performance_related_drivers(golden_dir):
    g = _load(golden_dir, "eval_air.npz")
    humans, preds, perf, alloc = _air_lists(g)
    for given in (True, False):
        same, diff, flag = O.pairs_eval_scanmatch_performance_related(humans, preds, perf, given)
        np.testing.assert_allclose(same, g["pesm_same_%d" % given], rtol=1e-13, equal_nan=True)
        np.testing.assert_allclose(diff, g["pesm_diff_%d" % given], rtol=1e-13, equal_nan=True)
        assert flag == bool(g["pesm_flag_%d" % given])
    good, poor, gp = O.gtpairs_eval_scanmatch_performance_related(humans, perf)
    np.testing.assert_allclose(good, g["gtp_good"], rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(poor, g["gtp_poor"], rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(gp, g["gtp_good_vs_poor"], rtol=1e-13, equal_nan=True)
    mean, std, per = O.evaluation_performance_related(humans, preds, perf, alloc)
    np.testing.assert_allclose(mean, g["epr_mean"], rtol=2e-6)       # float32 aggregation in the reference
    np.testing.assert_allclose(std, g["epr_std"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(per, g["epr_per_image"], rtol=1e-12)
