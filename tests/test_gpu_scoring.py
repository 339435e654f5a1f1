"""Parity of the CUDA scoring path (through the C ABI) with the oracle and with
the goldens recorded from the reference.  Gates: SED bit-exact; ScanMatch
(evaluated in f64 in the reference's operation order) bit-exact; STDE within
1e-12 relative (gate of north_star: 1e-5)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STDE_RTOL = 1e-12


@pytest.fixture(scope="module")
def S():
    from scanpaths_b200 import build, scoring
    build.build_library()
    return scoring


def _unpad(arr, lens):
    return [arr[i, :lens[i]].copy() for i in range(len(lens))]


def _pairs_identity(n, dev):
    i = torch.arange(n, dtype=torch.int32, device=dev)
    return i, i.clone()


def test_golden_random_pairs(S, golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring_random.npz"))
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)           # goldens are already in ms
    hp = S.pack_paths(_unpad(g["gt"], g["gt_len"]), cfg)
    pp = S.pack_paths(_unpad(g["pred"], g["pred_len"]), cfg)
    assert np.array_equal(hp.nwd.cpu().numpy(), g["n_wd_gt"])
    assert np.array_equal(pp.nwd.cpu().numpy(), g["n_wd_pred"])
    ph, ps = _pairs_identity(hp.n, cfg.device)
    out = S.score_pairs(hp, pp, ph, ps, cfg).cpu().numpy()
    assert np.array_equal(out[:, 0], g["wd"], equal_nan=True)
    assert np.array_equal(out[:, 1], g["wod"])
    assert np.array_equal(out[:, 2].astype(np.int64), g["sed"])
    np.testing.assert_allclose(out[:, 3], g["stde"], rtol=STDE_RTOL)


@pytest.mark.parametrize("name", ["main", "eval"])
def test_golden_mat_fixture(S, golden_dir, name):
    from test_oracle_golden import CFG
    g = np.load(os.path.join(golden_dir, "scoring_mat.npz"))
    cfgd, tb, shp, sc = CFG[name]
    cfg = S.ScoreConfig(TempBin=tb, stimulus_shape=shp, dur_scale=1.0, **cfgd)
    data = [g["data%d" % i] * [sc, sc, 1.0] for i in (1, 2, 3)]
    pack = S.pack_paths(data, cfg)
    rows = g["mat_" + name]
    dev = cfg.device
    out = S.score_pairs(pack, pack, torch.tensor(rows[:, 0], dtype=torch.int32, device=dev),
                        torch.tensor(rows[:, 1], dtype=torch.int32, device=dev), cfg).cpu().numpy()
    assert np.array_equal(out[:, 0], rows[:, 2]) and np.array_equal(out[:, 1], rows[:, 3])
    assert np.array_equal(out[:, 2], rows[:, 4])
    np.testing.assert_allclose(out[:, 3], rows[:, 5], rtol=STDE_RTOL)


def test_prep_micro_golden(S, golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring_mat.npz"))
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    pack = S.pack_paths([g["micro_in"]], cfg)
    sym = pack.sym[0].cpu().numpy(); run = pack.run[0].cpu().numpy()
    assert np.array_equal(sym, g["micro_wod"])
    assert np.array_equal(np.repeat(sym, run), g["micro_wd"])


def _random_set(rng, n, lh=(1, 20), lp=(1, 16), long_frac=0.0):
    from golden.make_goldens import human_paths, pred_paths
    H = [h * [1, 1, 1000.0] for h in human_paths(rng, n, *lh)]
    P = [p * [1, 1, 1000.0] for p in pred_paths(rng, n, *lp)]
    for i in range(int(n * long_frac)):
        P[i][:, 2] *= rng.uniform(5, 60)
        H[i][:, 2] *= rng.uniform(1, 8)
    for h in H[::7]:                                   # out-of-range coordinates
        h[:, 0] += rng.uniform(-60, 60); h[:, 1] -= rng.uniform(0, 50)
    return H, P


def test_random_vs_c_oracle(S):
    """20k pairs incl. long with-duration strings (multi-panel path) vs the C oracle."""
    from oracle import c_scoring as CO
    rng = np.random.default_rng(123)
    H, P = _random_set(rng, 4000, long_frac=0.1)
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    hp, pp = S.pack_paths(H, cfg), S.pack_paths(P, cfg)
    gi = rng.integers(0, len(H), 20000); pi = rng.integers(0, len(P), 20000)
    dev = cfg.device
    out = S.score_pairs(hp, pp, torch.tensor(gi, dtype=torch.int32, device=dev),
                        torch.tensor(pi, dtype=torch.int32, device=dev), cfg).cpu().numpy()
    ha, hl = S.pad_paths(H); pa, pl = S.pad_paths(P)
    ref = CO.score_pairs(ha, hl, pa, pl, gi, pi, threads=os.cpu_count())
    assert int(pp.nwd.max().item()) > 256, "the test must exercise the multi-panel path"
    assert np.array_equal(out[:, 0], ref[:, 0], equal_nan=True)
    assert np.array_equal(out[:, 1], ref[:, 1])
    assert np.array_equal(out[:, 2], ref[:, 2])
    np.testing.assert_allclose(out[:, 3], ref[:, 3], rtol=STDE_RTOL)


def test_human_vs_human_long_paths(S):
    """human_evaluation mode: both sides from the same pack, up to 40 fixations."""
    from oracle import c_scoring as CO
    rng = np.random.default_rng(7)
    H, _ = _random_set(rng, 60, lh=(1, 40))
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    hp = S.pack_paths(H, cfg)
    gi, pi = np.meshgrid(np.arange(60), np.arange(60), indexing="ij")
    gi, pi = gi.reshape(-1), pi.reshape(-1)
    dev = cfg.device
    out = S.score_pairs(hp, hp, torch.tensor(gi, dtype=torch.int32, device=dev),
                        torch.tensor(pi, dtype=torch.int32, device=dev), cfg).cpu().numpy()
    ha, hl = S.pad_paths(H)
    ref = CO.score_pairs(ha, hl, ha, hl, gi, pi, threads=os.cpu_count())
    assert np.array_equal(out[:, :3], ref[:, :3], equal_nan=True)
    np.testing.assert_allclose(out[:, 3], ref[:, 3], rtol=STDE_RTOL)
    d = out.reshape(60, 60, 4)
    assert np.all(np.diagonal(d[:, :, 2]) == 0)              # SED(x, x) = 0
    assert np.allclose(np.diagonal(d[:, :, 3]), 1.0)         # STDE(x, x) = 1
    assert np.array_equal(d[:, :, 1], d[:, :, 1].T)          # NW is symmetric, bit for bit
    assert np.array_equal(d[:, :, 2], d[:, :, 2].T)


@pytest.mark.parametrize("lmax", [24, 64, 100])
def test_long_scanpaths_on_both_sides(S, lmax):
    """AiR / COCO human scanpaths are not truncated: both packs with lmax 24 (fast path, 4 columns per lane for
    the fixation strings), 64 and 100 (warp-per-pair kernel, warps per block reduced to fit the STDE tiles;
    round 1 failed beyond ~42 with SPB_ERR_UNSUPPORTED)."""
    from oracle import c_scoring as CO
    rng = np.random.default_rng(lmax)
    n = 24
    H, _ = _random_set(rng, n, lh=(lmax // 2, lmax))
    H[0] = H[0][:1]; H[1] = H[1][:lmax]
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    hp = S.pack_paths(H, cfg, lmax=lmax)
    gi, pi = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    gi, pi = gi.reshape(-1), pi.reshape(-1)
    dev = cfg.device
    out = S.score_pairs(hp, hp, torch.tensor(gi, dtype=torch.int32, device=dev),
                        torch.tensor(pi, dtype=torch.int32, device=dev), cfg).cpu().numpy()
    ha, hl = S.pad_paths(H, lmax)
    ref = CO.score_pairs(ha, hl, ha, hl, gi, pi, threads=os.cpu_count())
    assert np.array_equal(out[:, :3], ref[:, :3], equal_nan=True)
    np.testing.assert_allclose(out[:, 3], ref[:, 3], rtol=STDE_RTOL)


def test_stale_pair_map_fails_loudly(S):
    """Pair indices outside their pack (a pair map built for other humans) raise instead of reading out of range."""
    from scanpaths_b200 import _lib
    rng = np.random.default_rng(1)
    H, P = _random_set(rng, 6)
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    hp, pp = S.pack_paths(H, cfg), S.pack_paths(P, cfg)
    dev = cfg.device
    bad_h = torch.tensor([0, 1, 99], dtype=torch.int32, device=dev)
    ok_s = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev)
    with pytest.raises(_lib.SpbError):
        S.score_pairs(hp, pp, bad_h, ok_s, cfg)
    out = S.score_pairs(hp, pp, bad_h, ok_s, cfg, check=False).cpu().numpy()
    assert np.isnan(out[2]).all() and np.isfinite(out[:2, :3]).all()


def test_custom_mask_from_array(S):
    """ScanMatch.maskFromArray (scanmatch.py:199-200): a [Yres, Xres] symbol table replaces the grid."""
    from oracle.scoring import ScanMatchOracle
    from scanpaths_b200.utils.evaltools.scanmatch import ScanMatch
    kw = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
    rng = np.random.default_rng(2)
    mask = rng.integers(0, 16 * 12, (240, 320)).astype(np.float64)
    sm, o = ScanMatch(**kw), ScanMatchOracle(**{k: v for k, v in kw.items() if k != "Offset"})
    sm.maskFromArray(mask)
    o.mask = mask
    data = np.stack([rng.uniform(0, 320, 9), rng.uniform(0, 240, 9), rng.uniform(50, 400, 9)], 1)
    assert np.array_equal(sm.fixationToSequence(data), o.fixationToSequence(data))
    a = sm.fixationToSequence(data).astype(np.int32)
    b = sm.fixationToSequence(data[::-1].copy()).astype(np.int32)
    assert sm.match(a, b)[0] == o.match_score(a, b)


def test_gap_value_general_path(S):
    """Non-zero GapValue (never used by the reference's drivers, supported by its class)."""
    from oracle.scoring import ScanMatchOracle
    rng = np.random.default_rng(3)
    H, P = _random_set(rng, 12)
    kw = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Threshold=3.5, GapValue=-0.4, TempBin=50)
    cfg = S.ScoreConfig(stimulus_shape=(240, 320, 3), dur_scale=1.0, **kw)
    o = ScanMatchOracle(**kw); o2 = ScanMatchOracle(**{**kw, "TempBin": 0})
    hp, pp = S.pack_paths(H, cfg), S.pack_paths(P, cfg)
    ph, ps = _pairs_identity(12, cfg.device)
    out = S.score_pairs(hp, pp, ph, ps, cfg).cpu().numpy()
    for i in range(12):
        assert out[i, 0] == o.match_score(o.fixationToSequence(H[i]), o.fixationToSequence(P[i]))
        assert out[i, 1] == o2.match_score(o2.fixationToSequence(H[i]), o2.fixationToSequence(P[i]))


def test_mirror_single_pair_api(S, golden_dir):
    """The reference's own call sequence (scanmatch.py:222-257, visual_attention_metrics.py:495-519)."""
    from scanpaths_b200.utils.evaltools.scanmatch import ScanMatch
    from scanpaths_b200.utils.evaltools.visual_attention_metrics import (
        scaled_time_delay_embedding_similarity, string_edit_distance)
    g = np.load(os.path.join(golden_dir, "scoring_mat.npz"))
    d1, d2 = g["data1"], g["data2"]
    wd = ScanMatch(Xres=1024, Yres=768, Xbin=12, Ybin=8, Offset=(0, 0), TempBin=100, Threshold=3.5)
    wod = ScanMatch(Xres=1024, Yres=768, Xbin=12, Ybin=8, Offset=(0, 0), Threshold=3.5)
    assert np.array_equal(wd.SubMatrix, g["sub_main"]) and np.array_equal(wd.mask.astype(np.int32), g["mask_main"])
    s1 = wd.fixationToSequence(d1).astype(np.int32); s2 = wd.fixationToSequence(d2).astype(np.int32)
    assert (len(s1), len(s2)) == (62, 61)
    score, align, f = wd.match(s1, s2)
    assert score == 0.6725138474550876
    t1 = wod.fixationToSequence(d1[:, :2]).astype(np.int32); t2 = wod.fixationToSequence(d2[:, :2]).astype(np.int32)
    assert wod.match(t1, t2)[0] == 0.6178313750019084
    stim = np.zeros((768, 1024, 3), dtype=np.float32)
    assert string_edit_distance(stim, d1, d2) == 9
    assert scaled_time_delay_embedding_similarity(d1, d2, stim) == pytest.approx(0.9064806433533912, rel=STDE_RTOL)
    assert scaled_time_delay_embedding_similarity(d2, d1, stim) == pytest.approx(0.8540590287740126, rel=STDE_RTOL)
    with pytest.raises(ValueError):
        ScanMatch(Foo=1)


def test_mirror_match_returns_alignment_and_matrix(S, golden_dir):
    """ScanMatch.match -> (score, align, F) as the reference returns them (scanmatch.py:135-197): the F matrix from
    spb_scanmatch_matrix, the walk back on the host.  All three exact against the recorded reference outputs:
    with / without duration strings of the .mat fixture, GapValue -0.5 / -1.25 / +0.75, short strings."""
    from scanpaths_b200.utils.evaltools.scanmatch import ScanMatch
    g = np.load(os.path.join(golden_dir, "scoring_align.npz"))
    for k in range(int(g["n_cases"])):
        sm = ScanMatch(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5, GapValue=float(g["c%d_gap" % k]))
        score, align, F = sm.match(g["c%d_A" % k], g["c%d_B" % k])
        assert score == g["c%d_score" % k], k
        assert align.shape == g["c%d_align" % k].shape and np.array_equal(align, g["c%d_align" % k]), k
        assert F.shape == g["c%d_F" % k].shape and np.array_equal(F, g["c%d_F" % k]), k
    sm = ScanMatch(Xres=320, Yres=240, Xbin=16, Ybin=12)
    with pytest.raises(IndexError):                       # the reference indexes SubMatrix[A, B]
        sm.match(np.array([5, 400]), np.array([1]))
    score, align, F = sm.match(np.array([], dtype=np.int64), np.array([3, 4]))      # one empty string
    assert score == 0.0 and F.shape == (3, 1) and np.array_equal(align, [[-1, 3], [-1, 4]])


def test_mirror_tde_family(S, golden_dir):
    """euclidean_distance, time_delay_embedding_distance (every k, 'Mean' and 'Hausdorff'),
    scaled_time_delay_embedding_distance (visual_attention_metrics.py:205-218, 332-390, 444-492) through
    spb_tde_distances against the recorded reference outputs; the reference's False / None returns."""
    from scanpaths_b200.utils.evaltools import visual_attention_metrics as V
    g = np.load(os.path.join(golden_dir, "vame_tde.npz"))
    stim = np.zeros((240, 320, 3), dtype=np.float32)
    for c in range(int(g["n_cases"])):
        h, s = g["c%d_h" % c], g["c%d_s" % c]
        kmax = min(len(h), len(s))
        for k in range(1, kmax + 1):
            assert V.time_delay_embedding_distance(h, s, k=k) == pytest.approx(g["c%d_mean" % c][k - 1], rel=STDE_RTOL)
            assert V.time_delay_embedding_distance(h, s, k=k, distance_mode="Hausdorff") == pytest.approx(
                g["c%d_haus" % c][k - 1], rel=STDE_RTOL)
        assert V.time_delay_embedding_distance(h, s, k=kmax + 1) is False
        assert V.time_delay_embedding_distance(h, s, k=1, distance_mode="nope") is False
        assert V.scaled_time_delay_embedding_distance(h, s, stim) == pytest.approx(float(g["c%d_scaled" % c]), rel=STDE_RTOL)
        e = V.euclidean_distance(h, s)
        if np.isnan(g["c%d_euclid" % c]):
            assert e is False
        else:
            assert e == pytest.approx(float(g["c%d_euclid" % c]), rel=STDE_RTOL)
    assert V.scaled_time_delay_embedding_distance(np.zeros((0, 3)), g["c0_s"], stim) is None


def test_mirror_evaluation_drivers(S, golden_dir):
    from test_oracle_golden import _flat, _struct_lists
    from scanpaths_b200.utils import evaluation as E
    g = np.load(os.path.join(golden_dir, "eval_drivers.npz"))
    humans, preds, N, K, Sn = _struct_lists(g)
    all_gt, all_pred = [], []
    for k in range(K):
        for i in range(N):
            all_gt.append(humans[i]); all_pred.append(preds[i][k])
    m, s, per = E.evaluation(all_gt, all_pred)
    np.testing.assert_allclose(_flat(m), g["evaluation_mean"], rtol=1e-12)
    np.testing.assert_allclose(_flat(s), g["evaluation_std"], rtol=1e-10)
    np.testing.assert_allclose(np.array(per)[:, 5:], g["evaluation_per_image"], rtol=1e-12)
    loader = [{"fix_vectors": humans[:3], "img_names": ["a", "b", "c"]},
              {"fix_vectors": humans[3:], "img_names": ["d", "e", "f"]}]
    m, s, per = E.human_evaluation(loader)
    np.testing.assert_allclose(_flat(m), g["human_mean"], rtol=1e-12)
    np.testing.assert_allclose(_flat(s), g["human_std"], rtol=1e-10)
    np.testing.assert_allclose(np.array([per[k] for k in "abcdef"])[:, 5:], g["human_per_image"], rtol=1e-12)
    ragged = [humans[i][:int(g["coco_human_sizes"][i])] for i in range(N)]
    loader = [{"fix_vectors": ragged[:4], "img_names": ["a", "b", "c", "d"]},
              {"fix_vectors": ragged[4:], "img_names": ["e", "f"]}]
    m, s, per = E.human_evaluation(loader, per_image_best=True)          # COCO-Search18 variant
    np.testing.assert_allclose(_flat(m), g["coco_human_mean"], rtol=1e-12)
    np.testing.assert_allclose(_flat(s), g["coco_human_std"], rtol=1e-10)
    np.testing.assert_allclose(np.array([per[k] for k in "abcdef"])[:, 5:], g["coco_human_per_image"], rtol=1e-12)
    for k in range(K):
        pe = E.pairs_eval(humans, [preds[i][k] for i in range(N)], None, None)
        np.testing.assert_allclose(pe[:, 5:], g["pairs_eval"][k][:, 5:], rtol=1e-6, equal_nan=True)
        ps = E.pairs_eval_scanmatch(humans, [preds[i][k] for i in range(N)], None, None)
        np.testing.assert_allclose(ps, g["pairs_eval_scanmatch"][k], rtol=1e-12)
    with pytest.raises(IndexError):
        from golden.make_goldens import to_struct
        E.evaluation([humans[0]], [to_struct(np.zeros((0, 3)))])


def test_reduce_reward(S):
    rng = np.random.default_rng(11)
    G, Sn = 37, 15
    sc = rng.uniform(0.05, 1, (G * Sn, 4)); sc[:, 2] = rng.integers(0, 17, G * Sn)
    sc[5, 0] = np.nan
    valid = (rng.uniform(size=(G, Sn)) > 0.2).astype(np.uint8); valid[3] = 0
    dev = torch.device("cuda")
    table, reward = S.reduce_pairs_eval(torch.tensor(sc, device=dev), Sn, torch.tensor(valid, device=dev))
    table, reward = table.cpu().numpy(), reward.cpu().numpy()
    for gi in range(G):
        rows = sc[gi * Sn:(gi + 1) * Sn]
        ok = valid[gi].astype(bool) & ~np.isnan(rows.sum(1))
        if not ok.any():
            assert np.isnan(table[gi]).all() and np.isnan(reward[gi]); continue
        r = rows[ok]
        exp = np.array([r[:, 1].sum() / Sn, r[:, 0].sum() / Sn, r[:, 2].sum() / Sn, r[:, 3].sum() / Sn,
                        r[:, 2].min(), r[:, 3].max()], dtype=np.float32)
        np.testing.assert_array_equal(table[gi, 5:], exp)
        assert (table[gi, :5] == 0).all()          # MultiMatch placeholder where the reference has numbers
        a, b = float(exp[0]), float(exp[1])
        assert reward[gi] == pytest.approx(2 / (1 / a + 1 / b), rel=1e-12)


def test_full_size_properties(S):
    """BASELINE.json config 2 pair count (4096 x 64 x 15 = 3.93 M pairs) through
    size-independent properties: scores of a pair do not depend on where it sits in
    the batch (checksum of a permuted duplicate), ranges, and a C-oracle spot check."""
    from oracle import c_scoring as CO
    rng = np.random.default_rng(2)
    N, K, Sn = 4096, 64, 15
    H, _ = _random_set(rng, 2048, lh=(6, 14))
    _, P = _random_set(rng, 8192, lp=(1, 16))
    cfg = S.ScoreConfig.evaluation(dur_scale=1.0)
    hp, pp = S.pack_paths(H, cfg), S.pack_paths(P, cfg)
    dev = cfg.device
    npairs = N * K * Sn
    gen = torch.Generator(device=dev).manual_seed(0)
    ph = torch.randint(0, len(H), (npairs,), generator=gen, device=dev, dtype=torch.int32)
    ps = torch.randint(0, len(P), (npairs,), generator=gen, device=dev, dtype=torch.int32)
    out = S.score_pairs(hp, pp, ph, ps, cfg)
    perm = torch.randperm(npairs, generator=gen, device=dev)
    out2 = S.score_pairs(hp, pp, ph[perm].contiguous(), ps[perm].contiguous(), cfg)
    assert torch.equal(out[perm], out2)
    assert float(out[:, 3].min()) > 0 and float(out[:, 3].max()) <= 1.0
    assert float(out[:, 1].max()) <= 1.0 and float(out[:, 2].min()) >= 0
    idx = torch.randint(0, npairs, (5000,), generator=gen, device=dev)
    ha, hl = S.pad_paths(H); pa, pl = S.pad_paths(P)
    ref = CO.score_pairs(ha, hl, pa, pl, ph[idx].cpu().numpy(), ps[idx].cpu().numpy(), threads=os.cpu_count())
    got = out[idx].cpu().numpy()
    assert np.array_equal(got[:, :3], ref[:, :3], equal_nan=True)
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=STDE_RTOL)


def test_mirror_air_performance_related_drivers(S, golden_dir):
    from test_oracle_golden import _air_lists
    from scanpaths_b200.utils import evaluation as E
    g = np.load(os.path.join(golden_dir, "eval_air.npz"))
    humans, preds, perf, alloc = _air_lists(g)
    for given in (True, False):
        same, diff, flag = E.pairs_eval_scanmatch_performance_related(humans, preds, None, None, perf, given)
        np.testing.assert_allclose(same, g["pesm_same_%d" % given], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(diff, g["pesm_diff_%d" % given], rtol=1e-12, equal_nan=True)
        assert flag == bool(g["pesm_flag_%d" % given])
    good, poor, gp = E.gtpairs_eval_scanmatch_performance_related(humans, None, None, perf)
    np.testing.assert_allclose(good, g["gtp_good"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(poor, g["gtp_poor"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(gp, g["gtp_good_vs_poor"], rtol=1e-12, equal_nan=True)
    m, s, per = E.evaluation_performance_related(humans, preds, perf, alloc)
    cats = ["all", "right_answer", "wrong_answer"]
    flat = lambda d: np.array([[d[c]["ScanMatch"]["w/o duration"], d[c]["ScanMatch"]["with duration"],
                                d[c]["VAME"]["SED"], d[c]["VAME"]["STDE"], d[c]["VAME"]["SED_best"],
                                d[c]["VAME"]["STDE_best"]] for c in cats])
    np.testing.assert_allclose(flat(m), g["epr_mean"], rtol=2e-6)
    np.testing.assert_allclose(flat(s), g["epr_std"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(np.array(per)[:, 5:], g["epr_per_image"], rtol=1e-12)


def test_reduce_pairs_counts_length_rule_accumulators(S):
    """spb_reduce_pairs in one pass: padded subjects skipped (divisor = real subject count), the < 3 fixations rule
    from the path lengths, mean over the surviving rows (AiR variant), per-group validity, and the `evaluation`
    accumulators -- against numpy, and bit-reproducible from call to call (no floating-point atomics)."""
    rng = np.random.default_rng(5)
    K, N, Sn = 7, 9, 6
    G = K * N
    sc = rng.uniform(0.05, 1, (G * Sn, 4)); sc[:, 2] = rng.integers(0, 17, G * Sn)
    counts = rng.integers(0, Sn + 1, N).astype(np.int32); counts[0] = Sn; counts[1] = 0
    len_h = rng.integers(1, 8, N * Sn).astype(np.int32)
    len_s = rng.integers(1, 8, G).astype(np.int32)
    pair_h = np.tile((np.arange(N)[:, None] * Sn + np.arange(Sn)[None, :]).reshape(-1), K).astype(np.int32)
    pair_s = np.repeat(np.arange(G), Sn).astype(np.int32)
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)
    acc = S.new_accumulator(dev)
    args = dict(n_images=N, group_count=t(counts), pair_h=t(pair_h), pair_s=t(pair_s), len_h=t(len_h), len_s=t(len_s),
                min_len_valid=3)
    tab, rew, gv = S.reduce_pairs(t(sc), Sn, acc=acc, **args)
    first = acc[:16].clone()
    acc2 = S.new_accumulator(dev)
    tab2, rew2, gv2 = S.reduce_pairs(t(sc), Sn, acc=acc2, **args)
    assert torch.equal(first, acc2[:16]) and torch.equal(torch.nan_to_num(tab), torch.nan_to_num(tab2))   # reproducible
    S.reduce_pairs(t(sc), Sn, acc=acc, **args)                       # accumulates over calls
    np.testing.assert_allclose(acc[:16].cpu().numpy(), 2 * first.cpu().numpy(), rtol=1e-14)
    tab, rew, gv, a = tab.cpu().numpy(), rew.cpu().numpy(), gv.cpu().numpy(), first.cpu().numpy()
    tabk, _, _ = S.reduce_pairs(t(sc), Sn, mean_over_kept=True, **args)
    tabk = tabk.cpu().numpy()
    rows_all, best = [], []
    for g in range(G):
        img = g % N
        c = int(counts[img])
        rows = sc[g * Sn:g * Sn + c]
        ok = (len_h[pair_h[g * Sn:g * Sn + c]] >= 3) & (len_s[g] >= 3)
        if c:
            rows_all.append(rows); best.append([rows[:, 2].min(), rows[:, 3].max()])
        if not ok.any():
            assert np.isnan(tab[g]).all() and np.isnan(rew[g]) and gv[g] == 0
            continue
        r = rows[ok]
        exp = np.array([r[:, 1].sum() / c, r[:, 0].sum() / c, r[:, 2].sum() / c, r[:, 3].sum() / c, r[:, 2].min(),
                        r[:, 3].max()], dtype=np.float32)
        np.testing.assert_array_equal(tab[g, 5:], exp)
        np.testing.assert_allclose(tabk[g, 5:9], exp[:4] * c / ok.sum(), rtol=1e-6)
        assert gv[g] == 1 and (tab[g, :5] == 0).all()
        assert rew[g] == pytest.approx(2 / (1 / float(exp[0]) + 1 / float(exp[1])), rel=1e-12)
    rows_all, best = np.concatenate(rows_all, 0), np.array(best)
    np.testing.assert_allclose(a[0:4], rows_all.sum(0), rtol=1e-12)
    np.testing.assert_allclose(a[4:8], (rows_all ** 2).sum(0), rtol=1e-12)
    np.testing.assert_allclose(a[8:10], best.sum(0), rtol=1e-12)
    np.testing.assert_allclose(a[10:12], (best ** 2).sum(0), rtol=1e-12)
    assert a[12] == len(rows_all) and a[13] == len(best) and a[14] == gv.sum()
