"""End-to-end pipeline (decode -> sample -> prep -> score -> reduce) against the oracle:
the sampled scanpaths are read back and re-scored / re-aggregated on the CPU."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(task, N, K, Sn, T, wave, seed=0):
    from scanpaths_b200 import build
    build.build_library()
    from scanpaths_b200.pipeline import ScanpathPipeline
    from scanpaths_b200.weights import random_state_dict, synthetic_features
    sys_path_hack = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path_hack)
    from bench import synth_humans
    dev = torch.device("cuda")
    pipe = ScanpathPipeline(random_state_dict(task, seed), task, T, K, 1, dev, wave, seed=77)
    hx, hl = synth_humans(N, Sn, 5, lo=2, hi=12)
    pipe.set_humans(hx, hl)
    if task == "OSIE":
        vf, att = synthetic_features(N, seed), None
    else:
        vf, att = synthetic_features(N, seed, attention=True)
    return pipe, vf, att, hx, hl


@pytest.mark.parametrize("task,wave", [("OSIE", 2), ("AiR", 3)])
def test_pipeline_table_and_metrics_match_oracle(task, wave):
    from oracle import c_scoring as CO
    from oracle import scoring as O
    from golden.make_goldens import to_struct
    N, K, Sn, T = 5, 4, 6, 5
    pipe, vf, att, hx, hl = _setup(task, N, K, Sn, T, wave)
    out = pipe.run(vf.cuda(), None if att is None else att.cuda(), keep_scores=True, keep_paths=True, valid_min_len=3)
    torch.cuda.synchronize()
    HD = pipe.decoder.heads
    assert out["table"].shape == (HD, K, N, 11) and out["reward"].shape == (HD, K, N)
    humans = [[to_struct(hx[i, s, :hl[i, s]]) for s in range(Sn)] for i in range(N)]
    for hd, n0, n1, smp in out["paths"]:
        n = n1 - n0
        xyd = smp["xyd"].cpu().numpy(); lens = smp["len"].cpu().numpy()
        # 1. raw scores of this wave vs the C oracle on the very same sampled paths
        ha = hx.reshape(N * Sn, -1, 3).copy(); ha[..., 2] *= 1000.0
        pa = xyd.copy(); pa[..., 2] *= 1000.0
        gi = np.array([(n0 + i) * Sn + s for k in range(K) for i in range(n) for s in range(Sn)])
        pi = np.array([k * n + i for k in range(K) for i in range(n) for s in range(Sn)])
        ref = CO.score_pairs(ha, hl.reshape(-1), pa, lens, gi, pi).reshape(K, n, Sn, 4)
        got = out["scores"][hd, :, n0:n1].cpu().numpy()
        assert np.array_equal(got[..., :3], ref[..., :3], equal_nan=True)
        np.testing.assert_allclose(got[..., 3], ref[..., 3], rtol=1e-12)
        # 2. the reduced table vs the oracle's pairs_eval on the same lists (per sample k)
        for k in range(K):
            preds = [to_struct(xyd[k * n + i, :lens[k * n + i]]) for i in range(n)]
            pe = O.pairs_eval(humans[n0:n1], preds)
            tab = out["table"][hd, k, n0:n1].cpu().numpy().astype(np.float64)
            np.testing.assert_allclose(tab[:, 5:], pe[:, 5:], rtol=1e-6, equal_nan=True)
            rew = out["reward"][hd, k, n0:n1].cpu().numpy()
            exp = 2.0 / (1.0 / pe[:, 5] + 1.0 / pe[:, 6])
            np.testing.assert_allclose(rew, exp, rtol=1e-6, equal_nan=True)
    # 3. the aggregate equals evaluation() over all (sample, image) entries of head 0
    all_gt, all_pred = [], []
    for hd, n0, n1, smp in out["paths"]:
        if hd != 0:
            continue
        xyd = smp["xyd"].cpu().numpy(); lens = smp["len"].cpu().numpy()
        for k in range(K):
            for i in range(n1 - n0):
                all_gt.append(humans[n0 + i]); all_pred.append(to_struct(xyd[k * (n1 - n0) + i, :lens[k * (n1 - n0) + i]]))
    m_ref, s_ref, _ = O.evaluation(all_gt, all_pred)
    m, s = pipe.metrics(out, head=0)
    for grp in ("ScanMatch", "VAME"):
        for key in m_ref[grp]:
            assert m[grp][key] == pytest.approx(m_ref[grp][key], rel=1e-10), key
            assert s[grp][key] == pytest.approx(s_ref[grp][key], rel=1e-6, abs=1e-9), key


def test_pipeline_coco_tasks_and_host_input():
    """COCO variant (per-image 5x5 weights by task id) and pinned-host input features."""
    N, K, Sn, T = 4, 3, 4, 3
    pipe, vf, att, hx, hl = _setup("COCO_Search18", N, K, Sn, T, wave=4)
    tasks = torch.tensor([0, 17, 5, 5])
    a = pipe.run(vf.cuda(), att.cuda(), tasks, keep_scores=True)
    pipe.sampler._calls = 0                                # same Philox streams again
    b = pipe.run(vf.pin_memory(), att.pin_memory(), tasks, keep_scores=True)
    assert torch.equal(a["scores"], b["scores"])
    assert torch.equal(torch.nan_to_num(a["table"], nan=-1.0), torch.nan_to_num(b["table"], nan=-1.0))
    ok = a["group_valid"].bool()
    assert (a["table"][ok][:, :5] == 0).all()               # MultiMatch slots: placeholder (out of scope)
    assert torch.isnan(a["table"][~ok]).all() and torch.isnan(a["reward"][~ok]).all()
    assert torch.isfinite(a["reward"][ok]).all()


def test_pipeline_ragged_subject_counts_and_new_humans():
    """Images with fewer subjects (pack_subject_lists pads them with empty scanpaths): the padded subjects enter
    neither the table nor the reward nor the aggregate, and the table divides by the real count (len(gt),
    OSIE/utils/evaluation.py:329).  A second set_humans() with another subject count must not reuse a pair map."""
    from oracle import scoring as O
    from golden.make_goldens import to_struct
    from scanpaths_b200 import scoring as S
    N, K, T = 4, 3, 6
    pipe, vf, att, hx, hl = _setup("OSIE", N, K, 7, T, wave=4)
    counts = [7, 3, 5, 1]
    lists = [[to_struct(hx[i, s, :hl[i, s]]) for s in range(counts[i])] for i in range(N)]
    xyd, lens, nsub = S.pack_subject_lists(lists)
    assert tuple(xyd.shape[:2]) == (N, 7) and nsub.tolist() == counts
    pipe.set_humans(xyd, lens, nsub)
    out = pipe.run(vf.cuda(), keep_paths=True)
    torch.cuda.synchronize()
    (hd, n0, n1, smp), = out["paths"]
    p_xyd, p_len = smp["xyd"].cpu().numpy(), smp["len"].cpu().numpy()
    rows, best = [], []
    for k in range(K):
        preds = [to_struct(p_xyd[k * N + i, :p_len[k * N + i]]) for i in range(N)]
        pe = O.pairs_eval(lists, preds)
        tab = out["table"][0, k].cpu().numpy().astype(np.float64)
        np.testing.assert_allclose(tab[:, 5:], pe[:, 5:], rtol=1e-6, equal_nan=True)
        assert np.array_equal(np.isnan(tab[:, 0]), np.isnan(pe[:, 5]))
        assert np.array_equal(out["group_valid"][0, k].cpu().numpy() == 0, np.isnan(pe[:, 5]))
        for i in range(N):
            sc = np.array([O.score_pair(O.structured_to_array(g), O.structured_to_array(preds[i])) for g in lists[i]], dtype=np.float64)
            rows.append(sc); best.append([sc[:, 2].min(), sc[:, 3].max()])
    allrows, best = np.concatenate(rows, 0), np.array(best)
    m, s = pipe.metrics(out)
    assert out["acc"][0, 12].item() == K * sum(counts) and out["acc"][0, 13].item() == K * N
    assert m["ScanMatch"]["with duration"] == pytest.approx(allrows[:, 0].mean(), rel=1e-10)
    assert m["VAME"]["SED"] == pytest.approx(allrows[:, 2].mean(), rel=1e-10)
    assert m["VAME"]["STDE"] == pytest.approx(allrows[:, 3].mean(), rel=1e-10)
    assert m["VAME"]["SED_best"] == pytest.approx(best[:, 0].mean(), rel=1e-10)
    assert s["VAME"]["STDE_best"] == pytest.approx(best[:, 1].std(), rel=1e-6, abs=1e-9)
    # other humans, other subject count: fresh pair map, scores equal a fresh pipeline's
    hx2, hl2 = hx[:, :2], hl[:, :2]
    pipe.set_humans(hx2, hl2)
    pipe.sampler._calls = 0
    again = pipe.run(vf.cuda(), keep_scores=True)
    fresh, _, _, _, _ = _setup("OSIE", N, K, 7, T, wave=4)
    fresh.set_humans(hx2, hl2)
    ref = fresh.run(vf.cuda(), keep_scores=True)
    assert torch.equal(again["scores"], ref["scores"]) and again["scores"].shape == (1, K, N, 2, 4)


def test_pipeline_overlap_and_checksums_bench_shape():
    """Bench-shaped slice (768 images = 3 waves x 64 samples x 15 subjects, 737,280 pairs): the tail of wave w on
    the side stream under the decode of wave w+1 gives bit-identical results to the single-stream order, pinned
    host input equals device input, and the reduced table is consistent with the device accumulators
    (a checksum of checksums: sum over groups of the table's SED / STDE means x S == the accumulated sums)."""
    from scanpaths_b200.pipeline import ScanpathPipeline
    from scanpaths_b200.weights import random_state_dict
    from bench import synth_humans
    dev = torch.device("cuda")
    N, K, Sn, T = 768, 64, 15, 16
    g = torch.Generator(device=dev).manual_seed(3)
    vf = torch.randn((N, 512, 30, 40), generator=g, device=dev).clamp_min_(0)
    hx, hl = synth_humans(N, Sn, 9)
    outs = []
    for overlap, host in ((True, False), (False, False), (True, True)):
        pipe = ScanpathPipeline(random_state_dict("OSIE", 0), "OSIE", T, K, 1, dev, 256, seed=5, overlap_tail=overlap)
        pipe.set_humans(hx, hl)
        src = vf.cpu().pin_memory() if host else vf
        o = pipe.run(src, valid_min_len=0)
        torch.cuda.synchronize()
        outs.append({k: o[k].clone() for k in ("table", "reward", "group_valid", "acc")})
        del pipe
    for other in outs[1:]:
        for k in ("table", "reward", "group_valid", "acc"):
            assert torch.equal(torch.nan_to_num(outs[0][k].double(), nan=-1.0), torch.nan_to_num(other[k].double(), nan=-1.0)), k
    tab, acc = outs[0]["table"][0].double(), outs[0]["acc"][0]
    assert acc[12].item() == N * K * Sn and acc[13].item() == N * K and acc[14].item() == N * K
    for slot, a in ((5, 1), (6, 0), (7, 2), (8, 3)):                  # table means x S vs accumulated sums (f32 rows)
        assert (tab[..., slot].sum() * Sn).item() == pytest.approx(acc[a].item(), rel=2e-6)
    assert tab[..., 9].sum().item() == pytest.approx(acc[8].item(), rel=1e-6)       # SED best
    assert tab[..., 10].sum().item() == pytest.approx(acc[9].item(), rel=1e-6)      # STDE best
    assert torch.isfinite(outs[0]["reward"]).all() and (outs[0]["reward"] > 0).all()
