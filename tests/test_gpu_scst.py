"""a14 + f4 on the GPU: the CUDA log-likelihood / loss kernels (csrc/loss.cu) and the fused SCST tail against
the reference's own functions + torch autograd (tests/golden/scst.npz) -- gate 1e-5 relative on the loss and
on the gradients w.r.t. all_actions_prob, log_normal_mu, log_normal_sigma2 -- and the SCST batch driver
(sample -> score -> reward -> loss) against the oracle, incl. the trial-rejection loop of train.py:223-239."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TRIALS = ["m1_t0_", "m1_t1_", "m1_t2_", "m2_t0_", "m2_t1_", "m2_t2_"]
RTOL = 1e-5


@pytest.fixture(scope="module")
def data(golden_dir):
    from scanpaths_b200 import build
    build.build_library()
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    s = np.load(os.path.join(golden_dir, "scst.npz"))
    dev = torch.device("cuda")
    stack = lambda key, dt=None: torch.tensor(np.stack([g[t + key] for t in TRIALS], 0), device=dev, dtype=dt)
    return g, s, dev, stack


def _close(got, ref, what, rtol=RTOL):
    got = got.detach().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max(), err_msg=what)


def test_scst_loss_matches_reference_autograd(data):
    from scanpaths_b200.scst import scst_loss
    g, s, dev, stack = data
    probs = torch.tensor(g["probs"], device=dev, requires_grad=True)
    mu = torch.tensor(g["mu"], device=dev, requires_grad=True)
    s2 = torch.tensor(g["sigma2"], device=dev, requires_grad=True)
    samples = {"selected_actions": stack("actions", torch.int32), "durations": stack("dur"),
               "action_masks": stack("action_mask"), "duration_masks": stack("duration_mask")}
    table = torch.tensor(s["table"], device=dev)
    # reward + per-image validity the way the pipeline's reduction delivers them
    a, b = table[..., 5].double(), table[..., 6].double()
    reward = 2.0 / (1.0 / a + 1.0 / b)
    group_valid = (~torch.isnan(table).any(-1)).to(torch.uint8)
    loss, aux = scst_loss(probs, mu, s2, samples, reward, group_valid, k_use=4)
    assert aux["trial_used"].cpu().tolist() == [1, 1, 0, 1, 1, 0] and int(aux["n_used"]) == 4
    used = list(s["used"])
    assert float(loss) == pytest.approx(float(s["loss"]), rel=RTOL)
    assert float(aux["loss_actions"]) == pytest.approx(float(s["loss_actions"]), rel=RTOL)
    assert float(aux["loss_duration"]) == pytest.approx(float(s["loss_duration"]), rel=RTOL)
    _close(aux["advantage"][used], s["advantage"], "advantage")
    assert (aux["advantage"][[2, 5]] == 0).all()
    _close(aux["neg_log_actions"][used], s["neg_log_actions"], "neg_log_actions")
    _close(aux["neg_log_durations"][used], s["neg_log_durations"], "neg_log_durations")
    loss.backward()
    _close(probs.grad, s["grad_probs"], "d loss / d all_actions_prob")
    _close(mu.grad, s["grad_mu"], "d loss / d log_normal_mu")
    _close(s2.grad, s["grad_sigma2"], "d loss / d log_normal_sigma2")
    nz = probs.grad != 0
    assert int(nz.sum()) <= 4 * probs.shape[0] * probs.shape[1]          # only sampled actions carry gradient


def test_scst_loss_parts_backward(data):
    """loss = loss_actions + loss_duration: back-propagating the parts separately sums to the gradient of loss."""
    from scanpaths_b200.scst import scst_loss
    g, s, dev, stack = data
    samples = {"selected_actions": stack("actions", torch.int32), "durations": stack("dur"),
               "action_masks": stack("action_mask"), "duration_masks": stack("duration_mask")}
    table = torch.tensor(s["table"], device=dev)
    reward = 2.0 / (1.0 / table[..., 5].double() + 1.0 / table[..., 6].double())
    gv = (~torch.isnan(table).any(-1)).to(torch.uint8)
    grads = []
    for which in ("loss", "parts"):
        probs = torch.tensor(g["probs"], device=dev, requires_grad=True)
        mu = torch.tensor(g["mu"], device=dev, requires_grad=True)
        s2 = torch.tensor(g["sigma2"], device=dev, requires_grad=True)
        loss, aux = scst_loss(probs, mu, s2, samples, reward, gv, k_use=4)
        (loss if which == "loss" else aux["loss_actions"] + aux["loss_duration"]).backward()
        grads.append((probs.grad, mu.grad, s2.grad))
    for a, b in zip(*grads):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-9)


def test_log_action_duration_mirror(data):
    from scanpaths_b200.models.loss import LogAction, LogDuration
    g, s, dev, stack = data
    tag = TRIALS[0]
    t = lambda k, rg=False: torch.tensor(g[k], device=dev, requires_grad=rg)
    p, x = t(tag + "sel_prob", True), t(tag + "dur", True)
    mu, s2 = t("mu", True), t("sigma2", True)
    am, dm = t(tag + "action_mask"), t(tag + "duration_mask")
    la, ld = LogAction(p, am), LogDuration(x, mu, s2, dm)
    _close(la, g[tag + "log_action"], "LogAction")
    _close(ld, g[tag + "log_duration"], "LogDuration")
    w = torch.tensor(s["row_w"], device=dev)
    ((la * w).sum() + (ld * w).sum()).backward()
    _close(p.grad, s["la_grad_p"], "dLogAction/dp")
    _close(mu.grad, s["ld_grad_mu"], "dLogDuration/dmu")
    _close(s2.grad, s["ld_grad_sigma2"], "dLogDuration/dsigma2")
    _close(x.grad, s["ld_grad_x"], "dLogDuration/dx")
    for tg in TRIALS[1:]:                                            # every recorded trial, forward
        _close(LogAction(t(tg + "sel_prob"), t(tg + "action_mask")), g[tg + "log_action"], tg)
        _close(LogDuration(t(tg + "dur"), t("mu"), t("sigma2"), t(tg + "duration_mask")), g[tg + "log_duration"], tg)


def test_supervised_losses_mirror(data):
    from scanpaths_b200.models.loss import CrossEntropyLoss, MLPLogNormalDistribution
    g, s, dev, _ = data
    logits = torch.tensor(g["loss_logits"], device=dev, requires_grad=True)
    gt = torch.zeros_like(logits).scatter_(2, torch.tensor(g["loss_gt_idx"], device=dev).unsqueeze(-1), 1.0).detach()
    gt[0, 0] = torch.tensor(s["ce_gt00"], device=dev)
    mask = torch.tensor(g["loss_mask"], device=dev)
    ce = CrossEntropyLoss(logits, gt, mask)
    assert float(ce) == pytest.approx(float(s["ce"]), rel=RTOL) 
    (ce * 1.7).backward()
    _close(logits.grad, s["ce_grad_logits"], "dCE/dlogits", rtol=2e-5)
    mu = torch.tensor(g["mu"], device=dev, requires_grad=True)
    s2 = torch.tensor(g["sigma2"], device=dev, requires_grad=True)
    nll = MLPLogNormalDistribution(mu, s2, torch.tensor(g["loss_gt_dur"], device=dev), mask)
    assert float(nll) == pytest.approx(float(s["nll"]), rel=RTOL)
    assert float(nll) == pytest.approx(float(g["loss_lognormal"]), rel=RTOL)
    (nll * 0.6).backward()
    _close(mu.grad, s["nll_grad_mu"], "dNLL/dmu")
    _close(s2.grad, s["nll_grad_sigma2"], "dNLL/dsigma2")
    with pytest.raises(Exception):
        CrossEntropyLoss(logits.cpu(), gt.cpu(), mask.cpu())            # no CPU path


def _humans(N, Sn, seed, short=()):
    from golden.make_goldens import to_struct
    rng = np.random.default_rng(seed)
    out = []
    for i in range(N):
        subs = []
        for j in range(Sn - (i % 2)):                                   # ragged subject counts
            L = 2 if (i, j) in short else int(rng.integers(4, 12))
            subs.append(to_struct(np.stack([rng.uniform(0, 320, L), rng.uniform(0, 240, L),
                                            np.exp(rng.normal(np.log(0.25), 0.4, L))], 1)))
        out.append(subs)
    return out


def test_scst_step_rewards_match_oracle_and_reject_short_trials(data):
    """The SCST batch on the device (N = 4 images, K = 5 trials, +3 spare): rewards equal the oracle's
    pairs_eval -> hmean on the very same samples; trials in which an image's prediction is shorter than 3
    fixations are rejected exactly as train.py:237-238 does; the loss equals the oracle's on those samples."""
    from oracle import scoring as O
    from oracle import scst as OS
    from golden.make_goldens import to_struct
    from scanpaths_b200.models.sampling import Sampling
    from scanpaths_b200.scst import ScstRewardStep
    g, s, dev, _ = data
    N, K = 4, 5
    p_np = g["probs"][:N].copy()
    p_np[1, 1] *= 0.7 / (1.0 - p_np[1, 1, 0]); p_np[1, 1, 0] = 0.3           # image 1 stops after ONE fixation in ~30 % of
    probs = torch.tensor(p_np, device=dev, requires_grad=True)               # the trials: those trials must be rejected
    mu = torch.tensor(g["mu"][:N], device=dev, requires_grad=True)
    s2 = torch.tensor(g["sigma2"][:N], device=dev, requires_grad=True)
    humans = _humans(N, 5, 3, short={(0, 1)})
    step = ScstRewardStep(Sampling(convLSTM_length=16, min_length=1, seed=5), dev, rl_sample_number=K, spare=3)
    step.set_humans(humans)
    loss, aux = step(probs, mu, s2)
    smp = aux["samples"]
    xyd, lens = smp["xyd"].cpu().numpy(), smp["len"].cpu().numpy()
    KT = K + 3
    table = np.stack([O.pairs_eval(humans, [to_struct(xyd[k * N + i, :lens[k * N + i]]) for i in range(N)])
                      for k in range(KT)], 0)
    got_tab = aux["table"].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got_tab[..., 5:], table[..., 5:], rtol=1e-6, equal_nan=True)
    rejected = np.isnan(table[..., 5]).any(1)
    assert rejected.any() and (~rejected).sum() >= K, "the case must exercise the rejection rule"
    used = aux["trial_used"].cpu().numpy().astype(bool)
    assert used.tolist() == [bool(u) for u in np.isin(np.arange(KT), np.flatnonzero(~rejected)[:K])]
    tab32 = table.astype(np.float32)
    tab32[..., :5] = np.where(np.isnan(table[..., 5:6]), np.nan, 0.0)        # MultiMatch slots: placeholder
    r = OS.scst_loss(probs.detach().cpu().numpy(), mu.detach().cpu().numpy(), s2.detach().cpu().numpy(),
                     smp["selected_actions"].cpu().numpy(), smp["durations"].cpu().numpy(),
                     smp["action_masks"].cpu().numpy(), smp["duration_masks"].cpu().numpy(), tab32, K)
    assert float(loss) == pytest.approx(float(r["loss"]), rel=RTOL, abs=1e-7)
    loss.backward()
    _close(probs.grad, r["grad_probs"], "grad probs")
    _close(mu.grad, r["grad_mu"], "grad mu")
    _close(s2.grad, r["grad_sigma2"], "grad sigma2")


def test_train_loop_shape_with_mirror_api(data):
    """The reference's own SCST loop shape (OSIE/train.py:223-250) on the mirror modules: `pairs_eval` returns
    NaN only where the reference does, so `np.any(np.isnan(metrics_reward))` accepts ordinary trials and the
    loop terminates (with NaN MultiMatch slots it would spin forever)."""
    from scanpaths_b200.models.loss import LogAction, LogDuration
    from scanpaths_b200.models.sampling import Sampling
    from scanpaths_b200.utils.evaluation import pairs_eval
    g, s, dev, _ = data
    N = 3
    probs = torch.tensor(g["probs"][3:3 + N], device=dev, requires_grad=True)
    mu = torch.tensor(g["mu"][3:3 + N], device=dev, requires_grad=True)
    s2 = torch.tensor(g["sigma2"][3:3 + N], device=dev, requires_grad=True)
    humans = _humans(N, 4, 8)
    sampling = Sampling(convLSTM_length=16, min_length=1, seed=3)
    images = torch.zeros(N, 3, 2, 2, device=dev)
    trial, attempts, rewards, nla, nld = 0, 0, [], [], []
    while True:
        if trial >= 3:
            break
        attempts += 1
        assert attempts < 200, "the rejection loop does not terminate"
        samples = sampling.random_sample(probs, mu, s2)
        fix, am, dm = sampling.generate_scanpath(images, samples["selected_actions_probs"], samples["durations"],
                                                 samples["selected_actions"])
        t = samples["durations"].data.clone()
        metrics_reward = pairs_eval(humans, fix, None, None)
        if np.any(np.isnan(metrics_reward)):
            continue
        trial += 1
        rewards.append(torch.tensor(metrics_reward, dtype=torch.float32, device=dev).unsqueeze(0))
        nla.append(-LogAction(samples["selected_actions_probs"], am).unsqueeze(0))
        nld.append(-LogDuration(t, mu, s2, dm).unsqueeze(0))
    rw = torch.cat(rewards, 0)
    hm = 2.0 / (1.0 / rw[:, :, 5] + 1.0 / rw[:, :, 6])
    adv = hm - hm.mean(0, keepdim=True)
    loss = (torch.cat(nla, 0) * adv).sum() + (torch.cat(nld, 0) * adv).sum()
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(probs.grad).all() and (probs.grad != 0).any()
    assert torch.isfinite(mu.grad).all() and torch.isfinite(s2.grad).all()


def test_air_scst_step_matches_oracle(data):
    """AiR's SCST batch: same / diff rewards of both heads equal the oracle's
    pairs_eval_scanmatch_performance_related on the very same samples, and the loss equals the reference's
    effective expression (AiR/train.py:286-340: the lambda_5 lines are no-op statements) evaluated by the
    SCST oracle per head."""
    from oracle import scoring as O
    from oracle import scst as OS
    from golden.make_goldens import to_struct
    from scanpaths_b200.models.sampling import Sampling
    from scanpaths_b200.scst import AirScstStep
    g, s, dev, _ = data
    N, K = 4, 3
    rng = np.random.default_rng(2)
    humans = _humans(N, 5, 13)
    perf = [[bool(b) for b in rng.integers(0, 2, len(h))] for h in humans]
    perf[2] = [True] * len(humans[2])                       # image 2 has no incorrect answerers: diff group empty
    predict = {}
    for name, lo in (("good", 0), ("poor", 2)):
        predict[name + "_all_actions_prob"] = torch.tensor(g["probs"][lo:lo + N], device=dev, requires_grad=True)
        predict[name + "_log_normal_mu"] = torch.tensor(g["mu"][lo:lo + N], device=dev, requires_grad=True)
        predict[name + "_log_normal_sigma2"] = torch.tensor(g["sigma2"][lo:lo + N], device=dev, requires_grad=True)
    step = AirScstStep(Sampling(convLSTM_length=16, min_length=1, seed=11), dev, rl_sample_number=K)
    step.set_humans(humans, perf)
    loss, aux = step(predict)
    expect = 0.0
    for name, given in (("good", True), ("poor", False)):
        a = aux[name]
        smp = a["samples"]
        xyd, lens = smp["xyd"].cpu().numpy(), smp["len"].cpu().numpy()
        table = np.zeros((K, N, 11), dtype=np.float32)
        for k in range(K):
            preds = [to_struct(xyd[k * N + i, :lens[k * N + i]]) for i in range(N)]
            same, diff, flag = O.pairs_eval_scanmatch_performance_related(humans, preds, perf, given)
            assert flag
            np.testing.assert_allclose(a["same_table"][k].cpu().numpy(), same, rtol=1e-6, equal_nan=True)
            np.testing.assert_allclose(a["diff_table"][k].cpu().numpy(), diff, rtol=1e-6, equal_nan=True)
            table[k, :, 5:7] = np.nan_to_num(same)             # train.py:283: NaN -> 0, then hmean (:300)
        empty = "diff" if given else "same"                    # image 2: everybody answered correctly
        assert np.isnan(a[empty + "_table"][:, 2].cpu().numpy()).all() and (a[empty][:, 2] == 0).all()
        r = OS.scst_loss(predict[name + "_all_actions_prob"].detach().cpu().numpy(),
                         predict[name + "_log_normal_mu"].detach().cpu().numpy(),
                         predict[name + "_log_normal_sigma2"].detach().cpu().numpy(),
                         smp["selected_actions"].cpu().numpy(), smp["durations"].cpu().numpy(),
                         smp["action_masks"].cpu().numpy(), smp["duration_masks"].cpu().numpy(), table, K)
        np.testing.assert_allclose(a["same"].cpu().numpy(), OS.hmean_reward(table), rtol=1e-6)
        expect += float(r["loss"])
        aux[name]["oracle"] = r
    assert float(loss) == pytest.approx(expect, rel=RTOL, abs=1e-7)
    loss.backward()
    for name in ("good", "poor"):
        _close(predict[name + "_all_actions_prob"].grad, aux[name]["oracle"]["grad_probs"], name + " grad probs")
        _close(predict[name + "_log_normal_mu"].grad, aux[name]["oracle"]["grad_mu"], name + " grad mu")
