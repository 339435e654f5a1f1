"""The SCST / loss oracle against the reference's own functions + torch autograd (tests/golden/scst.npz,
recorded by tests/golden/make_scst_goldens.py from OSIE/models/loss.py and OSIE/train.py:223-258)."""
import os

import numpy as np
import pytest

from oracle import scst as OS

TRIALS = ["m1_t0_", "m1_t1_", "m1_t2_", "m2_t0_", "m2_t1_", "m2_t2_"]


def load(golden_dir):
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    s = np.load(os.path.join(golden_dir, "scst.npz"))
    stack = lambda key: np.stack([g[t + key] for t in TRIALS], 0)
    return g, s, stack("actions"), stack("dur"), stack("action_mask"), stack("duration_mask")


def test_scst_loss_and_gradients(golden_dir):
    g, s, actions, dur, am, dm = load(golden_dir)
    r = OS.scst_loss(g["probs"], g["mu"], g["sigma2"], actions, dur, am, dm, s["table"], 4)
    assert list(r["used"]) == list(s["used"]) == [0, 1, 3, 4]         # trial 2 carries a NaN row: rejected
    for key in ("loss", "loss_actions", "loss_duration"):
        assert float(r[key]) == pytest.approx(float(s[key]), rel=2e-5), key
    np.testing.assert_allclose(r["advantage"], s["advantage"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r["neg_log_actions"], s["neg_log_actions"], rtol=1e-5)
    np.testing.assert_allclose(r["neg_log_durations"], s["neg_log_durations"], rtol=1e-5)
    for key in ("grad_probs", "grad_mu", "grad_sigma2"):
        ref = s[key]
        np.testing.assert_allclose(r[key], ref, rtol=2e-5, atol=1e-5 * np.abs(ref).max(), err_msg=key)


def test_supervised_loss_gradients(golden_dir):
    g, s, *_ = load(golden_dir)
    A = g["loss_logits"].shape[-1]
    gt = np.zeros_like(g["loss_logits"])
    np.put_along_axis(gt, g["loss_gt_idx"][..., None], 1.0, -1)
    gt[0, 0] = s["ce_gt00"]
    loss, grad = OS.cross_entropy_grad(g["loss_logits"], gt, g["loss_mask"], upstream=1.7)
    assert loss == pytest.approx(float(s["ce"]), rel=1e-5)
    np.testing.assert_allclose(grad, s["ce_grad_logits"], rtol=1e-4, atol=1e-6 * np.abs(s["ce_grad_logits"]).max())
    loss, gmu, gs2 = OS.lognormal_nll_grad(g["mu"], g["sigma2"], g["loss_gt_dur"], g["loss_mask"], upstream=0.6)
    assert loss == pytest.approx(float(s["nll"]), rel=1e-5)
    np.testing.assert_allclose(gmu, s["nll_grad_mu"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(gs2, s["nll_grad_sigma2"], rtol=1e-4, atol=1e-6)
    assert A == 1201
