"""CPU oracle: the self-critical (SCST) reward / loss tail (TEST INFRASTRUCTURE ONLY).

numpy float64 restatement of OSIE/train.py:223-258 -- trial rejection on a NaN pairs_eval table (:237-238),
LogAction / LogDuration per accepted trial (:242-243, models/loss.py:34-45), harmonic-mean reward of table
slots 5, 6 (:252), mean-over-trials baseline (:254), loss (:256-258) -- with the analytic gradients w.r.t.
all_actions_prob (through the gather of models/sampling.py:23-24), log_normal_mu and log_normal_sigma2, and of
the supervised losses (loss.py:10-32).  Pinned by tests/golden/scst.npz: the reference's own functions and
torch autograd run on the recorded samples (tests/golden/make_scst_goldens.py).
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-7


def _logpdf(x, mu, s2):
    return np.log(1.0 / (x + EPS) * 1.0 / np.sqrt(2 * math.pi * s2)) - (np.log(x + EPS) - mu) ** 2 / (2 * s2)


def accepted_trials(table, k_use):
    """Indices of the first k_use trials whose [N,11] table has no NaN (train.py:237-238)."""
    used = [k for k in range(table.shape[0]) if not np.any(np.isnan(table[k]))]
    return used[:k_use]


def hmean_reward(table):
    """scipy.stats.hmean over slots 5, 6 of the float32 table (train.py:252); 0 if either is 0."""
    a, b = table[..., 5].astype(np.float64), table[..., 6].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = 2.0 / (1.0 / a + 1.0 / b)
    return np.where((a > 0) & (b > 0), r, 0.0)


def scst_loss(probs, mu, s2, actions, dur, am, dm, table, k_use):
    """probs [N,T,A], mu/s2 [N,T]; actions/dur/am/dm [K,N,T]; table [K,N,11].
    Returns dict(loss, loss_actions, loss_duration, used, advantage [len(used),N], neg_log_actions,
    neg_log_durations, grad_probs, grad_mu, grad_sigma2)."""
    probs, mu, s2 = (np.asarray(v, np.float64) for v in (probs, mu, s2))
    used = accepted_trials(table, k_use)
    N, T, A = probs.shape
    n_idx, t_idx = np.meshgrid(np.arange(N), np.arange(T), indexing="ij")
    r = hmean_reward(table[used])
    adv = r - r.mean(0, keepdims=True)
    nla, nld = [], []
    gp, gmu, gs2 = np.zeros_like(probs), np.zeros_like(mu), np.zeros_like(s2)
    for j, k in enumerate(used):
        a, x = np.asarray(actions[k]), np.asarray(dur[k], np.float64)
        ma, md = np.asarray(am[k], np.float64), np.asarray(dm[k], np.float64)
        p = probs[n_idx, t_idx, a]
        nla.append(-(np.log(p + EPS) * ma).sum(-1) / ma.sum())
        nld.append(-(_logpdf(x, mu, s2) * md).sum(-1) / md.sum())
        w = -adv[j][:, None]                                  # d loss / d Log*[k, n]
        np.add.at(gp, (n_idx, t_idx, a), w * ma / (ma.sum() * (p + EPS)))
        d = np.log(x + EPS) - mu
        gmu += w * md / md.sum() * d / s2
        gs2 += w * md / md.sum() * (-0.5 / s2 + d * d / (2 * s2 * s2))
    nla, nld = np.array(nla), np.array(nld)
    la, ld = (nla * adv).sum(), (nld * adv).sum()
    return dict(loss=la + ld, loss_actions=la, loss_duration=ld, used=used, advantage=adv, neg_log_actions=nla,
                neg_log_durations=nld, grad_probs=gp, grad_mu=gmu, grad_sigma2=gs2)


def cross_entropy_grad(logits, gt, mask, upstream=1.0):
    """(loss, d loss / d logits) of CrossEntropyLoss (loss.py:10-14) with a dense target."""
    z = np.asarray(logits, np.float64)
    gt, mask = np.asarray(gt, np.float64), np.asarray(mask, np.float64)
    e = np.exp(z - z.max(-1, keepdims=True))
    p = e / e.sum(-1, keepdims=True)
    loss = -(gt * np.log(p + EPS) * mask[..., None]).sum() / mask.sum()
    w = gt * p / (p + EPS)
    grad = -(mask[..., None] / mask.sum()) * (w - p * w.sum(-1, keepdims=True)) * upstream
    return loss, grad


def lognormal_nll_grad(mu, s2, gt, mask, upstream=1.0):
    """(loss, d/d mu, d/d sigma2) of MLPLogNormalDistribution (loss.py:27-32)."""
    mu, s2, gt, mask = (np.asarray(v, np.float64) for v in (mu, s2, gt, mask))
    on = mask == 1
    loss = -_logpdf(gt, mu, s2)[on].sum() / mask.sum()
    d = np.log(gt + EPS) - mu
    gmu = np.where(on, -(d / s2), 0.0) / mask.sum() * upstream
    gs2 = np.where(on, -(-0.5 / s2 + d * d / (2 * s2 * s2)), 0.0) / mask.sum() * upstream
    return loss, gmu, gs2
