"""CPU oracle: post-hoc scanpath sampling + log-likelihoods (TEST INFRASTRUCTURE ONLY).

numpy restatement of
  * Sampling.random_sample     /root/reference/OSIE/models/sampling.py:16-46
  * Sampling.generate_scanpath /root/reference/OSIE/models/sampling.py:48-77
  * LogAction / LogDuration / CrossEntropyLoss / MLPLogNormalDistribution
                               /root/reference/OSIE/models/loss.py:10-45
with the random draws INJECTED: the reference draws, in this order,
``q = empty(N*T, A).exponential_(1)`` inside ``Categorical.sample`` (torch
multinomial == exponential race ``argmax((p/sum p)/q)``) and ``z = randn(N, T)``.
Pinned by tests/golden/sampling.npz (reference outputs under a seeded generator
together with the q, z it consumed).
"""
from __future__ import annotations

import math

import numpy as np

EPS = np.float32(1e-7)


def random_sample(probs, mu, sigma2, q, z, min_length):
    """All float32.  probs [N,T,A]; mu, sigma2, z [N,T]; q [N,T,A]."""
    probs = np.asarray(probs, dtype=np.float32)
    p = probs.copy()
    p[:, :min_length, 0] = 0                                  # sampling.py:20
    # Categorical(probs=p) normalises: p / p.sum(-1, keepdim)  (float32)
    pn = p / p.sum(-1, keepdims=True, dtype=np.float32)
    ratio = pn / np.asarray(q, dtype=np.float32)
    actions = ratio.argmax(-1)                                 # first index on ties
    sel = np.take_along_axis(probs, actions[..., None], -1)[..., 0]   # from the UNMASKED probs (:23-24)
    dur = np.exp(np.asarray(z, np.float32) * np.asarray(sigma2, np.float32) + np.asarray(mu, np.float32),
                 dtype=np.float32)                             # variance used as scale (:27)
    N, T = actions.shape
    length = np.zeros(N, dtype=np.float32)
    for t in range(T):                                         # :29-33 (a stop at t=0 is not recorded)
        hit = np.logical_and(length == 0, actions[:, t] == 0)
        length[hit] = t
    length[length == 0] = T
    return dict(selected_actions=actions.astype(np.int64), selected_actions_probs=sel,
                durations=dur, scanpath_length=length[:, None])


def generate_scanpath(actions, durations, map_width=40, map_height=30, width=320, height=240):
    """Returns (list of [L,3] f64 arrays (x, y, seconds), action_masks, duration_masks)."""
    xg, yg = float(width / map_width), float(height / map_height)
    N, T = actions.shape
    am = np.zeros((N, T), dtype=np.float32)
    dm = np.zeros((N, T), dtype=np.float32)
    fix = []
    for n in range(N):
        rows = []
        for t in range(T):
            a = int(actions[n, t])
            am[n, t] = 1
            if a == 0:
                break
            cell = a - 1
            rows.append(((cell % map_width) * xg + xg / 2, (cell // map_width) * yg + yg / 2,
                         float(durations[n, t])))
            dm[n, t] = 1
        fix.append(np.array(rows, dtype=np.float64).reshape(-1, 3))
    return fix, am, dm


def log_action(sel_prob, mask):
    """loss.py:34-37 -- each row's masked sum over the WHOLE batch's mask.sum()."""
    sel_prob = np.asarray(sel_prob, np.float32)
    mask = np.asarray(mask, np.float32)
    return (np.log(sel_prob + EPS) * mask).sum(-1, dtype=np.float32) / mask.sum(dtype=np.float32)


def log_duration(dur, mu, sigma2, mask):
    """loss.py:39-45."""
    dur, mu, sigma2, mask = (np.asarray(v, np.float32) for v in (dur, mu, sigma2, mask))
    two_pi = np.float32(2 * math.pi)
    item = np.log(np.float32(1) / (dur + EPS) * np.float32(1) / np.sqrt(two_pi * sigma2)) \
        + (-(np.log(dur + EPS) - mu) ** 2 / (np.float32(2) * sigma2))
    return (item * mask).sum(-1, dtype=np.float32) / mask.sum(dtype=np.float32)


def cross_entropy_loss(logits, gt_idx, mask):
    """loss.py:10-14 with a one-hot target given as indices."""
    x = np.asarray(logits, np.float64)
    x = x - x.max(-1, keepdims=True)
    sm = np.exp(x) / np.exp(x).sum(-1, keepdims=True)
    picked = np.take_along_axis(sm, np.asarray(gt_idx)[..., None], -1)[..., 0]
    mask = np.asarray(mask, np.float64)
    return -(np.log(picked + 1e-7) * mask).sum() / mask.sum()


def lognormal_nll(mu, sigma2, gt, mask):
    """loss.py:27-32."""
    mu, sigma2, gt, mask = (np.asarray(v, np.float64) for v in (mu, sigma2, gt, mask))
    logpdf = np.log(1 / (gt + 1e-7) * 1 / np.sqrt(2 * math.pi * sigma2)) \
        + (-(np.log(gt + 1e-7) - mu) ** 2 / (2 * sigma2))
    return -(logpdf[mask == 1]).sum() / mask.sum()
