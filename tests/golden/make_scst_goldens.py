"""SCST / loss goldens: the reference's own loss functions (models/loss.py) and the reward / baseline /
loss arithmetic of OSIE/train.py:242-258, executed here with torch autograd on the recorded samples of
sampling.npz, plus the gradients of the two supervised losses.  Inputs come from sampling.npz (probs,
mu, sigma2, six recorded trials: min_length 1 and 2 x three seeds) and a seeded synthetic reward table.
Run in the authoring container (needs /root/reference); writes scst.npz next to this file."""
import os
import sys

import numpy as np
import scipy.stats
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refload  # noqa: E402

TRIALS = ["m1_t0_", "m1_t1_", "m1_t2_", "m2_t0_", "m2_t1_", "m2_t2_"]
REJECTED = 2            # trial index whose table carries a NaN row (train.py:237-238 rejects it)
K_USE = 4               # rl_sample_number: the first 4 accepted trials count


def reward_table(n_trials, n_images, seed=5):
    rng = np.random.default_rng(seed)
    tab = np.zeros((n_trials, n_images, 11), dtype=np.float32)
    tab[:, :, 5:7] = rng.uniform(0.05, 0.9, (n_trials, n_images, 2)).astype(np.float32)
    tab[:, :, 7] = rng.integers(3, 15, (n_trials, n_images))
    tab[:, :, 8:] = rng.uniform(0.5, 0.95, (n_trials, n_images, 3)).astype(np.float32)
    tab[REJECTED, 1, :] = np.nan
    return tab


def gen_scst():
    ns = refload.load_reference("OSIE")
    g = np.load(os.path.join(HERE, "sampling.npz"))
    probs = torch.tensor(g["probs"], requires_grad=True)
    mu = torch.tensor(g["mu"], requires_grad=True)
    sigma2 = torch.tensor(g["sigma2"], requires_grad=True)
    N = probs.shape[0]
    table = reward_table(len(TRIALS), N)
    out = {"table": table}
    # ---- the loop body of train.py:223-250 on the recorded samples
    rewards, nla_b, nld_b, used = [], [], [], []
    for k, tag in enumerate(TRIALS):
        if len(used) >= K_USE:
            break
        metrics_reward = table[k]
        if np.any(np.isnan(metrics_reward)):
            continue
        used.append(k)
        actions = torch.tensor(g[tag + "actions"])
        prob_sample_actions = torch.gather(probs, dim=2, index=actions.unsqueeze(-1)).squeeze(-1)   # sampling.py:23-24
        t = torch.tensor(g[tag + "dur"]).data.clone()
        am, dm = torch.tensor(g[tag + "action_mask"]), torch.tensor(g[tag + "duration_mask"])
        nla_b.append((-ns.loss.LogAction(prob_sample_actions, am)).unsqueeze(0))
        nld_b.append((-ns.loss.LogDuration(t, mu, sigma2, dm)).unsqueeze(0))
        rewards.append(torch.tensor(metrics_reward, dtype=torch.float32).unsqueeze(0))
    # ---- train.py:248-258
    neg_log_actions_tensor = torch.cat(nla_b, dim=0)
    neg_log_durations_tensor = torch.cat(nld_b, dim=0)
    metrics_reward_tensor = torch.cat(rewards, dim=0)
    metrics_reward_hmean = scipy.stats.hmean(metrics_reward_tensor[:, :, 5:7].cpu(), axis=-1)
    metrics_reward_hmean_tensor = torch.tensor(metrics_reward_hmean)
    baseline_reward_hmean_tensor = metrics_reward_hmean_tensor.mean(0, keepdim=True)
    loss_actions = (neg_log_actions_tensor * (metrics_reward_hmean_tensor - baseline_reward_hmean_tensor)).sum()
    loss_duration = (neg_log_durations_tensor * (metrics_reward_hmean_tensor - baseline_reward_hmean_tensor)).sum()
    loss = loss_actions + loss_duration
    loss.backward()
    out.update(used=np.array(used), loss=loss.detach().numpy(), loss_actions=loss_actions.detach().numpy(),
               loss_duration=loss_duration.detach().numpy(), neg_log_actions=neg_log_actions_tensor.detach().numpy(),
               neg_log_durations=neg_log_durations_tensor.detach().numpy(),
               reward_hmean=np.asarray(metrics_reward_hmean),
               advantage=(metrics_reward_hmean_tensor - baseline_reward_hmean_tensor).numpy(),
               grad_probs=probs.grad.numpy(), grad_mu=mu.grad.numpy(), grad_sigma2=sigma2.grad.numpy())
    # ---- gradients of LogAction / LogDuration on their own (one trial, random upstream weights)
    tag = TRIALS[0]
    w = torch.tensor(np.random.default_rng(9).normal(size=N).astype(np.float32))
    p = torch.tensor(g[tag + "sel_prob"], requires_grad=True)
    mu2 = torch.tensor(g["mu"], requires_grad=True); s22 = torch.tensor(g["sigma2"], requires_grad=True)
    x = torch.tensor(g[tag + "dur"], requires_grad=True)
    am, dm = torch.tensor(g[tag + "action_mask"]), torch.tensor(g[tag + "duration_mask"])
    ((ns.loss.LogAction(p, am) * w).sum() + (ns.loss.LogDuration(x, mu2, s22, dm) * w).sum()).backward()
    out.update(row_w=w.numpy(), la_grad_p=p.grad.numpy(), ld_grad_mu=mu2.grad.numpy(), ld_grad_sigma2=s22.grad.numpy(),
               ld_grad_x=x.grad.numpy())
    # ---- supervised losses and their gradients (loss.py:10-32; train.py:170-173)
    logits = torch.tensor(g["loss_logits"], requires_grad=True)
    idx = torch.tensor(g["loss_gt_idx"])
    gt = torch.zeros_like(logits).scatter_(2, idx.unsqueeze(-1), 1.0)
    # a soft target row too (the function takes a dense gt)
    gt[0, 0] = torch.softmax(torch.tensor(np.random.default_rng(3).normal(size=logits.shape[-1]).astype(np.float32)), 0)
    mask = torch.tensor(g["loss_mask"])
    ce = ns.loss.CrossEntropyLoss(logits, gt, mask)
    (ce * 1.7).backward()
    mu3 = torch.tensor(g["mu"], requires_grad=True); s23 = torch.tensor(g["sigma2"], requires_grad=True)
    gt_dur = torch.tensor(g["loss_gt_dur"])
    nll = ns.loss.MLPLogNormalDistribution(mu3, s23, gt_dur, mask)
    (nll * 0.6).backward()
    out.update(ce_gt00=gt[0, 0].numpy(), ce=ce.detach().numpy(), ce_grad_logits=logits.grad.numpy(),
               nll=nll.detach().numpy(), nll_grad_mu=mu3.grad.numpy(), nll_grad_sigma2=s23.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "scst.npz"), **out)
    print("scst goldens written: loss %.6f, used trials %s" % (float(loss), used))


if __name__ == "__main__":
    gen_scst()
