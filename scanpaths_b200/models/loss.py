"""The four loss / log-likelihood functions the reference's training loops call
(models/loss.py:10-45), kept as differentiable torch expressions: they are
O(N*T) element-wise tails of the SCST step whose gradients flow back into the
PyTorch forward, so they stay in autograd.  The unused
losses of that file (Rayleigh, SmoothL1, NSS, CC, KLD) are out of scope.
"""
import math

import torch
import torch.nn.functional as F

epsilon = 1e-7


def _lognormal_logpdf(x, mu, sigma2):
    # log-normal density with sigma2 as the variance
    return torch.log(1 / (x + epsilon) * 1 / (torch.sqrt(2 * math.pi * sigma2))) \
        + (-(torch.log(x + epsilon) - mu) ** 2 / (2 * sigma2))


def CrossEntropyLoss(input, gt, mask):
    prob = F.softmax(input, dim=-1)
    return -(gt * torch.log(prob + epsilon) * mask.unsqueeze(-1)).sum() / mask.sum()


def MLPLogNormalDistribution(log_normal_mu, log_normal_sigma2, gt, mask):
    logpdf = _lognormal_logpdf(gt, log_normal_mu, log_normal_sigma2)
    return -(logpdf[mask == 1]).sum() / mask.sum()


def LogAction(input, mask):
    # each row's masked sum over the WHOLE batch's mask count (loss.py:36)
    return (torch.log(input + epsilon) * mask).sum(dim=-1) / mask.sum()


def LogDuration(input, log_normal_mu, log_normal_sigma2, mask):
    return (_lognormal_logpdf(input, log_normal_mu, log_normal_sigma2) * mask).sum(dim=-1) / mask.sum()
