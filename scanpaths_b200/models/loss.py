"""Drop-in for the four loss / log-likelihood functions the reference's training loops call
(models/loss.py:10-45), on the GPU: forward and analytic backward are CUDA kernels (csrc/loss.cu)
behind ``torch.autograd.Function``s, so they slot into the reference's autograd graph
(``train.py:170-173`` supervised, ``:242-243`` SCST) with the same names and signatures.

    CrossEntropyLoss(input, gt, mask)                      loss.py:10-14
    MLPLogNormalDistribution(mu, sigma2, gt, mask)         loss.py:27-32
    LogAction(input, mask)                                 loss.py:34-37
    LogDuration(input, mu, sigma2, mask)                   loss.py:39-45

The whole SCST tail (both log-likelihoods of K trials, the self-critical baseline and the loss) as
ONE fused op lives in ``scanpaths_b200.scst``.  The unused losses of that file (Rayleigh, SmoothL1,
NSS, CC, KLD) are dead code in the reference and out of scope.  There is no CPU path.
"""
from __future__ import annotations

import torch

from .. import _lib

epsilon = 1e-7


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _need_cuda(t):
    if not t.is_cuda:
        raise _lib.SpbError("scanpaths_b200 has no CPU path: tensors must be on the GPU")


class _LogLikRows(torch.autograd.Function):
    """(LogAction rows, LogDuration rows) of K stacked calls; either half may be absent (None inputs)."""

    @staticmethod
    def forward(ctx, p, mask_a, x, mu, s2, mask_d):
        lib = _lib.load()
        ref = p if p is not None else x
        _need_cuda(ref)
        K, N, T = ref.shape
        dev = ref.device
        ctx.has_a, ctx.has_d = p is not None, x is not None
        pa = _f32(p) if ctx.has_a else None
        ma = _f32(mask_a) if ctx.has_a else None
        xd, mud, s2d, md = (_f32(x), _f32(mu), _f32(s2), _f32(mask_d)) if ctx.has_d else (None,) * 4
        out_a = torch.empty((K, N), dtype=torch.float32, device=dev) if ctx.has_a else None
        out_d = torch.empty((K, N), dtype=torch.float32, device=dev) if ctx.has_d else None
        msum = torch.empty((K, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spb_loglik_rows(_lib.ptr(pa), None, None, _lib.ptr(xd), _lib.ptr(mud), _lib.ptr(s2d),
                                           _lib.ptr(ma), _lib.ptr(md), K, N, T, 0, _lib.ptr(out_a), _lib.ptr(out_d),
                                           _lib.ptr(msum), _lib.current_stream()), "spb_loglik_rows")
        ctx.save_for_backward(*[t for t in (pa, ma, xd, mud, s2d, md, msum) if t is not None])
        ctx.shape = (K, N, T)
        return out_a, out_d

    @staticmethod
    def backward(ctx, ga, gd):
        lib = _lib.load()
        K, N, T = ctx.shape
        saved = list(ctx.saved_tensors)
        pa, ma = (saved.pop(0), saved.pop(0)) if ctx.has_a else (None, None)
        xd, mud, s2d, md = (saved.pop(0), saved.pop(0), saved.pop(0), saved.pop(0)) if ctx.has_d else (None,) * 4
        msum = saved.pop(0)
        dev = msum.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        use_a = ctx.has_a and ga is not None
        use_d = ctx.has_d and gd is not None
        gp = f(K, N, T) if use_a else None
        gmu, gs2, gx = (f(N, T), f(N, T), f(K, N, T)) if use_d else (None, None, None)
        with torch.cuda.device(dev):
            _lib.check(lib.spb_loglik_rows_backward(
                _lib.ptr(pa), _lib.ptr(xd), _lib.ptr(mud), _lib.ptr(s2d), _lib.ptr(ma), _lib.ptr(md), _lib.ptr(msum),
                _lib.ptr(_f32(ga)) if use_a else None, _lib.ptr(_f32(gd)) if use_d else None, K, N, T, _lib.ptr(gp),
                _lib.ptr(gmu), _lib.ptr(gs2), _lib.ptr(gx), _lib.current_stream()), "spb_loglik_rows_backward")
        return gp, None, gx, gmu, gs2, None


def LogAction(input, mask):
    """[N,T] selected-action probabilities and mask -> [N]: each row's masked log sum over the WHOLE batch's
    mask.sum() (loss.py:36).  Differentiable w.r.t. `input` (SCST back-propagates through it, train.py:242)."""
    out, _ = _LogLikRows.apply(input.unsqueeze(0), mask.unsqueeze(0), None, None, None, None)
    return out[0]


def LogDuration(input, log_normal_mu, log_normal_sigma2, mask):
    """[N,T] durations under the log-normal (mu, sigma2 = variance) -> [N] (loss.py:39-45).  Differentiable
    w.r.t. mu, sigma2 (and input)."""
    _, out = _LogLikRows.apply(None, None, input.unsqueeze(0), log_normal_mu, log_normal_sigma2, mask.unsqueeze(0))
    return out[0]


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, gt, mask):
        lib = _lib.load()
        _need_cuda(logits)
        z, g, m = _f32(logits), _f32(gt), _f32(mask)
        A = z.shape[-1]
        rows = z.numel() // A
        dev = z.device
        row_loss = torch.empty((rows,), dtype=torch.float32, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        msum = torch.empty((1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spb_cross_entropy(_lib.ptr(z), _lib.ptr(g), _lib.ptr(m), rows, A, _lib.ptr(row_loss),
                                             _lib.ptr(loss), _lib.ptr(msum), None, None, _lib.current_stream()),
                       "spb_cross_entropy")
        ctx.save_for_backward(z, g, m, msum)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        z, g, m, msum = ctx.saved_tensors
        A = z.shape[-1]
        rows = z.numel() // A
        gz = torch.empty_like(z)
        go = _f32(grad).reshape(1)
        with torch.cuda.device(z.device):
            _lib.check(lib.spb_cross_entropy(_lib.ptr(z), _lib.ptr(g), _lib.ptr(m), rows, A, None, None, _lib.ptr(msum),
                                             _lib.ptr(go), _lib.ptr(gz), _lib.current_stream()), "spb_cross_entropy")
        return gz, None, None


def CrossEntropyLoss(input, gt, mask):
    """logits [N,T,A], dense target [N,T,A], mask [N,T] -> scalar (loss.py:10-14)."""
    return _CrossEntropy.apply(input, gt, mask)


class _LogNormalNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, s2, gt, mask):
        lib = _lib.load()
        _need_cuda(mu)
        a, b, g, m = _f32(mu), _f32(s2), _f32(gt), _f32(mask)
        n, dev = a.numel(), a.device
        item = torch.empty((n,), dtype=torch.float32, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        msum = torch.empty((1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spb_lognormal_nll(_lib.ptr(a), _lib.ptr(b), _lib.ptr(g), _lib.ptr(m), n, _lib.ptr(item),
                                             _lib.ptr(loss), _lib.ptr(msum), None, None, None, _lib.current_stream()),
                       "spb_lognormal_nll")
        ctx.save_for_backward(a, b, g, m, msum)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        a, b, g, m, msum = ctx.saved_tensors
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        go = _f32(grad).reshape(1)
        with torch.cuda.device(a.device):
            _lib.check(lib.spb_lognormal_nll(_lib.ptr(a), _lib.ptr(b), _lib.ptr(g), _lib.ptr(m), a.numel(), None, None,
                                             _lib.ptr(msum), _lib.ptr(go), _lib.ptr(ga), _lib.ptr(gb),
                                             _lib.current_stream()), "spb_lognormal_nll")
        return ga, gb, None, None


def MLPLogNormalDistribution(log_normal_mu, log_normal_sigma2, gt, mask):
    """[N,T] -> scalar: -sum over mask == 1 of the log-normal log density / mask.sum() (loss.py:27-32)."""
    return _LogNormalNLL.apply(log_normal_mu, log_normal_sigma2, gt, mask)
