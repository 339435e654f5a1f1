"""Scratch: the systematic (truncation) bias of the tensor-core accumulation, as a multiplicative factor:
fit c in  got = (1 + c) * exact  over the outputs of spb_wino_gemm / spb_conv_gemm for several input distributions."""
import sys, torch
sys.path.insert(0, '.')
from scanpaths_b200 import _lib
from scanpaths_b200.models.baseline_attention import split_pair
lib = _lib.load()
dev = torch.device('cuda')

def split_dev(x):
    hi = torch.empty_like(x, dtype=torch.float16); lo = torch.empty_like(hi)
    _lib.check(lib.spb_split_fp16(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), 1, 1, 0, 1.0, 0, _lib.current_stream()))
    return hi, lo

def fit(got, ref):
    got, ref = got.double().flatten(), ref.flatten()
    c = float(((got - ref) * ref).sum() / (ref * ref).sum())
    res = got - ref * (1 + c)
    return c, float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), float(res.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())

def wino_case(name, u, w):
    rows, cols = u.shape[1], w.shape[0] // 24
    u_hi, u_lo = split_dev(u)
    w_hi, w_lo, inv = split_pair(w)
    # exact inputs as represented
    ue = u_hi.double() + u_lo.double() / 2048
    we = (w_hi.double() + w_lo.double() / 2048) * inv
    out = torch.empty((12, cols // 128, rows, 128), device=dev)
    _lib.check(lib.spb_wino_gemm(_lib.ptr(u_hi), _lib.ptr(u_lo), _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(out), rows, cols, inv, 0, _lib.current_stream()))
    m = torch.einsum("prk,pck->prc", ue, we.view(24, cols, 512)).view(6, 4, rows, cols)
    ref = torch.stack([m[:, 0] + m[:, 1] + m[:, 2], m[:, 1] - m[:, 2] - m[:, 3]], 1).reshape(12, rows, cols)
    got = out.permute(0, 2, 1, 3).reshape(12, rows, cols)
    c, e0, e1 = fit(got, ref)
    print('wino %-28s c = %+.3e   rel rms err %.3e -> %.3e after removing the factor' % (name, c, e0, e1))

def conv_case(name, a, w, ks):
    n, cols = a.shape[0], w.shape[0]
    a_hi, a_lo = split_dev(a)
    w_hi, w_lo, inv = split_pair(w.reshape(cols, -1))
    ae = a_hi.double() + a_lo.double() / 2048
    we = ((w_hi.double() + w_lo.double() / 2048) * inv).view(cols, ks, ks, 512)
    out = torch.empty((n * 1200, cols), device=dev)
    _lib.check(lib.spb_conv_gemm(_lib.ptr(a_hi), _lib.ptr(a_lo), _lib.ptr(w_hi), _lib.ptr(w_lo), None, cols, None, _lib.ptr(out), cols, n, cols, ks, inv, 1, _lib.current_stream()))
    ref = torch.nn.functional.conv2d(ae.permute(0, 3, 1, 2), we.permute(0, 3, 1, 2), padding=ks // 2).permute(0, 2, 3, 1).reshape(-1, cols)
    c, e0, e1 = fit(out, ref)
    print('conv ks=%d %-22s c = %+.3e   rel rms err %.3e -> %.3e after removing the factor' % (ks, name, c, e0, e1))

g = torch.Generator(device=dev).manual_seed(0)
R, C = 512, 256
un = torch.randn(24, R, 512, generator=g, device=dev)
wn = torch.randn(24 * C, 512, generator=g, device=dev) * 0.05
wino_case('u~N, w~N', un, wn)
wino_case('u>=0, w~N', un.clamp_min(0), wn)
wino_case('u>=0, w>=0', un.abs(), wn.abs())
wino_case('u~N*decay(k), w~N', un * torch.linspace(2, 0.05, 512, device=dev), wn)
wino_case('u~U(-1,1)*4, w~N*0.01', (torch.rand(24, R, 512, generator=g, device=dev) * 2 - 1) * 4, wn * 0.2)
an = torch.randn(2, 30, 40, 512, generator=g, device=dev) * 0.7
w3 = torch.randn(256, 3, 3, 512, generator=g, device=dev) * 0.02
conv_case('a~N, w~N', an, w3, 3)
conv_case('a>=0, w~N', an.clamp_min(0), w3, 3)
conv_case('a>=0, w>=0', an.abs(), w3.abs(), 3)
w1 = torch.randn(256, 1, 1, 512, generator=g, device=dev) * 0.05
conv_case('a~N, w~N', an, w1, 1)
w5 = torch.randn(128, 5, 5, 512, generator=g, device=dev) * 0.02
conv_case('a~N, w~N', an, w5, 5)
