"""Scratch: error structure + timing of the tcgen05 conv."""
import sys, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, '.')
from scanpaths_b200 import _lib
from scanpaths_b200.models.baseline_attention import split_pair
lib = _lib.load()
dev = torch.device('cuda')
def run(ks, n_images, cols, use_tc, relu_a=False, reps=1):
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(n_images, 30, 40, 512, generator=g, device=dev) * 0.7
    if relu_a: a = a.clamp_min(0)
    w = torch.randn(cols, ks, ks, 512, generator=g, device=dev) * 0.02
    a_hi = torch.empty_like(a, dtype=torch.float16); a_lo = torch.empty_like(a_hi)
    _lib.check(lib.spb_split_fp16(_lib.ptr(a), _lib.ptr(a_hi), _lib.ptr(a_lo), a.numel(), 1, 1, 0, 1.0, 0, _lib.current_stream()))
    w_hi, w_lo, inv = split_pair(w.reshape(cols, -1))
    out = torch.empty((n_images * 1200, cols), device=dev)
    def call():
        _lib.check(lib.spb_conv_gemm(_lib.ptr(a_hi), _lib.ptr(a_lo), _lib.ptr(w_hi), _lib.ptr(w_lo), None, cols, None, _lib.ptr(out), cols, n_images, cols, ks, inv, int(use_tc), _lib.current_stream()))
    call(); torch.cuda.synchronize()
    if reps > 1:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * n_images * 1200 * cols * ks * ks * 512
        print('  time %.3f ms  -> %.1f TFLOP/s algorithmic (x3 issued = %.1f)' % (ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
    nref = min(n_images, 2)
    ref = F.conv2d(a[:nref].permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), padding=ks // 2).permute(0, 2, 3, 1).reshape(-1, cols)
    o = out[:nref * 1200].double()
    d = o - ref
    big = ref.abs() > ref.abs().median()
    signed = (d * ref.sign() / ref.abs())[big]
    print('ks=%d tc=%d relu=%d: max abs err %.3e (max|ref| %.2f rms %.2f); signed rel err on big outputs: mean %.3e rms %.3e' % (
        ks, use_tc, relu_a, d.abs().max().item(), ref.abs().max().item(), ref.pow(2).mean().sqrt().item(), signed.mean().item(), signed.pow(2).mean().sqrt().item()))
for relu in (False, True):
    for ks in (3, 5):
        for tc in (0, 1):
            run(ks, 2, 512, tc, relu)
print('timing, 64 images:')
run(3, 64, 2048, 1, reps=5)
run(5, 64, 512, 1, reps=5)
