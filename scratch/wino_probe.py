"""Scratch: timing of the Winograd gate GEMM (spb_wino_gemm) at the product size (256 images: 38400 tiles),
32-k-step accumulators (h-gates) vs 8-k-step accumulators (fine drain)."""
import sys, torch
sys.path.insert(0, '.')
from scanpaths_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda')
rows, cols = 38400, 2048
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator(device=dev).manual_seed(0)
u_hi = (torch.randn(24, rows, 512, generator=g, device=dev) * 100).to(torch.float16)
u_lo = (torch.randn(24, rows, 512, generator=g, device=dev) * 0.05).to(torch.float16)
w_hi = (torch.randn(24 * cols, 512, generator=g, device=dev) * 100).to(torch.float16)
w_lo = (torch.randn(24 * cols, 512, generator=g, device=dev) * 0.05).to(torch.float16)
out = torch.empty((12, cols // 128, rows, 128), device=dev)
for fine in (0, 1):
    def call():
        _lib.check(lib.spb_wino_gemm(_lib.ptr(u_hi), _lib.ptr(u_lo), _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(out), rows, cols, 1.0, fine, _lib.current_stream()))
    call(); call(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * 24 * rows * cols * 512
    print('wino gemm fine=%d: %.3f ms -> %.1f TFLOP/s algorithmic, %.1f issued' % (fine, ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
