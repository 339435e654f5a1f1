"""HBM bandwidth by read/write mix (torch kernels): pure write (fill), copy (1R:1W), read only (sum), 1R:3W."""
import torch
dev = torch.device("cuda")
n = 1 << 29                                   # 2 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device=dev); y = torch.empty_like(x)
def t(fn, bytes_, name, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-28s %.3f ms  %.2f TB/s" % (name, ms, bytes_ / ms / 1e9))
t(lambda: x.fill_(1.0), 4 * n, "write only (fill 2 GiB)")
t(lambda: y.copy_(x), 8 * n, "copy 1R:1W")
t(lambda: x.sum(), 4 * n, "read only (sum)")
q = n // 4
src = x[:q]
dst = y[:3 * q].view(3, q)
t(lambda: dst.copy_(src.unsqueeze(0).expand(3, q)), 4 * 4 * q, "1R:3W (broadcast copy)")
