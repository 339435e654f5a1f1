"""Timing of spb_sal_conv (the encoder's last layer on tcgen05): 256 images."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scanpaths_b200.models.baseline_attention import baseline
dev = torch.device("cuda")
m = baseline(task="OSIE", wave=256)
m.sal_conv = torch.nn.Conv2d(2048, 512, kernel_size=3, padding=1)
m = m.cuda()
x = torch.randn((256, 2048, 30, 40), device=dev).clamp_min_(0)
for _ in range(2):
    m.sal_conv_cuda(x)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0.record()
for _ in range(5):
    m.sal_conv_cuda(x)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 5
print("sal_conv 256 images: %.2f ms (%.0f TFLOP/s algorithmic, %.0f issued) incl. the NCHW->NHWC operand split" % (
    ms, 256 * 22.65e9 / ms / 1e9, 3 * 256 * 22.65e9 / ms / 1e9))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for _ in range(2):
    torch.relu(m.sal_conv(x))
torch.cuda.synchronize(); t0.record()
for _ in range(5):
    torch.relu(m.sal_conv(x))
t1.record(); torch.cuda.synchronize()
print("torch/cuDNN fp32 (TF32 off): %.2f ms" % (t0.elapsed_time(t1) / 5))
