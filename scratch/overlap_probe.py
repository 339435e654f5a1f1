"""Scratch: do two waves on two streams overlap (GEMM of one with the memory-bound kernels of the other)
when the persistent GEMM leaves some SMs free (SPB_GEMM_CTAS)?"""
import os, sys, time, torch
sys.path.insert(0, '.')
from scanpaths_b200.models.baseline_attention import CudaDecoder
from scanpaths_b200.weights import random_state_dict, synthetic_features
dev = torch.device('cuda')
sd = random_state_dict("OSIE", 0, calibrated=True)
W = 256
d0 = CudaDecoder(sd, "OSIE", 16, dev, wave=W); d1 = CudaDecoder(sd, "OSIE", 16, dev, wave=W)
vf0 = synthetic_features(W, 1).to(dev); vf1 = synthetic_features(W, 2).to(dev)
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
def run(two, reps):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if two:
        s0.wait_stream(torch.cuda.current_stream()); s1.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            with torch.cuda.stream(s0): d0.decode(vf0)
            with torch.cuda.stream(s1): d1.decode(vf1)
        torch.cuda.current_stream().wait_stream(s0); torch.cuda.current_stream().wait_stream(s1)
    else:
        for _ in range(reps):
            d0.decode(vf0); d1.decode(vf1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * reps)
run(False, 1); run(True, 1)
print('SPB_GEMM_CTAS=%s: one stream %.1f ms per wave, two streams %.1f ms per wave' % (
    os.environ.get('SPB_GEMM_CTAS', '148'), run(False, 3), run(True, 3)))
