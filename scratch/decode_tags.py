"""Per-tag CUDA-event times of a 256-image decode (the library's profile brackets), plus the untagged total."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scanpaths_b200 import _lib
from scanpaths_b200.models.baseline_attention import CudaDecoder
from scanpaths_b200.weights import random_state_dict
TAGS = {1: "conv3x3_x", 2: "wino_gemm_h", 4: "cell", 5: "head", 6: "feedback", 7: "rank1", 8: "prep", 9: "wino_input"}
dev = torch.device("cuda")
lib = _lib.load()
task = sys.argv[1] if len(sys.argv) > 1 else "OSIE"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, T = 256, 16
sd = random_state_dict(task, 5, calibrated=True, bias_std=0.05)
g = torch.Generator(device=dev).manual_seed(21)
vf = torch.randn((N, 512, 30, 40), generator=g, device=dev).clamp_min_(0)
att = torch.rand((N, 1, 30, 40), generator=g, device=dev) if task != "OSIE" else None
dec = CudaDecoder(sd, task, T, dev, wave=N)
for _ in range(3): dec.decode(vf, att)
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): dec.decode(vf, att)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 4
    cap = 4000
    _lib.check(lib.spb_profile_enable(cap), "enable")
    dec.decode(vf, att)
    torch.cuda.synchronize()
    ms = np.zeros(cap, dtype=np.float32); tg = np.zeros(cap, dtype=np.int32); n = C.c_int32(0)
    _lib.check(lib.spb_profile_collect(_lib.ptr(ms), _lib.ptr(tg), cap, C.byref(n)), "collect")
    lib.spb_profile_enable(0)
    ms, tg = ms[:n.value], tg[:n.value]
    line = ", ".join("%s %.2f (%d x %.3f)" % (TAGS.get(k, k), ms[tg == k].sum(), (tg == k).sum(), ms[tg == k].mean()) for k in sorted(set(tg.tolist())))
    print("%s: %.2f ms per wave | tagged sum %.2f | %s" % (task, total, ms.sum(), line), flush=True)
