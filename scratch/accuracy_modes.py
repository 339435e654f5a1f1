"""Decode error vs the float64 oracle at T = 16 for feature scales 1, 2, 4 and every tensor-core route
(1 = Winograd h + Winograd x, 2 = direct h + direct x, 3 = Winograd h + direct x)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import decoder as OD
from scanpaths_b200.models.baseline_attention import CudaDecoder
from scanpaths_b200.weights import random_state_dict, synthetic_features
dev = torch.device("cuda"); T = 16
torch.set_num_threads(os.cpu_count())
sd = random_state_dict("OSIE", 21, calibrated=True, bias_std=0.05)
for scale in (1.0, 2.0, 4.0):
    vf = synthetic_features(1, 21) * scale
    with torch.no_grad():
        p64 = OD.decode(sd, vf.double(), "OSIE", steps=T)["all_actions_prob"].numpy()
        p32 = OD.decode(sd, vf.float(), "OSIE", steps=T)["all_actions_prob"].double().numpy()
    rel = lambda a: float((np.abs(a - p64) / p64).max())
    line = ["scale %g  f32-ref %.2e" % (scale, rel(p32))]
    for mode in (1, 5, 2, 3, 4):  # 1 = product path (Winograd F(2x4) h, F(2x2) x), 5 = F(2x4) h + direct x, 2 = direct, 3 = F(2x4) both, 4 = same, fine-drain x
        try:
            p = CudaDecoder(sd, "OSIE", T, dev, wave=1, use_tensor_cores=mode).decode(vf.to(dev))[0][0]
            line.append("mode%d %.2e" % (mode, rel(p.double().cpu().numpy())))
        except Exception as e:
            line.append("mode%d failed: %s" % (mode, e))
    print("  ".join(line), flush=True)
