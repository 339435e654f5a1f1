"""ncu / timing probe of the scoring kernels: one launch of the OSIE wave shape (256 images x 64 samples x 15
subjects = 245,760 pairs) and one of the human_evaluation shape.  Usage: python scratch/score_probe.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_humans
from scanpaths_b200 import scoring as S
from scanpaths_b200.models.sampling import Sampling

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda")
cfg = S.ScoreConfig.evaluation(device=dev, dur_scale=1000.0)
N, K, Sn, T, A = 256, 64, 15, 16, 1201
hx, hl = synth_humans(N, Sn, 5)
hp = S.prep_paths(torch.from_numpy(hx.reshape(N * Sn, -1, 3)).to(dev), torch.from_numpy(hl.reshape(-1)).to(dev), cfg)
g = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(N, T, A, generator=g, device=dev); logits[:, :, 0] += 4.6
probs = torch.softmax(logits, -1)
mu = torch.full((N, T), float(np.log(0.25)), device=dev); s2 = torch.full((N, T), 0.15, device=dev)
smp = Sampling(convLSTM_length=T, min_length=1, seed=1).sample_paths(probs, mu, s2, K)
pp = S.prep_paths(smp["xyd"], smp["len"], cfg)
ph, ps = S.grid_pairs(N, K, Sn, dev)
ws = S.Workspace(int(hp.nwd.max().item()), dev)
out = torch.empty((ph.numel(), 4), dtype=torch.float64, device=dev)
print("pairs", ph.numel(), "pred len mean %.1f nwd mean %.1f max %d; human nwd mean %.1f" % (
    pp.len.float().mean(), pp.nwd.float().mean(), int(pp.nwd.max()), hp.nwd.float().mean()))
for _ in range(2):
    S.score_pairs(hp, pp, ph, ps, cfg, workspace=ws, out=out, check=False)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0.record()
for _ in range(reps):
    S.score_pairs(hp, pp, ph, ps, cfg, workspace=ws, out=out, check=False)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
print("OSIE wave: %.3f ms per launch, %.1f M pairs/s" % (ms, ph.numel() / ms / 1e3))
