"""Scratch: time prep / sample / score at the OSIE config on one GPU."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from scanpaths_b200 import scoring as S
from scanpaths_b200.models.sampling import Sampling
from golden.make_goldens import human_paths
dev = torch.device('cuda')
N, T, A, K, Sn = 4096, 16, 1201, 64, 15
gen = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(N, T, A, generator=gen, device=dev); logits[:, :, 0] += 4.5
probs = torch.softmax(logits, -1)
mu = torch.full((N, T), -1.4, device=dev); s2 = torch.full((N, T), 0.15, device=dev)
rng = np.random.default_rng(0)
H = human_paths(rng, N * Sn)
cfg = S.ScoreConfig.evaluation()
hp = S.pack_paths(H, cfg)
ph, ps = S.grid_pairs(N, K, Sn, dev)
sampler = Sampling(convLSTM_length=T, min_length=1, seed=1)
ws = S.Workspace(int(hp.nwd.max().item()), dev)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    e0 = ev(); out = sampler.sample_paths(probs, mu, s2, K=K)
    e1 = ev(); pp = S.prep_paths(out['xyd'], out['len'], cfg)
    e2 = ev(); sc = S.score_pairs(hp, pp, ph, ps, cfg, workspace=ws, check=False)
    e3 = ev(); tab, rew = S.reduce_pairs_eval(sc, Sn)
    e4 = ev(); torch.cuda.synchronize()
    print('iter', it, 'sample %.2f ms prep %.2f ms score %.2f ms reduce %.2f ms' % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)))
lens = out['len'].float(); print('pred len mean', lens.mean().item(), 'nwd pred mean', pp.nwd.float().mean().item(), 'nwd human mean', hp.nwd.float().mean().item())
lh, lp = hp.len[ph.long()].double(), pp.len[ps.long()].double()
m = torch.minimum(lh, lp)
stde = m * (lp + 1) * (lh + 1) - (lp + lh + 2) * m * (m + 1) / 2 + m * (m + 1) * (2 * m + 1) / 6
cells = (hp.nwd[ph.long()].double() * pp.nwd[ps.long()].double() + 2 * lh * lp + stde).sum().item()
print('pairs', ph.numel(), 'cell updates (wd + wod + sed + stde windows) %.4e' % cells, 'mean scores', sc.nanmean(0).tolist())
