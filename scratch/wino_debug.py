import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from scanpaths_b200.models.baseline_attention import CudaDecoder
from scanpaths_b200.weights import random_state_dict, synthetic_features
from golden.make_decoder_goldens import CASES, COCO_TASKS
dev = torch.device('cuda')
def run(name, mode, poison):
    task, n, T, wseed, fseed, bstd = CASES[name]
    g = np.load('tests/golden/decoder_%s.npz' % name)
    sd = random_state_dict(task, wseed, calibrated=True, bias_std=bstd)
    if poison is not None:
        x = torch.full((3 * 2**28,), poison, device=dev); torch.cuda.synchronize(); del x
    dec = CudaDecoder(sd, task, T, dev, wave=n, use_tensor_cores=mode)
    if task == 'OSIE': vf, att, tasks = synthetic_features(n, fseed), None, None
    else:
        vf, att = synthetic_features(n, fseed, attention=True); tasks = COCO_TASKS[:n] if task == 'COCO_Search18' else None
    probs, mu, s2, amap = dec.decode(vf.to(dev), None if att is None else att.to(dev), tasks)
    pre = 'good_' if task == 'AiR' else ''
    ref = g['f64_' + pre + 'all_actions_prob'][:, :T]
    rel = np.abs(probs[0].cpu().numpy().astype(np.float64) - ref) / ref
    per_step = [float(np.nanmax(rel[:, t])) if not np.isnan(rel[:, t]).all() else float('nan') for t in range(T)]
    print(name, 'mode', mode, 'poison', poison, 'max rel per step', ['%.1e' % v for v in per_step], 'nan count', int(np.isnan(rel).sum()))
for name in ('coco', 'osie'):
    run(name, 1, None)
    run(name, 1, float('nan'))
    run(name, 1, 1e30)
    run(name, 2, float('nan'))
