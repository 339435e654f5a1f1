import sys, torch, numpy as np
sys.path.insert(0, '.')
from scanpaths_b200.models.baseline_attention import CudaDecoder
from scanpaths_b200.weights import random_state_dict, synthetic_features
from oracle import decoder as OD
torch.set_num_threads(16)
dev = torch.device('cuda')
for seed, scale in ((12, 4.0), (12, 2.0), (12, 1.0)):
    sd = random_state_dict('OSIE', seed, calibrated=True, bias_std=0.05)
    vf = synthetic_features(1, seed) * scale
    T = 16
    with torch.no_grad():
        ref = OD.decode(sd, vf.double(), 'OSIE', steps=T)
        ref32 = OD.decode(sd, vf.float(), 'OSIE', steps=T)
    p64 = ref['all_actions_prob'].numpy()
    rel32 = np.abs(ref32['all_actions_prob'].double().numpy() - p64) / p64
    print('scale', scale, 'torch-fp32 CPU vs fp64: max rel per step', ['%.1e' % rel32[:, t].max() for t in range(T)])
    amap = ref['action_map'].numpy()
    print('   max logit', amap.max(), 'stop prob', p64[0, :, 0].round(3))
    for mode in (0, 2, 1):
        dec = CudaDecoder(sd, 'OSIE', T, dev, wave=1, use_tensor_cores=mode)
        probs, mu, s2, am = dec.decode(vf.to(dev))
        rel = np.abs(probs[0].double().cpu().numpy() - p64) / p64
        print('   mode', mode, 'max rel per step', ['%.1e' % rel[:, t].max() for t in range(T)])
