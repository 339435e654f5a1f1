"""Drop-in for the scanpath metrics of the reference's
``utils/evaltools/visual_attention_metrics.py`` (SED :301-317, STDE :393-441 and
the STDE family's other entry points: euclidean_distance :205-218,
time_delay_embedding_distance :332-390, scaled_time_delay_embedding_distance
:444-492) on the GPU (csrc/prep.cu + csrc/score_pairs.cu).  The saliency-map metrics of that
file (AUC_Judd, KLdiv, NSS) are never called by the reference's pipelines and are
out of scope.  One pair per call; the batched path is ``scanpaths_b200.scoring``.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import scoring as S

_cfg_cache = {}


def _config(shape, n=5):
    key = (tuple(int(v) for v in shape), int(n))
    if key not in _cfg_cache:
        _cfg_cache[key] = S.ScoreConfig(Xres=max(int(shape[1]), 1), Yres=max(int(shape[0]), 1), Xbin=1, Ybin=1,
                                        stimulus_shape=key[0], sed_n=n, dur_scale=1.0)
    return _cfg_cache[key]


def _score(human, simulated, shape, n=5):
    cfg = _config(shape, n)
    h = np.asarray(human, dtype=np.float64)
    s = np.asarray(simulated, dtype=np.float64)

    def three(a):
        a = a.reshape(-1, a.shape[-1]) if a.ndim == 2 else a.reshape(0, 3)
        if a.shape[1] == 2:
            a = np.concatenate([a, np.zeros((a.shape[0], 1))], 1)
        return a[:, :3]
    hp, sp = S.pack_paths([three(h)], cfg), S.pack_paths([three(s)], cfg)
    z = torch.zeros(1, dtype=torch.int32, device=cfg.device)
    return S.score_pairs(hp, sp, z, z, cfg)[0].cpu().numpy()


def string_edit_distance(stimulus, human_scanpath, simulated_scanpath, n=5, substitution_cost=1, msg=False):
    """visual_attention_metrics.py:301-317 (the reference ignores substitution_cost
    too: it calls _Levenshtein with the default unit cost)."""
    return int(_score(human_scanpath, simulated_scanpath, np.shape(stimulus), n)[2])


def scaled_time_delay_embedding_similarity(human_scanpath, simulated_scanpath, image, toPlot=False, msg=False):
    """visual_attention_metrics.py:393-441; None when either scanpath is empty."""
    v = float(_score(human_scanpath, simulated_scanpath, np.shape(image))[3])
    return None if np.isnan(v) else v


def _xy(scanpath):
    a = np.asarray(scanpath, dtype=np.float64)
    return a.reshape(-1, a.shape[-1])[:, :2] if a.size else np.zeros((0, 2))


def euclidean_distance(human_scanpath, simulated_scanpath, msg=False):
    """visual_attention_metrics.py:205-218: the sum of the point distances of two equally long scanpaths, else
    False."""
    if len(human_scanpath) != len(simulated_scanpath):
        if msg:
            print('Error: The two sequences must have the same length!')
        return False
    if len(human_scanpath) == 0:
        return 0.0
    return float(S.tde_table(_xy(human_scanpath), _xy(simulated_scanpath))[-1, 2])


def time_delay_embedding_distance(human_scanpath, simulated_scanpath, k=3, distance_mode='Mean', msg=False):
    """visual_attention_metrics.py:332-390; False when k exceeds a scanpath's length or the mode is unknown."""
    if len(human_scanpath) < k or len(simulated_scanpath) < k:
        if msg:
            print('ERROR: Too large value for the time-embedding vector dimension')
        return False
    if distance_mode not in ('Mean', 'Hausdorff'):
        if msg:
            print('ERROR: distance mode not defined.')
        return False
    row = S.tde_table(_xy(human_scanpath), _xy(simulated_scanpath))[k - 1]
    return float(row[0] if distance_mode == 'Mean' else row[1])


def scaled_time_delay_embedding_distance(human_scanpath, simulated_scanpath, image, toPlot=False, msg=False):
    """visual_attention_metrics.py:444-492: coordinates rescaled by the image's largest dimension, then the mean
    over k of the 'Mean' distances; None when either scanpath is empty."""
    max_dim = float(max(np.shape(image)))
    t = S.tde_table(_xy(human_scanpath) / max_dim, _xy(simulated_scanpath) / max_dim)
    if len(t) == 0:
        return None
    return float(sum(t[:, 0]) / len(t))
