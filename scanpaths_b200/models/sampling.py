"""Drop-in for the reference's ``models/sampling.py`` (Sampling.random_sample :16-46,
Sampling.generate_scanpath :48-77) on the GPU (csrc/sample.cu).

``random_sample`` / ``generate_scanpath`` keep the reference's signatures and
return types.  ``sample_paths`` is the batched form the evaluation and SCST loops
want: K samples per image in one launch, the predicted scanpaths left packed on
the device for the scoring kernels (no per-image ``.cpu()`` round trips).
Random draws come from Philox4x32-10 streams keyed by ``seed`` and a call
counter; pass ``q`` (Exp(1), shape [..., N, T, A]) and ``z`` (N(0,1), [..., N, T])
to inject the reference's own draws, in which case the samples are identical.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

FIX_DTYPE = {'names': ('start_x', 'start_y', 'duration'), 'formats': ('f8', 'f8', 'f8')}


class Sampling():
    def __init__(self, convLSTM_length=16, min_length=2, map_width=40, map_height=30, width=320, height=240,
                 seed=None):
        self.convLSTM_length = convLSTM_length
        self.min_length = min_length
        self.map_width = map_width
        self.map_height = map_height
        self.width = width
        self.height = height
        self.x_granularity = float(self.width / self.map_width)
        self.y_granularity = float(self.height / self.map_height)
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0

    def _geom(self):
        return _lib.SampleGeom(self.map_width, self.map_height, self.width, self.height, self.min_length, 0)

    # ------------------------------------------------------------------ batched
    def sample_paths(self, all_actions_prob, log_normal_mu, log_normal_sigma2, K=1, q=None, z=None):
        """K samples per image.  Returns a dict of device tensors:
        selected_actions i32 [K,N,T], selected_actions_probs / durations /
        action_masks / duration_masks f32 [K,N,T], scanpath_length f32 [K,N],
        xyd f64 [K*N,T,3] (x, y, seconds), len i32 [K*N] (sample-major: k*N + image)."""
        _lib.require_cuda()
        lib = _lib.load()
        probs = all_actions_prob.detach().float().contiguous()
        mu = log_normal_mu.detach().float().contiguous()
        s2 = log_normal_sigma2.detach().float().contiguous()
        assert probs.is_cuda, "scanpaths_b200 has no CPU path: tensors must be on the GPU"
        N, T, A = probs.shape
        dev = probs.device
        f32 = dict(dtype=torch.float32, device=dev)
        out = {
            "selected_actions": torch.empty((K, N, T), dtype=torch.int32, device=dev),
            "selected_actions_probs": torch.empty((K, N, T), **f32),
            "durations": torch.empty((K, N, T), **f32),
            "action_masks": torch.empty((K, N, T), **f32),
            "duration_masks": torch.empty((K, N, T), **f32),
            "scanpath_length": torch.empty((K, N), **f32),
            "xyd": torch.empty((K * N, T, 3), dtype=torch.float64, device=dev),
            "len": torch.empty((K * N,), dtype=torch.int32, device=dev),
        }
        if q is not None:
            q = q.detach().float().contiguous().reshape(K, N, T, A)
        if z is not None:
            z = z.detach().float().contiguous().reshape(K, N, T)
        seed = (self.seed + 0x9E3779B97F4A7C15 * self._calls) & 0xFFFFFFFFFFFFFFFF
        self._calls += 1
        geom = self._geom()
        with torch.cuda.device(dev):
            _lib.check(lib.spb_sample_paths(
                _lib.ptr(probs), _lib.ptr(mu), _lib.ptr(s2), _lib.ptr(q), _lib.ptr(z), C.c_uint64(seed), N, T, A, K,
                C.byref(geom), _lib.ptr(out["selected_actions"]), _lib.ptr(out["selected_actions_probs"]),
                _lib.ptr(out["durations"]), _lib.ptr(out["action_masks"]), _lib.ptr(out["duration_masks"]),
                _lib.ptr(out["scanpath_length"]), _lib.ptr(out["xyd"]), _lib.ptr(out["len"]),
                _lib.current_stream()), "spb_sample_paths")
        return out

    # ---------------------------------------------------------- reference API
    def random_sample(self, all_actions_prob, log_normal_mu, log_normal_sigma2, q=None, z=None):
        s = self.sample_paths(all_actions_prob, log_normal_mu, log_normal_sigma2, 1, q, z)
        actions = s["selected_actions"][0].long()
        if all_actions_prob.requires_grad:      # SCST: the gradient flows through this gather (train.py:242)
            sel = torch.gather(all_actions_prob, dim=2, index=actions.unsqueeze(-1)).squeeze(-1)
        else:
            sel = s["selected_actions_probs"][0]
        predicts = {}
        predicts["scanpath_length"] = s["scanpath_length"][0].unsqueeze(-1)
        predicts["durations"] = s["durations"][0]
        predicts["selected_actions_probs"] = sel
        predicts["selected_actions"] = actions
        return predicts

    def generate_scanpath(self, images, prob_sample_actions, durations, sample_actions):
        packed = self.generate_scanpath_packed(sample_actions, durations)
        xyd = packed["xyd"].cpu().numpy()
        lens = packed["len"].cpu().numpy()
        fix = []
        for n in range(xyd.shape[0]):
            rows = [tuple(r) for r in xyd[n, :lens[n]]]
            fix.append(np.array(rows, dtype=FIX_DTYPE))
        am = packed["action_masks"].to(images.dtype) if torch.is_tensor(images) else packed["action_masks"]
        dm = packed["duration_masks"].to(images.dtype) if torch.is_tensor(images) else packed["duration_masks"]
        return fix, am, dm

    def generate_scanpath_packed(self, sample_actions, durations):
        """generate_scanpath without leaving the device: masks + packed scanpaths."""
        _lib.require_cuda()
        lib = _lib.load()
        acts = sample_actions.detach().to(torch.int32).contiguous()
        dur = durations.detach().float().contiguous()
        assert acts.is_cuda
        n, T = acts.shape
        dev = acts.device
        out = {"action_masks": torch.empty((n, T), dtype=torch.float32, device=dev),
               "duration_masks": torch.empty((n, T), dtype=torch.float32, device=dev),
               "scanpath_length": torch.empty((n,), dtype=torch.float32, device=dev),
               "xyd": torch.empty((n, T, 3), dtype=torch.float64, device=dev),
               "len": torch.empty((n,), dtype=torch.int32, device=dev)}
        geom = self._geom()
        with torch.cuda.device(dev):
            _lib.check(lib.spb_generate_scanpaths(_lib.ptr(acts), _lib.ptr(dur), n, T, C.byref(geom),
                                                  _lib.ptr(out["action_masks"]), _lib.ptr(out["duration_masks"]),
                                                  _lib.ptr(out["scanpath_length"]), _lib.ptr(out["xyd"]),
                                                  _lib.ptr(out["len"]), _lib.current_stream()),
                       "spb_generate_scanpaths")
        return out


def predictions_to_records(sampled, img_names, n_images):
    """The prediction records test.py:135-148 dumps to JSON, built from a packed ``sample_paths`` result
    (sample-major order k*N + image) with one device->host read: list of dicts
    {name, repeat_id, X, Y, T (ms), length}."""
    xyd = sampled["xyd"].cpu().numpy()
    lens = sampled["len"].cpu().numpy()
    out = []
    for idx in range(xyd.shape[0]):
        k, n = divmod(idx, n_images)
        L = int(lens[idx])
        out.append({"name": img_names[n], "repeat_id": k + 1, "X": list(xyd[idx, :L, 0]), "Y": list(xyd[idx, :L, 1]),
                    "T": list(xyd[idx, :L, 2] * 1000), "length": L})
    return out
