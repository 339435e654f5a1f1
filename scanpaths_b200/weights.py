"""Seeded random-init decoder weights under the reference's state_dict key names.

There is no network for checkpoints, so tests and bench.py use weights drawn the
way the reference initialises them (xavier-normal convs, N(0, 0.01) linears, zero
biases: /root/reference/OSIE/models/baseline_attention.py:50-57, 399-408), from a
numpy PCG64 stream so that the authoring container and the GPU box regenerate
bit-identical tensors from a seed.  A real checkpoint's ``state_dict`` (keys
``lstm.input_h.weight``, ``object_head.sal_layer_3.weight``, ...) is consumed the
same way (checkpointing.py:79-110 saves exactly these names).
"""
from __future__ import annotations

import math

import numpy as np
import torch

COCO_OBJECTS = ["bottle", "bowl", "car", "chair", "clock", "cup", "fork", "keyboard", "knife",
                "laptop", "microwave", "mouse", "oven", "potted plant", "sink", "stop sign",
                "toilet", "tv"]

E = 512


def decoder_param_shapes(task: str = "OSIE"):
    """Ordered (name, shape, kind) for every decoder-side parameter of `task`."""
    p = []
    gates_h = ["input_x", "forget_x", "output_x", "memory_x", "input_h", "forget_h", "output_h", "memory_h"]
    if task == "AiR":
        gates_m = ["input_pos", "forget_pos", "output_pos", "input_neg", "forget_neg", "output_neg"]
    else:
        gates_m = ["input", "forget", "output"]
    for g in gates_h + gates_m:
        p.append(("lstm.%s" % g, (E, E, 3, 3), "conv"))
    p.append(("semantic_embed", (E, E), "linear"))
    p.append(("spatial_embed", (1200, 1200), "linear"))
    p.append(("semantic_att.semantic_lists", (E, E), "linear"))
    p.append(("semantic_att.semantic_cur", (E, E), "linear"))
    p.append(("semantic_att.semantic_attention", (1, E), "linear"))
    p.append(("spatial_att.spatial_lists", (1, 1, 3, 3), "conv"))
    p.append(("spatial_att.spatial_cur", (1, 1, 3, 3), "conv"))
    p.append(("spatial_att.spatial_attention", (1, 1, 30, 40), "conv"))
    if task == "AiR":
        for k in ("False", "True"):
            p.append(("performance_sal_layer.%s" % k, (E, E, 5, 5), "conv"))
    elif task == "COCO_Search18":
        for k in COCO_OBJECTS:
            p.append(("object_sal_layer.%s" % k, (E, E, 5, 5), "conv"))
    else:
        p.append(("performance_sal_layer", (E, E, 5, 5), "conv"))
    p.append(("object_head.sal_layer_2", (1, E, 1, 1), "conv"))
    p.append(("object_head.sal_layer_3", (1, E, 1, 1), "conv"))
    p.append(("object_head.drt_layer_1", (1, E, 7, 7), "conv"))
    p.append(("object_head.drt_layer_2", (2, 1, 6, 8), "conv"))
    return p


def random_state_dict(task: str = "OSIE", seed: int = 0, calibrated: bool = True, bias_std: float = 0.0):
    """float32 CPU tensors.  `calibrated` applies the stated bias calibration of
    SURVEY.md section 8d so the outputs are data-like: durations ~ log-normal
    around 0.25 s (drt_layer_2.bias = (log 0.25, log 0.15)) and a per-step stop
    probability of roughly 0.05-0.15 (sal_layer_2.bias).  `bias_std` > 0 draws
    non-zero biases everywhere (exercises the bias paths in tests)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, kind in decoder_param_shapes(task):
        if kind == "conv":
            rf = shape[2] * shape[3]
            std = math.sqrt(2.0 / (shape[1] * rf + shape[0] * rf))
        else:
            std = 0.01
        w = rng.standard_normal(shape, dtype=np.float32) * np.float32(std)
        b = (rng.standard_normal(shape[0], dtype=np.float32) * np.float32(bias_std)) if bias_std > 0 \
            else np.zeros(shape[0], dtype=np.float32)
        sd[name + ".weight"] = torch.from_numpy(w)
        sd[name + ".bias"] = torch.from_numpy(b)
    if calibrated:
        sd["object_head.drt_layer_2.bias"] = sd["object_head.drt_layer_2.bias"] + torch.tensor(
            [math.log(0.25), math.log(0.15)], dtype=torch.float32)
        sd["object_head.sal_layer_2.bias"] = sd["object_head.sal_layer_2.bias"] + torch.tensor(
            [STOP_BIAS], dtype=torch.float32)
    return sd


# sal_layer_2 bias giving a stop share of roughly 0.05-0.15 per step with the
# synthetic relu(N(0,1)) features (measured, see DESIGN.md "calibration")
STOP_BIAS = 5.0


def synthetic_features(n: int, seed: int = 0, attention: bool = False):
    """visual_feature [n,512,30,40] = relu(N(0,1)) (encoder bypassed),
    optional attention map [n,1,30,40] ~ U(0,1)/max."""
    rng = np.random.default_rng(seed + 1000003)
    vf = np.maximum(rng.standard_normal((n, E, 30, 40), dtype=np.float32), 0)
    out = [torch.from_numpy(vf)]
    if attention:
        a = rng.random((n, 1, 30, 40), dtype=np.float32)
        a /= a.reshape(n, -1).max(1).reshape(n, 1, 1, 1)
        out.append(torch.from_numpy(a))
    return out if attention else out[0]
