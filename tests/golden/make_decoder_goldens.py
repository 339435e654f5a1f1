"""Decoder goldens: run the reference ``baseline`` modules (encoder replaced by
identity so that `images` := visual_feature) in float64 and float32 on seeded
weights/features and record the outputs.  Weights and features are regenerated
from their seeds (scanpaths_b200.weights), only outputs are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refload  # noqa: E402
from scanpaths_b200.weights import random_state_dict, synthetic_features  # noqa: E402

CASES = {
    # name: (task, n_images, steps, weight seed, feature seed, bias_std)
    "osie": ("OSIE", 2, 16, 0, 0, 0.02),
    "air": ("AiR", 2, 16, 1, 1, 0.02),
    "coco": ("COCO_Search18", 3, 16, 2, 2, 0.02),
}
COCO_TASKS = [3, 17, 3]


def run_case(name):
    task, n, steps, wseed, fseed, bstd = CASES[name]
    ns = refload.load_reference_model(task)
    model = ns.model.baseline(convLSTM_length=steps)
    model.resnet = torch.nn.Identity()
    model.sal_conv = torch.nn.Identity()
    sd = random_state_dict(task, wseed, calibrated=True, bias_std=bstd)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    model.eval()
    out = {}
    if task == "OSIE":
        vf = synthetic_features(n, fseed)
        args64, args32 = (vf.double(),), (vf,)
    else:
        vf, att = synthetic_features(n, fseed, attention=True)
        if task == "AiR":
            args64, args32 = (vf.double(), att.double()), (vf, att)
        else:
            tasks = torch.tensor(COCO_TASKS[:n])
            args64, args32 = (vf.double(), att.double(), tasks), (vf, att, tasks)
            out["tasks"] = tasks.numpy()
    with torch.no_grad():
        r64 = model.double()(*args64)
        r32 = model.float()(*args32)
    for k, v in r64.items():
        out["f64_" + k] = v.numpy()
        out["f32_" + k] = r32[k].numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_%s.npz" % name), **out)
    p = r64[[k for k in r64 if k.endswith("all_actions_prob")][0]]
    print(name, "stop prob/step:", p[0, :, 0].numpy().round(3))


def gen_decoder():
    torch.set_num_threads(os.cpu_count())
    for name in CASES:
        run_case(name)


if __name__ == "__main__":
    gen_decoder()
