"""Decoder oracle (oracle/decoder.py, float64) against the reference's own
outputs recorded in tests/golden/decoder_*.npz."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder as OD
from scanpaths_b200.weights import random_state_dict, synthetic_features
from golden.make_decoder_goldens import CASES, COCO_TASKS


@pytest.mark.parametrize("name", ["coco", "air", "osie"])
def test_decoder_oracle_matches_reference(golden_dir, name):
    task, n, steps, wseed, fseed, bstd = CASES[name]
    if name == "osie":
        steps = 5            # CPU-suite budget: first 5 of the 16 recorded steps
    g = np.load(os.path.join(golden_dir, "decoder_%s.npz" % name))
    sd = random_state_dict(task, wseed, calibrated=True, bias_std=bstd)
    torch.set_num_threads(os.cpu_count())
    if task == "OSIE":
        vf, att, tasks = synthetic_features(n, fseed), None, None
    else:
        vf, att = synthetic_features(n, fseed, attention=True)
        tasks = COCO_TASKS[:n] if task == "COCO_Search18" else None
    with torch.no_grad():
        out = OD.decode(sd, vf.double(), task, None if att is None else att.double(), tasks, steps=steps)
    for k, v in out.items():
        ref = g["f64_" + k][:, :steps]
        np.testing.assert_allclose(v.numpy(), ref, rtol=1e-10, atol=1e-13, err_msg=k)
