"""f2: the data formats on either side of the hot path.

IN  -- the evaluation / RL datasets of the reference hand the scoring loop a Python list per image of
       per-subject structured arrays (``OSIE/dataset/dataset.py:196-248`` ``OSIE_evaluation.__getitem__`` /
       ``collate_func``; ``COCO_Search18/dataset/dataset.py:283-368``; ``AiR/dataset/dataset.py`` likewise).
       ``fixation_record_to_fix_vector`` is that conversion (float32 division of the JSON coordinates by the
       resize scale, milliseconds -> seconds, f8 structured rows); ``PackedCollate`` wraps a reference
       ``collate_func`` so that every batch ALSO carries the dense pinned layout the CUDA path consumes
       (``fix_packed = (xyd [N,S,Lmax,3] f64, lens [N,S] i32, n_subjects [N] i32)``), built once in the loader
       worker instead of per metric call; ``concat_packed`` joins the batches of a whole split.
OUT -- ``write_prediction_records`` is the JSON dump of ``OSIE/test.py:135-152`` fed from the packed device
       output of ``Sampling.sample_paths`` (``models.sampling.predictions_to_records``).
Host code only (numpy / torch CPU tensors); nothing here launches a kernel.
"""
from __future__ import annotations

import json

import numpy as np
import torch

from .models.sampling import FIX_DTYPE, predictions_to_records
from .scoring import pack_subject_lists


def fixation_record_to_fix_vector(fixation, resizescale_x, resizescale_y):
    """One JSON fixation record {"X", "Y", "T" (ms), "length"} -> the reference's structured array
    (start_x, start_y, duration[s]) exactly as the datasets build it: the divisions happen in float32
    (dataset.py:207-209), the rows are stored as f8 (:216-217)."""
    x = np.array(fixation["X"]).astype(np.float32) / resizescale_x
    y = np.array(fixation["Y"]).astype(np.float32) / resizescale_y
    t = np.array(fixation["T"]).astype(np.float32) / 1000.0
    n = int(fixation["length"])
    out = np.zeros((n,), dtype=FIX_DTYPE)
    out["start_x"], out["start_y"], out["duration"] = x[:n], y[:n], t[:n]
    return out


def group_fixations_by_image(fixations):
    """imgid_to_sub of the reference datasets (dataset.py:183-186): image name -> record indices, in file order."""
    groups = {}
    for index, fixation in enumerate(fixations):
        groups.setdefault(fixation["name"], []).append(index)
    return groups


def image_fix_vectors(fixations, indices, origin_size=(600, 800), resize=(240, 320)):
    """``__getitem__``'s fix_vectors of one image (OSIE geometry by default: 800x600 -> 320x240)."""
    sx, sy = origin_size[1] / resize[1], origin_size[0] / resize[0]
    return [fixation_record_to_fix_vector(fixations[i], sx, sy) for i in indices]


class PackedCollate:
    """Drop-in for a reference dataset's ``collate_func``: same dict, plus ``fix_packed``."""

    def __init__(self, collate_func, pin=True):
        self.collate_func, self.pin = collate_func, pin

    def __call__(self, batch):
        data = self.collate_func(batch)
        data["fix_packed"] = pack_subject_lists(data["fix_vectors"], pin=self.pin)
        return data


def concat_packed(packs):
    """[(xyd, lens, nsub), ...] of several batches -> one packed set (subjects / lengths padded to the maxima)."""
    smax = max(p[0].shape[1] for p in packs)
    lmax = max(p[0].shape[2] for p in packs)
    n = sum(p[0].shape[0] for p in packs)
    xyd = torch.zeros((n, smax, lmax, 3), dtype=torch.float64)
    lens = torch.zeros((n, smax), dtype=torch.int32)
    o = 0
    for x, l, _ in packs:
        xyd[o:o + x.shape[0], :x.shape[1], :x.shape[2]] = x
        lens[o:o + x.shape[0], :l.shape[1]] = l
        o += x.shape[0]
    return xyd, lens, torch.cat([p[2] for p in packs], 0)


def write_prediction_records(path, sampled, img_names, n_images):
    """test.py:135-152: one record per (trial, image) with X, Y, T (ms), length, dumped with indent=2."""
    records = predictions_to_records(sampled, img_names, n_images)
    with open(path, "w") as f:
        json.dump(records, f, indent=2)
    return records
