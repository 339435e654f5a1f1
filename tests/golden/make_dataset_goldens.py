"""Dataset golden: instantiate the reference's OSIE_evaluation (OSIE/dataset/dataset.py:150-248) on a tiny
throw-away split (3 images, ragged subject counts) written to a temp dir, and record the fix_vectors its
__getitem__ / collate_func produce together with the JSON they came from."""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refload  # noqa: E402


def gen_dataset():
    for name in ("skimage", "skimage.io", "skimage.transform"):          # import-only in this code path
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    for fn in ("rescale", "resize", "downscale_local_mean"):
        setattr(sys.modules["skimage.transform"], fn, lambda *a, **k: None)
    refload.load_reference("OSIE")
    import importlib
    ds = importlib.import_module("dataset.dataset")
    from PIL import Image
    rng = np.random.default_rng(4)
    records = []
    for img, n_sub in (("1001.jpg", 3), ("1002.jpg", 1), ("1003.jpg", 4)):
        for s in range(n_sub):
            L = int(rng.integers(1, 9))
            records.append({"name": img, "subject": s, "X": [float(v) for v in rng.uniform(0, 800, L).round(1)],
                            "Y": [float(v) for v in rng.uniform(0, 600, L).round(1)],
                            "T": [float(v) for v in rng.integers(80, 900, L)], "length": L})
    with tempfile.TemporaryDirectory() as tmp:
        stim, fix = os.path.join(tmp, "stimuli"), os.path.join(tmp, "fix")
        os.makedirs(stim); os.makedirs(fix)
        for img in ("1001.jpg", "1002.jpg", "1003.jpg"):
            Image.new("RGB", (800, 600)).save(os.path.join(stim, img))
        json.dump(records, open(os.path.join(fix, "osie_fixations_validation.json"), "w"))
        d = ds.OSIE_evaluation(stim, fix, type="validation", transform=lambda im: __import__("torch").zeros(3, 2, 2))
        batch = d.collate_func([d[i] for i in range(len(d))])
    out = {"records_json": np.array(json.dumps(records)), "img_names": np.array(batch["img_names"])}
    for i, fvs in enumerate(batch["fix_vectors"]):
        out["n_sub_%d" % i] = np.array(len(fvs))
        for j, fv in enumerate(fvs):
            out["fix_%d_%d" % (i, j)] = np.stack([fv["start_x"], fv["start_y"], fv["duration"]], 1).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, "dataset_osie.npz"), **out)
    print("dataset golden written:", [int(out["n_sub_%d" % i]) for i in range(3)])


if __name__ == "__main__":
    gen_dataset()
