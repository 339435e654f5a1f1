"""f2 host logic (no GPU): the packed layout the evaluation datasets' fix_vectors go into
(OSIE/dataset/dataset.py:196-248 -> scoring.pack_subject_lists) and the prediction records test.py dumps
(OSIE/test.py:135-148 -> models.sampling.predictions_to_records), against the reference's own
generate_scanpath output recorded in tests/golden/sampling.npz."""
import os

import numpy as np
import torch

FIX_DTYPE = {'names': ('start_x', 'start_y', 'duration'), 'formats': ('f8', 'f8', 'f8')}


def _struct(a):
    return np.array([tuple(r) for r in a], dtype=FIX_DTYPE)


def test_pack_subject_lists_layout():
    from scanpaths_b200.scoring import pack_subject_lists, structured_to_xyd
    rng = np.random.default_rng(0)
    counts, lists = [3, 1, 4], []
    for c in counts:
        lists.append([_struct(rng.uniform(0, 300, (int(rng.integers(1, 9)), 3))) for _ in range(c)])
    lists[2][1] = _struct(np.zeros((0, 3)))                         # an empty scanpath stays a real (length 0) subject
    xyd, lens, nsub = pack_subject_lists(lists, pin=False)
    assert xyd.dtype == torch.float64 and lens.dtype == torch.int32 and nsub.tolist() == counts
    lmax = max(len(s) for f in lists for s in f)
    assert tuple(xyd.shape) == (3, 4, lmax, 3) and tuple(lens.shape) == (3, 4)
    for i, fvs in enumerate(lists):
        for j in range(4):
            if j < len(fvs):
                a = structured_to_xyd(fvs[j])
                assert lens[i, j] == len(a)
                np.testing.assert_array_equal(xyd[i, j, :len(a)].numpy(), a)      # bit-exact f8 copy
                assert (xyd[i, j, len(a):] == 0).all()
            else:
                assert lens[i, j] == 0 and (xyd[i, j] == 0).all()                 # padding subject


def test_predictions_to_records_match_test_py(golden_dir):
    """test.py:135-148 on the reference's own generate_scanpath output == the records built from the packed
    sample_paths layout (sample-major k*N + image, seconds -> ms)."""
    from scanpaths_b200.models.sampling import predictions_to_records
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    tags = ["m1_t0_", "m1_t1_", "m1_t2_"]
    N = g[tags[0] + "fix"].shape[0]
    names = ["img_%d.jpg" % i for i in range(N)]
    xyd = np.concatenate([g[t + "fix"] for t in tags], 0)           # [K*N, 16, 3]
    lens = np.concatenate([g[t + "fix_len"] for t in tags], 0)
    recs = predictions_to_records({"xyd": torch.from_numpy(xyd), "len": torch.from_numpy(lens)}, names, N)
    # the reference's loop, verbatim semantics
    expect = []
    for trial, t in enumerate(tags):
        for index in range(N):
            L = int(g[t + "fix_len"][index])
            fix_vector_array = np.array(_struct(g[t + "fix"][index, :L]).tolist()).reshape(-1, 3)
            expect.append({"name": names[index], "repeat_id": trial + 1, "X": list(fix_vector_array[:, 0]),
                           "Y": list(fix_vector_array[:, 1]), "T": list(fix_vector_array[:, 2] * 1000),
                           "length": L})
    assert len(recs) == len(expect)
    for r, e in zip(recs, expect):
        assert r["name"] == e["name"] and r["repeat_id"] == e["repeat_id"] and r["length"] == e["length"]
        assert r["X"] == e["X"] and r["Y"] == e["Y"] and r["T"] == e["T"]
    import json
    json.dumps(recs)                                                 # test.py:151 dumps them as JSON


def test_dataset_mirror_matches_reference_dataset(golden_dir, tmp_path):
    """scanpaths_b200.dataset against the reference's OSIE_evaluation.__getitem__ / collate_func output recorded
    in tests/golden/dataset_osie.npz (same JSON records in, bit-equal f8 rows out), the packed layout built from
    it, and the JSON writer of test.py:151-152."""
    import json
    from scanpaths_b200 import dataset as D
    g = np.load(os.path.join(golden_dir, "dataset_osie.npz"))
    records = json.loads(str(g["records_json"]))
    groups = D.group_fixations_by_image(records)
    assert list(groups) == [str(n) for n in g["img_names"]]
    lists = []
    for i, (name, idx) in enumerate(groups.items()):
        fvs = D.image_fix_vectors(records, idx)
        assert len(fvs) == int(g["n_sub_%d" % i])
        for j, fv in enumerate(fvs):
            ref = g["fix_%d_%d" % (i, j)]
            got = np.stack([fv["start_x"], fv["start_y"], fv["duration"]], 1).reshape(-1, 3)
            np.testing.assert_array_equal(got, ref)                      # float32 division, stored as f8: bit-equal
        lists.append(fvs)
    # collate wrapper: reference dict + the packed layout
    ref_collate = lambda batch: {"fix_vectors": [b["fix_vectors"] for b in batch], "img_names": [b["img_name"] for b in batch]}
    coll = D.PackedCollate(ref_collate, pin=False)
    b1 = coll([{"fix_vectors": lists[0], "img_name": "a"}, {"fix_vectors": lists[1], "img_name": "b"}])
    b2 = coll([{"fix_vectors": lists[2], "img_name": "c"}])
    xyd, lens, nsub = D.concat_packed([b1["fix_packed"], b2["fix_packed"]])
    assert nsub.tolist() == [3, 1, 4] and tuple(xyd.shape[:2]) == (3, 4)
    for i in range(3):
        for j in range(int(nsub[i])):
            ref = g["fix_%d_%d" % (i, j)]
            assert int(lens[i, j]) == len(ref)
            np.testing.assert_array_equal(xyd[i, j, :len(ref)].numpy(), ref)
        assert (lens[i, int(nsub[i]):] == 0).all()
    # prediction writer
    sampled = {"xyd": torch.tensor([[[4.0, 12.0, 0.25], [36.0, 20.0, 0.5]], [[100.0, 60.0, 0.125], [0, 0, 0]]], dtype=torch.float64),
               "len": torch.tensor([2, 1], dtype=torch.int32)}
    path = tmp_path / "predicts.json"
    D.write_prediction_records(str(path), sampled, ["x.jpg"], 1)
    back = json.load(open(path))
    assert back == [{"name": "x.jpg", "repeat_id": 1, "X": [4.0, 36.0], "Y": [12.0, 20.0], "T": [250.0, 500.0], "length": 2},
                    {"name": "x.jpg", "repeat_id": 2, "X": [100.0], "Y": [60.0], "T": [125.0], "length": 1}]
