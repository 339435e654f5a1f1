"""The C restatement of the scoring oracle against the reference goldens and the
Python oracle."""
import os

import numpy as np

from oracle import c_scoring as C
from oracle import scoring as O


def test_c_oracle_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring_random.npz"))
    n = len(g["gt_len"])
    idx = np.arange(n)
    out = C.score_pairs(g["gt"], g["gt_len"], g["pred"], g["pred_len"], idx, idx)
    assert np.array_equal(out[:, 0], g["wd"], equal_nan=True)       # f64 DP in the same order: bit-exact
    assert np.array_equal(out[:, 1], g["wod"])
    assert np.array_equal(out[:, 2].astype(np.int64), g["sed"])
    np.testing.assert_allclose(out[:, 3], g["stde"], rtol=1e-13)


def test_c_oracle_mat_eval_cfg(golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring_mat.npz"))
    data = [g["data%d" % i] * [0.3125, 0.3125, 1.0] for i in (1, 2, 3)]
    lmax = max(len(d) for d in data)
    arr = np.zeros((3, lmax, 3)); lens = np.array([len(d) for d in data], np.int32)
    for i, d in enumerate(data):
        arr[i, :len(d)] = d
    rows = g["mat_eval"]
    out = C.score_pairs(arr, lens, arr, lens, rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64))
    assert np.array_equal(out[:, 0], rows[:, 2]) and np.array_equal(out[:, 1], rows[:, 3])
    assert np.array_equal(out[:, 2], rows[:, 4])
    np.testing.assert_allclose(out[:, 3], rows[:, 5], rtol=1e-13)


def test_c_oracle_vs_python_oracle_random():
    rng = np.random.default_rng(5)
    H, P, L = 40, 40, 18
    hum = np.zeros((H, L, 3)); prd = np.zeros((P, L, 3))
    hl = rng.integers(1, L + 1, H).astype(np.int32); pl = rng.integers(1, L + 1, P).astype(np.int32)
    hum[..., 0] = rng.uniform(-20, 340, (H, L)); hum[..., 1] = rng.uniform(-20, 260, (H, L))
    hum[..., 2] = rng.uniform(0, 900, (H, L))
    prd[..., 0] = rng.integers(0, 40, (P, L)) * 8 + 4; prd[..., 1] = rng.integers(0, 30, (P, L)) * 8 + 4
    prd[..., 2] = np.float32(rng.uniform(0, 900, (P, L)))
    idx = np.arange(H)
    out = C.score_pairs(hum, hl, prd, pl, idx, idx)
    for i in range(H):
        wd, wod, sed, stde = O.score_pair(hum[i, :hl[i]], prd[i, :pl[i]])
        assert np.array_equal(np.float64(wd), out[i, 0], equal_nan=True)
        assert wod == out[i, 1] and sed == out[i, 2]
        assert abs(stde - out[i, 3]) <= 1e-13 * abs(stde)
