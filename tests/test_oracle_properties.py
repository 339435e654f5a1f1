"""Property tests of the oracle (hypothesis): invariants the domain offers, used as
size-independent checks of the CUDA path as well (tests/test_gpu_scoring.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_scoring as C
from oracle import scoring as O

paths = st.lists(st.tuples(st.floats(-20, 340), st.floats(-20, 260), st.floats(0, 1500)), min_size=1, max_size=10)


def _arr(p):
    return np.array(p, dtype=np.float64).reshape(-1, 3)


@settings(max_examples=40, deadline=None)
@given(paths, paths)
def test_symmetry_and_ranges(a, b):
    a, b = _arr(a), _arr(b)
    wd1, wod1, sed1, stde1 = O.score_pair(a, b)
    wd2, wod2, sed2, _ = O.score_pair(b, a)
    assert np.array_equal(np.float64(wd1), np.float64(wd2), equal_nan=True)   # NW with gap 0 is symmetric
    assert wod1 == wod2 and sed1 == sed2
    assert 0 <= sed1 <= max(len(a), len(b))
    assert abs(len(a) - len(b)) <= sed1
    assert 0 < stde1 <= 1.0
    assert wod1 <= 1.0 + 1e-12


@settings(max_examples=25, deadline=None)
@given(paths)
def test_self_comparison(a):
    a = _arr(a)
    wd, wod, sed, stde = O.score_pair(a, a)
    assert sed == 0 and abs(stde - 1.0) < 1e-12 and abs(wod - 1.0) < 1e-12
    assert np.isnan(wd) or abs(wd - 1.0) < 1e-12          # NaN when every duration bins to zero symbols


@settings(max_examples=25, deadline=None)
@given(paths, paths)
def test_c_oracle_agrees(a, b):
    a, b = _arr(a), _arr(b)
    ref = O.score_pair(a, b)
    out = C.score_pairs(a[None], np.array([len(a)], np.int32), b[None], np.array([len(b)], np.int32), [0], [0])[0]
    assert np.array_equal(np.float64(ref[0]), out[0], equal_nan=True) and ref[1] == out[1] and ref[2] == out[2]
    assert abs(ref[3] - out[3]) <= 1e-13


@settings(max_examples=25, deadline=None)
@given(paths, paths, st.integers(1, 5))
def test_duration_scaling_changes_only_scanmatch_wd(a, b, k):
    """SED, STDE and ScanMatch w/o duration ignore durations entirely."""
    a, b = _arr(a), _arr(b)
    a2, b2 = a.copy(), b.copy()
    a2[:, 2] *= k; b2[:, 2] = b2[:, 2] * k + 7
    r1, r2 = O.score_pair(a, b), O.score_pair(a2, b2)
    assert r1[1:] == r2[1:]
