"""C-ABI checks that need no GPU: the library loads, exports every symbol the
header declares, and the host-side table helper is bit-exact."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from scanpaths_b200 import build, _lib
    build.build_library()
    return _lib.load()


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", txt))
    return names


def test_every_declared_symbol_is_exported_and_bound(lib):
    from scanpaths_b200 import _lib
    declared = _declared()
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_string(lib):
    assert lib.spb_version() >= 100
    assert isinstance(lib.spb_last_error(), bytes)


def test_bad_arguments_fail_loudly(lib):
    from scanpaths_b200 import _lib
    rc = lib.spb_scanmatch_tables(None, None, None, None, None, None)
    assert rc == -1 and b"cfg is null" in lib.spb_last_error()
    with pytest.raises(_lib.SpbError):
        _lib.check(rc, "spb_scanmatch_tables")


@pytest.mark.parametrize("cfg", [dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Threshold=3.5),
                                 dict(Xres=1024, Yres=768, Xbin=12, Ybin=8, Threshold=3.5),
                                 dict(Xres=1024, Yres=768, Xbin=8, Ybin=6, Threshold=1.5)])
def test_tables_bit_exact_vs_oracle(lib, cfg):
    from scanpaths_b200 import _lib
    from oracle.scoring import ScanMatchOracle
    o = ScanMatchOracle(**cfg)
    c = _lib.ScanMatchCfg(cfg["Xres"], cfg["Yres"], cfg["Xbin"], cfg["Ybin"], cfg["Threshold"], 0.0, 0.0, 0.0, 0.0)
    nb = cfg["Xbin"] * cfg["Ybin"]
    delta = np.zeros(nb); full = np.zeros((nb, nb))
    xl = np.zeros(cfg["Xres"], np.uint8); yl = np.zeros(cfg["Yres"], np.uint8)
    mx = C.c_double()
    assert lib.spb_scanmatch_tables(C.byref(c), _lib.ptr(delta), _lib.ptr(full), _lib.ptr(xl), _lib.ptr(yl),
                                    C.byref(mx)) == 0
    assert np.array_equal(full, o.SubMatrix)
    assert mx.value == np.max(o.SubMatrix)
    assert np.array_equal(yl.astype(np.int64)[:, None] * cfg["Xbin"] + xl.astype(np.int64)[None, :],
                          o.mask.astype(np.int64))
    rows, cols = np.arange(nb) // cfg["Xbin"], np.arange(nb) % cfg["Xbin"]
    dr = np.abs(rows[:, None] - rows[None, :]); dc = np.abs(cols[:, None] - cols[None, :])
    assert np.array_equal(delta[dr * cfg["Xbin"] + dc], o.SubMatrix)


def test_product_path_has_no_oracle_import():
    """The shipped package must never route through oracle/ (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scanpaths_b200")):
        for fn in files:
            if fn.endswith(".py"):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, fn)


def test_cuda_missing_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scanpaths_b200 import _lib
    from scanpaths_b200.utils.evaltools.scanmatch import ScanMatch
    with pytest.raises(_lib.SpbError):
        ScanMatch(Xres=320, Yres=240, Xbin=16, Ybin=12)
