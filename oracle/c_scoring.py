"""ctypes wrapper over oracle/c/libscoring_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = None


class Cfg(ctypes.Structure):
    _fields_ = [("Xres", ctypes.c_int), ("Yres", ctypes.c_int), ("Xbin", ctypes.c_int), ("Ybin", ctypes.c_int),
                ("Threshold", ctypes.c_double), ("GapValue", ctypes.c_double), ("TempBin", ctypes.c_double),
                ("OffsetX", ctypes.c_double), ("OffsetY", ctypes.c_double)]


def build(force: bool = False):
    so = os.path.join(_DIR, "libscoring_oracle.so")
    src = os.path.join(_DIR, "scoring_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-B", "libscoring_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.spo_nw_score.restype = ctypes.c_double
        _LIB.spo_stde.restype = ctypes.c_double
        _LIB.spo_fixation_to_sequence.restype = ctypes.c_long
    return _LIB


def eval_cfg(tempbin=50.0):
    return Cfg(320, 240, 16, 12, 3.5, 0.0, float(tempbin), 0.0, 0.0)


def score_pairs(human, hlen, pred, plen, gi, pi, cfg=None, height=240, width=320, threads=1):
    """human [H,Lh,3] f64 (x,y,ms) + lengths, pred [P,Lp,3] + lengths, index arrays
    gi/pi [npairs] -> [npairs,4] f64 (SM-wd, SM-wod, SED, STDE)."""
    cfg = cfg or eval_cfg()
    human = np.ascontiguousarray(human, dtype=np.float64)
    pred = np.ascontiguousarray(pred, dtype=np.float64)
    hlen = np.ascontiguousarray(hlen, dtype=np.int32)
    plen = np.ascontiguousarray(plen, dtype=np.int32)
    gi = np.ascontiguousarray(gi, dtype=np.int64)
    pi = np.ascontiguousarray(pi, dtype=np.int64)
    out = np.zeros((len(gi), 4), dtype=np.float64)
    P = ctypes.c_void_p
    fn = lib().spo_score_pairs

    def run(lo, hi):
        fn(ctypes.byref(cfg), P(human.ctypes.data), P(hlen.ctypes.data), ctypes.c_int(human.shape[1]),
           P(pred.ctypes.data), P(plen.ctypes.data), ctypes.c_int(pred.shape[1]),
           P(gi.ctypes.data + 8 * lo), P(pi.ctypes.data + 8 * lo), ctypes.c_int64(hi - lo),
           ctypes.c_int(height), ctypes.c_int(width), P(out.ctypes.data + 32 * lo))

    n = len(gi)
    if threads <= 1 or n < 4 * threads:
        run(0, n)
    else:
        from concurrent.futures import ThreadPoolExecutor
        cuts = np.linspace(0, n, threads * 8 + 1).astype(np.int64)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda k: run(int(cuts[k]), int(cuts[k + 1])), range(threads * 8)))
    return out
