"""Drop-in for the reference's ``utils/evaltools/scanmatch.py`` on the GPU.

Same constructor keywords, attributes and methods as the reference's
``ScanMatch`` (scanmatch.py:39-197); the work runs in csrc/prep.cu and
csrc/score_pairs.cu.  ``match`` returns ``(score, align, F)`` like the
reference's: the F matrix comes from the device (spb_scanmatch_matrix, bit-identical),
the O(n + m) walk back through it runs on the host.  Every caller in the reference
discards align and F (OSIE/utils/evaluation.py:186,192): for throughput use
``scanpaths_b200.scoring`` / ``utils.evaluation`` (batched, score only); this class
handles one pair per call.
"""
from __future__ import annotations

import numpy as np

from ... import scoring as S


class ScanMatch(object):
    _KEYS = ("Xres", "Yres", "Xbin", "Ybin", "Threshold", "GapValue", "TempBin", "Offset")

    def __init__(self, **kw):
        self.Xres, self.Yres, self.Xbin, self.Ybin = 1024, 768, 8, 6
        self.Threshold, self.GapValue, self.TempBin, self.Offset = 3.5, 0.0, 0.0, (0, 0)
        for k in kw.keys():
            if k not in self._KEYS:
                raise ValueError('Unknown parameter: %s.' % k)          # scanmatch.py:81
            setattr(self, k, kw[k])
        self._cfg = None
        self._custom_mask = None
        self.CreateSubMatrix()
        self.GridMask()

    # -- tables (host helper of the C ABI; bit-equal to the reference's numpy tables)
    def _config(self):
        if self._cfg is None:
            self._cfg = S.ScoreConfig(Xres=self.Xres, Yres=self.Yres, Xbin=self.Xbin, Ybin=self.Ybin,
                                      Threshold=self.Threshold, GapValue=self.GapValue, TempBin=self.TempBin,
                                      Offset=self.Offset, stimulus_shape=(self.Yres, self.Xres, 3), dur_scale=1.0)
            if self._custom_mask is not None:
                self._cfg.set_mask(self._custom_mask)
        return self._cfg

    def CreateSubMatrix(self, Threshold=None):
        if Threshold is not None:
            self.Threshold = Threshold
            self._cfg = None
        self.SubMatrix = self._config().full_sub_matrix()

    def GridMask(self):
        c = self._config()
        self.mask = (c.ylut.astype(np.float64)[:, None] * self.Xbin + c.xlut.astype(np.float64)[None, :])

    # -- scanmatch.py:116-133
    def fixationToSequence(self, data):
        data = np.asarray(data, dtype=np.float64)
        if data.shape[1] == 2:                                           # the w/o-duration call style (:248)
            data = np.concatenate([data, np.zeros((data.shape[0], 1))], 1)
        pack = S.pack_paths([data], self._config())
        L = data.shape[0]
        sym = pack.sym[0, :L].cpu().numpy().astype(np.float64)
        if self.TempBin != 0:
            return np.repeat(sym, pack.run[0, :L].cpu().numpy())
        return sym

    # -- scanmatch.py:135-197
    def match(self, A, B):
        A = np.asarray(A).astype(np.int64).reshape(-1)
        B = np.asarray(B).astype(np.int64).reshape(-1)
        n, m = len(A), len(B)
        F = S.scanmatch_matrix(A, B, self._config())                     # [(n+1), (m+1)], scanmatch.py:138-150
        # walk back from F[n, m] (:152-185): diagonal when the cell is its diagonal neighbour plus the
        # substitution score, else along A when it is F[i-1, j] + GapValue, else along B; -1 marks a gap
        sub, gap = self.SubMatrix, self.GapValue
        i, j, ra, rb = n, m, [], []
        while i > 0 and j > 0:
            if F[i, j] == F[i - 1, j - 1] + sub[A[i - 1], B[j - 1]]:
                ra.append(A[i - 1]); rb.append(B[j - 1]); i -= 1; j -= 1
            elif F[i, j] == F[i - 1, j] + gap:
                ra.append(A[i - 1]); rb.append(-1); i -= 1
            else:
                ra.append(-1); rb.append(B[j - 1]); j -= 1
        ra += list(A[:i][::-1]) + [-1] * j
        rb += [-1] * i + list(B[:j][::-1])
        align = np.array([ra[::-1], rb[::-1]], dtype=np.float64).T.reshape(-1, 2)
        with np.errstate(invalid="ignore", divide="ignore"):
            score = np.float64(np.max(F)) / np.float64(np.max(sub) * max(m, n))        # :190-193
        return score, align, F.transpose()

    def maskFromArray(self, array):
        """scanmatch.py:199-200: `array` [Yres, Xres] replaces the grid mask; fixationToSequence then reads the
        symbol of a fixation from it (K1 takes the table through spb_score_cfg.d_mask)."""
        self.mask = array
        self._custom_mask = np.asarray(array)
        self._config().set_mask(self._custom_mask)

    def subMatrixFromArray(self, array):
        self.SubMarix = array                                            # reference typo kept: it has no effect there either
