"""Drop-in for the reference's ``utils/evaltools/scanmatch.py`` on the GPU.

Same constructor keywords, attributes and methods as the reference's
``ScanMatch`` (scanmatch.py:39-197); the work runs in csrc/prep.cu and
csrc/score_pairs.cu.  ``match`` returns ``(score, None, None)``: every caller in
the reference discards the alignment and the F matrix
(OSIE/utils/evaluation.py:186,192), so only the score is computed.
For throughput use ``scanpaths_b200.scoring`` / ``utils.evaluation`` (batched);
this class scores one pair per call.
"""
from __future__ import annotations

import numpy as np
import torch

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

    # -- scanmatch.py:135-197 (score only)
    def match(self, A, B):
        A = np.asarray(A).astype(np.int64).reshape(-1)
        B = np.asarray(B).astype(np.int64).reshape(-1)
        packs = [_rle_pack(A, self._config()), _rle_pack(B, self._config())]
        if packs[1].lmax > 256 and packs[0].lmax <= 256:                 # NW is symmetric in (A, B)
            packs = packs[::-1]
        dev = self._config().device
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        out = S.score_pairs(packs[0], packs[1], z, z, self._config())
        return float(out[0, 0].item()), None, None

    def maskFromArray(self, array):
        """scanmatch.py:199-200: `array` [Yres, Xres] replaces the grid mask; fixationToSequence then reads the
        symbol of a fixation from it (K1 takes the table through spb_score_cfg.d_mask)."""
        self.mask = array
        self._custom_mask = np.asarray(array)
        self._config().set_mask(self._custom_mask)

    def subMatrixFromArray(self, array):
        self.SubMarix = array                                            # reference typo kept: it has no effect there either


def _rle_pack(seq, cfg):
    """Explicit symbol string -> PathPack in run-length form (sym, run)."""
    dev = cfg.device
    if len(seq) == 0:
        sym, run = np.zeros(1, np.uint8), np.zeros(1, np.int32)
        n = 0
    else:
        cut = np.flatnonzero(np.diff(seq)) + 1
        starts = np.concatenate([[0], cut])
        sym = seq[starts].astype(np.uint8)
        run = np.diff(np.concatenate([starts, [len(seq)]])).astype(np.int32)
        n = len(sym)
    L = len(sym)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    return S.PathPack(xyd=torch.zeros((1, L, 3), dtype=torch.float64, device=dev),
                      len=torch.tensor([n], dtype=torch.int32, device=dev),
                      sym=t(sym[None], torch.uint8), run=t(run[None], torch.int32),
                      nwd=torch.tensor([int(len(seq))], dtype=torch.int32, device=dev),
                      sed=torch.zeros((1, L), dtype=torch.int32, device=dev),
                      xyn=torch.zeros((1, L, 2), dtype=torch.float64, device=dev))
