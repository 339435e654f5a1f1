"""Batched scoring on the GPU: packed scanpaths -> (ScanMatch-wd, ScanMatch-wod, SED, STDE).

Host side of kernels K1-K4 (csrc/prep.cu, csrc/score_pairs.cu).  The mirror
modules under ``scanpaths_b200.utils`` (reference names and signatures) are thin
wrappers over this file.  torch is used for device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

# evaluation.py:159-162 -- the configuration every driver of the reference uses
EVAL_SCANMATCH = dict(Xres=320, Yres=240, Xbin=16, Ybin=12, Offset=(0, 0), Threshold=3.5)
EVAL_TEMPBIN = 50
EVAL_STIMULUS = (240, 320, 3)


class ScoreConfig:
    """ScanMatch tables + SED / STDE geometry, resident on one device."""

    def __init__(self, Xres=1024, Yres=768, Xbin=8, Ybin=6, Threshold=3.5, GapValue=0.0, TempBin=0.0,
                 Offset=(0, 0), stimulus_shape=EVAL_STIMULUS, sed_n=5, dur_scale=1.0, device=None):
        _lib.require_cuda()
        lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self.sm = _lib.ScanMatchCfg(int(Xres), int(Yres), int(Xbin), int(Ybin), float(Threshold), float(GapValue),
                                    float(TempBin), float(Offset[0]), float(Offset[1]))
        nb = int(Xbin) * int(Ybin)
        self.sub_delta = np.zeros(nb, dtype=np.float64)
        self.xlut = np.zeros(int(Xres), dtype=np.uint8)
        self.ylut = np.zeros(int(Yres), dtype=np.uint8)
        mx = C.c_double(0.0)
        _lib.check(lib.spb_scanmatch_tables(C.byref(self.sm), _lib.ptr(self.sub_delta), None, _lib.ptr(self.xlut),
                                            _lib.ptr(self.ylut), C.byref(mx)), "spb_scanmatch_tables")
        self.max_sub = mx.value
        self.d_sub_delta = torch.from_numpy(self.sub_delta).to(self.device)
        self.d_xlut = torch.from_numpy(self.xlut).to(self.device)
        self.d_ylut = torch.from_numpy(self.ylut).to(self.device)
        h, w = int(stimulus_shape[0]), int(stimulus_shape[1])
        self.d_mask = None
        self.cfg = _lib.ScoreCfg(self.sm, h, w, int(sed_n), 0, float(max(stimulus_shape)), float(dur_scale),
                                 self.max_sub, self.d_sub_delta.data_ptr(), self.d_xlut.data_ptr(),
                                 self.d_ylut.data_ptr(), None)

    def set_mask(self, array):
        """ScanMatch.maskFromArray (scanmatch.py:199-200): a [Yres, Xres] pixel -> symbol table replaces the
        regular grid.  Symbols must stay below Xbin*Ybin (they index the substitution table)."""
        m = np.ascontiguousarray(np.asarray(array))
        if m.shape != (self.sm.Yres, self.sm.Xres):
            raise ValueError("mask must be [Yres, Xres] = [%d, %d]" % (self.sm.Yres, self.sm.Xres))
        if m.min() < 0 or m.max() >= self.sm.Xbin * self.sm.Ybin:
            raise ValueError("mask symbols must lie in [0, Xbin*Ybin)")
        self.d_mask = torch.from_numpy(m.astype(np.uint8)).to(self.device)
        self.cfg.d_mask = self.d_mask.data_ptr()

    @classmethod
    def evaluation(cls, device=None, dur_scale=1000.0):
        """The drivers' configuration: 320x240, 16x12 bins, TempBin 50 ms, durations given in seconds."""
        return cls(TempBin=EVAL_TEMPBIN, stimulus_shape=EVAL_STIMULUS, dur_scale=dur_scale, device=device,
                   **EVAL_SCANMATCH)

    def full_sub_matrix(self):
        nb = self.sm.Xbin * self.sm.Ybin
        full = np.zeros((nb, nb), dtype=np.float64)
        _lib.check(_lib.load().spb_scanmatch_tables(C.byref(self.sm), None, _lib.ptr(full), None, None, None),
                   "spb_scanmatch_tables")
        return full


@dataclass
class PathPack:
    """Symbol pack of n scanpaths (output of K1), all tensors on the GPU."""
    xyd: torch.Tensor      # [n, lmax, 3] f64 (x, y, duration)
    len: torch.Tensor      # [n] i32
    sym: torch.Tensor      # [n, lmax] u8
    run: torch.Tensor      # [n, lmax] i32
    nwd: torch.Tensor      # [n] i32
    sed: torch.Tensor      # [n, lmax] i32
    xyn: torch.Tensor      # [n, lmax, 2] f64

    @property
    def n(self):
        return self.xyd.shape[0]

    @property
    def lmax(self):
        return self.xyd.shape[1]

    def c_struct(self):
        return _lib.PathPack(self.sym.data_ptr(), self.run.data_ptr(), self.nwd.data_ptr(), self.sed.data_ptr(),
                             self.xyn.data_ptr(), self.len.data_ptr(), self.n, self.lmax, 0)


def pad_paths(paths, lmax=None):
    """list of [L,3] arrays -> (padded [n,lmax,3] f64 numpy, lengths i32 numpy)."""
    n = len(paths)
    lens = np.array([len(p) for p in paths], dtype=np.int32)
    lmax = int(lmax or max(1, int(lens.max()) if n else 1))
    out = np.zeros((n, lmax, 3), dtype=np.float64)
    for i, p in enumerate(paths):
        if len(p):
            out[i, :len(p)] = np.asarray(p, dtype=np.float64).reshape(-1, 3)
    return out, lens


def structured_to_xyd(fix_vector):
    """Reference structured array (start_x, start_y, duration) -> [L,3] f64."""
    if len(fix_vector) == 0:
        return np.zeros((0, 3), dtype=np.float64)
    if getattr(fix_vector, "dtype", None) is not None and fix_vector.dtype.names:
        names = fix_vector.dtype.names
        return np.stack([fix_vector[names[0]], fix_vector[names[1]], fix_vector[names[2]]], 1).astype(np.float64)
    return np.array([list(r) for r in list(fix_vector)], dtype=np.float64)


def pack_subject_lists(batch_fix_vectors, pin=True):
    """The evaluation datasets' `fix_vectors` (a list per image of per-subject structured arrays,
    dataset.py `*_evaluation.__getitem__` / `collate_func`) -> dense [N, Smax, Lmax, 3] f64 + lens
    [N, Smax] i32 (+ n_subjects [N]) in pinned host memory, ready for ScanpathPipeline.set_humans.
    Images with fewer subjects are padded with empty scanpaths (length 0)."""
    N = len(batch_fix_vectors)
    smax = max(len(f) for f in batch_fix_vectors)
    lmax = max(1, max(len(s) for f in batch_fix_vectors for s in f))
    xyd = torch.zeros((N, smax, lmax, 3), dtype=torch.float64)
    lens = torch.zeros((N, smax), dtype=torch.int32)
    nsub = torch.zeros((N,), dtype=torch.int32)
    for i, fvs in enumerate(batch_fix_vectors):
        nsub[i] = len(fvs)
        for j, fv in enumerate(fvs):
            a = structured_to_xyd(fv)
            xyd[i, j, :len(a)] = torch.from_numpy(a)
            lens[i, j] = len(a)
    if pin and torch.cuda.is_available():
        xyd, lens = xyd.pin_memory(), lens.pin_memory()
    return xyd, lens, nsub


def prep_paths(xyd: torch.Tensor, lens: torch.Tensor, cfg: ScoreConfig) -> PathPack:
    """K1: xyd [n,lmax,3] f64 + lens [n] i32 (device tensors) -> PathPack."""
    lib = _lib.load()
    assert xyd.dtype == torch.float64 and lens.dtype == torch.int32 and xyd.is_cuda and lens.is_cuda
    xyd = xyd.contiguous()
    lens = lens.contiguous()
    n, lmax = xyd.shape[0], xyd.shape[1]
    dev = xyd.device
    pack = PathPack(xyd, lens,
                    torch.empty((n, lmax), dtype=torch.uint8, device=dev),
                    torch.empty((n, lmax), dtype=torch.int32, device=dev),
                    torch.empty((n,), dtype=torch.int32, device=dev),
                    torch.empty((n, lmax), dtype=torch.int32, device=dev),
                    torch.empty((n, lmax, 2), dtype=torch.float64, device=dev))
    with torch.cuda.device(dev):
        _lib.check(lib.spb_prep_paths(_lib.ptr(xyd), _lib.ptr(lens), n, lmax, C.byref(cfg.cfg), _lib.ptr(pack.sym),
                                      _lib.ptr(pack.run), _lib.ptr(pack.nwd), _lib.ptr(pack.sed), _lib.ptr(pack.xyn),
                                      _lib.current_stream()), "spb_prep_paths")
    return pack


def pack_paths(paths, cfg: ScoreConfig, lmax=None) -> PathPack:
    """Host lists of [L,3] arrays -> device PathPack (one H2D copy + K1)."""
    arr, lens = pad_paths(paths, lmax)
    return prep_paths(torch.from_numpy(arr).to(cfg.device), torch.from_numpy(lens).to(cfg.device), cfg)


def scanmatch_matrix(A, B, cfg: "ScoreConfig"):
    """ScanMatch.match's F matrix of one pair of symbol strings (scanmatch.py:138-150) from the device kernel
    spb_scanmatch_matrix: numpy [(n+1), (m+1)] f64, bit-identical to the reference's."""
    A = np.asarray(A).astype(np.int64).reshape(-1)
    B = np.asarray(B).astype(np.int64).reshape(-1)
    nb = int(cfg.cfg.sm.Xbin) * int(cfg.cfg.sm.Ybin)
    if (len(A) and (A.min() < 0 or A.max() >= nb)) or (len(B) and (B.min() < 0 or B.max() >= nb)):
        raise IndexError("symbol outside the %d bins of the substitution matrix" % nb)      # the reference's SubMatrix[A, B]
    dev = cfg.device
    n, m = len(A), len(B)
    da = torch.from_numpy(A.astype(np.int32)).to(dev)
    db = torch.from_numpy(B.astype(np.int32)).to(dev)
    F = torch.empty((n + 1, m + 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().spb_scanmatch_matrix(_lib.ptr(da) if n else None, n, _lib.ptr(db) if m else None, m,
                                                    C.byref(cfg.cfg), _lib.ptr(F), _lib.current_stream()),
                   "spb_scanmatch_matrix")
    return F.cpu().numpy()


def tde_table(human_xy, simulated_xy, device=None):
    """[min(Lh, Ls), 3] f64 numpy table of spb_tde_distances for one pair: per window length k the 'Mean' and the
    'Hausdorff' time-delay-embedding distance (visual_attention_metrics.py:332-390) and the sum of the first k
    point distances.  Coordinates are taken as given (no rescaling)."""
    _lib.require_cuda()
    dev = torch.device(device or "cuda")
    h = torch.as_tensor(np.ascontiguousarray(np.asarray(human_xy, dtype=np.float64)[:, :2])).to(dev)
    s = torch.as_tensor(np.ascontiguousarray(np.asarray(simulated_xy, dtype=np.float64)[:, :2])).to(dev)
    Lh, Ls = h.shape[0], s.shape[0]
    kmax = min(Lh, Ls)
    if kmax == 0:
        return np.zeros((0, 3))
    lib = _lib.load()
    nbytes = lib.spb_tde_work_bytes(Lh, Ls)
    work = torch.empty((nbytes // 8,), dtype=torch.float64, device=dev)
    out = torch.empty((kmax, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spb_tde_distances(_lib.ptr(h), Lh, _lib.ptr(s), Ls, _lib.ptr(work), nbytes, _lib.ptr(out),
                                         _lib.current_stream()), "spb_tde_distances")
    return out.cpu().numpy()


class Workspace:
    """Boundary-column workspace for with-duration strings longer than 256 symbols."""

    def __init__(self, max_human_nwd: int, device):
        nbytes = _lib.load().spb_score_workspace_bytes(int(max_human_nwd))
        self.buf = torch.empty((max(nbytes, 8) // 8,), dtype=torch.float64, device=device)
        self.nbytes = nbytes


def score_pairs(human: PathPack, sim: PathPack, pair_h: torch.Tensor, pair_s: torch.Tensor, cfg: ScoreConfig,
                workspace: Workspace | None = None, out: torch.Tensor | None = None, check: bool = True):
    """K2-K4.  pair_h / pair_s: i32 device tensors [P] indexing `human` / `sim`.
    Returns scores [P,4] f64 on the device = (SM-wd, SM-wod, SED, STDE).
    `workspace=None` sizes one from the packs (costs a device->host read of the
    string lengths); pass a Workspace to stay asynchronous.  With `check`, a
    too-small workspace raises instead of leaving NaNs."""
    lib = _lib.load()
    assert pair_h.dtype == torch.int32 and pair_s.dtype == torch.int32
    P = pair_h.numel()
    dev = pair_h.device
    if out is None:
        out = torch.empty((P, 4), dtype=torch.float64, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    if workspace is None and P > 0 and sim.n > 0 and int(sim.nwd.max().item()) > 128:   # longer than one panel
        workspace = Workspace(int(human.nwd.max().item()), dev)
    hp, sp = human.c_struct(), sim.c_struct()
    with torch.cuda.device(dev):
        _lib.check(lib.spb_score_pairs(C.byref(hp), C.byref(sp), _lib.ptr(pair_h.contiguous()),
                                       _lib.ptr(pair_s.contiguous()), P, C.byref(cfg.cfg), _lib.ptr(out),
                                       _lib.ptr(workspace.buf) if workspace else None,
                                       workspace.nbytes if workspace else 0, _lib.ptr(err), _lib.current_stream()),
                   "spb_score_pairs")
    if check:
        e = int(err.item())
        if e == 2:
            raise _lib.SpbError("spb_score_pairs: a pair index lies outside its path pack (stale pair map?)")
        if e != 0:
            raise _lib.SpbError("spb_score_pairs: workspace too small for the with-duration strings")
    return out


def reduce_pairs_eval(scores: torch.Tensor, group_size: int, valid: torch.Tensor | None = None):
    """a7: scores [G*group_size, 4] -> (table [G,11] f32, reward [G] f64)."""
    lib = _lib.load()
    G = scores.shape[0] // group_size
    dev = scores.device
    out = torch.empty((G, 11), dtype=torch.float32, device=dev)
    reward = torch.empty((G,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spb_reduce_pairs_eval(_lib.ptr(scores.contiguous()),
                                             _lib.ptr(valid.contiguous()) if valid is not None else None, G,
                                             group_size, _lib.ptr(out), _lib.ptr(reward), _lib.current_stream()),
                   "spb_reduce_pairs_eval")
    return out, reward


def new_accumulator(device) -> torch.Tensor:
    """Zeroed f64 buffer for reduce_pairs(acc=...): slots [0:16] are the running sums (see the header)."""
    n = _lib.load().spb_reduce_acc_bytes() // 8
    return torch.zeros((n,), dtype=torch.float64, device=device)


def reduce_pairs(scores: torch.Tensor, group_size: int, *, n_images: int = 0, group_count: torch.Tensor | None = None,
                 pair_h: torch.Tensor | None = None, pair_s: torch.Tensor | None = None,
                 len_h: torch.Tensor | None = None, len_s: torch.Tensor | None = None, min_len_valid: int = 0,
                 valid: torch.Tensor | None = None, acc: torch.Tensor | None = None, out: torch.Tensor | None = None,
                 reward: torch.Tensor | None = None, group_valid: torch.Tensor | None = None,
                 mean_over_kept: bool = False):
    """a7 in one pass (spb_reduce_pairs): table [G,11] f32, reward [G] f64, group_valid [G] u8, and the running
    sums of `evaluation` added into `acc` (new_accumulator).  Padded subjects (index >= group_count[image]) are
    skipped; with min_len_valid the MultiMatch NaN rule is evaluated on the device from the path lengths."""
    lib = _lib.load()
    G = scores.shape[0] // group_size
    dev = scores.device
    if out is None:
        out = torch.empty((G, 11), dtype=torch.float32, device=dev)
    if reward is None:
        reward = torch.empty((G,), dtype=torch.float64, device=dev)
    if group_valid is None:
        group_valid = torch.empty((G,), dtype=torch.uint8, device=dev)
    assert out.is_contiguous() and reward.is_contiguous() and group_valid.is_contiguous() and scores.is_contiguous()
    a = _lib.ReduceArgs()
    p = lambda t: None if t is None else t.data_ptr()
    a.d_scores, a.d_valid = scores.data_ptr(), p(valid)
    a.d_pair_h, a.d_pair_s, a.d_len_h, a.d_len_s = p(pair_h), p(pair_s), p(len_h), p(len_s)
    a.d_group_count = p(group_count)
    a.d_out, a.d_reward, a.d_group_valid = out.data_ptr(), reward.data_ptr(), group_valid.data_ptr()
    a.d_acc, a.acc_bytes = p(acc), (acc.numel() * 8 if acc is not None else 0)
    a.n_groups, a.group_size, a.n_images, a.min_len_valid = G, int(group_size), int(n_images), int(min_len_valid)
    a.mean_over_kept = 1 if mean_over_kept else 0
    with torch.cuda.device(dev):
        _lib.check(lib.spb_reduce_pairs(C.byref(a), _lib.current_stream()), "spb_reduce_pairs")
    return out, reward, group_valid


def grid_pairs(n_images: int, k_samples: int, n_subjects: int, device):
    """Pair map of the evaluation drivers for sample-major predictions
    (sim index = k * n_images + image, as test.py:124-131 extends its lists) and
    image-major humans (human index = image * n_subjects + s):
    pair order = (k, image, s), i.e. `for pred: for subject` (evaluation.py:166-170)."""
    sim = torch.arange(k_samples * n_images, device=device, dtype=torch.int32)
    img = sim % n_images
    s = torch.arange(n_subjects, device=device, dtype=torch.int32)
    pair_s = sim[:, None].expand(-1, n_subjects).reshape(-1).contiguous()
    pair_h = (img[:, None] * n_subjects + s[None, :]).reshape(-1).contiguous()
    return pair_h, pair_s
