"""The whole hot path as one call: decode -> sample K scanpaths per image -> score every
(sample, subject) pair -> reduce.  This is what test.py:111-149 / validation() /
the SCST reward loop (train.py:216-254) do per batch, in waves of images that stay on
the device from the feature map to the reduced score table.

Host side only: every stage is one C-ABI call (spb_decode, spb_sample_paths,
spb_prep_paths, spb_score_pairs, spb_reduce_pairs_eval).
"""
from __future__ import annotations

import numpy as np
import torch

from . import scoring as S
from .models.baseline_attention import CudaDecoder
from .models.sampling import Sampling


class ScanpathPipeline:
    def __init__(self, state_dict, task="OSIE", steps=16, k_samples=10, min_length=1, device="cuda", wave=256,
                 seed=0, use_tensor_cores=True):
        self.device = torch.device(device)
        self.task, self.K, self.steps, self.wave = task, int(k_samples), int(steps), int(wave)
        self.decoder = CudaDecoder(state_dict, task, steps, self.device, wave, use_tensor_cores)
        self.sampler = Sampling(convLSTM_length=steps, min_length=min_length, seed=seed)
        self.cfg = S.ScoreConfig.evaluation(device=self.device, dur_scale=1000.0)
        self.humans = None
        self.n_subjects = 0
        self._pairs = {}
        self._ws = None

    # ---- human scanpaths: packed once, resident on the device
    def set_humans(self, xyd, lens):
        """xyd [N, S, Lmax, 3] f64 (x, y, seconds), lens [N, S] i32 (numpy or torch, host or device)."""
        xyd = torch.as_tensor(xyd, dtype=torch.float64)
        lens = torch.as_tensor(lens, dtype=torch.int32)
        N, Sn, L, _ = xyd.shape
        self.n_subjects = Sn
        self.humans = S.prep_paths(xyd.reshape(N * Sn, L, 3).to(self.device, non_blocking=True),
                                   lens.reshape(N * Sn).to(self.device, non_blocking=True), self.cfg)
        self._ws = S.Workspace(int(self.humans.nwd.max().item()), self.device)
        return self.humans

    def _copy_stream(self):
        if getattr(self, "_cs", None) is None:
            self._cs = torch.cuda.Stream(device=self.device)
        return self._cs

    def _pair_map(self, n, n0):
        key = (n, n0)
        if key not in self._pairs:
            ph, ps = S.grid_pairs(n, self.K, self.n_subjects, self.device)
            self._pairs[key] = ((ph + n0 * self.n_subjects).contiguous(), ps)
            if len(self._pairs) > 64:
                self._pairs.pop(next(iter(self._pairs)))
        return self._pairs[key]

    def run(self, visual_feature, attention_maps=None, tasks=None, keep_scores=False, valid_min_len=0,
            keep_paths=False):
        """visual_feature [N,512,30,40] f32 on the device or in (pinned) host memory.
        Returns dict: table [HD, K, N, 11] f32 (pairs_eval layout per sample), reward [HD, K, N] f64,
        metrics (the `evaluation` aggregate over all pairs) and optionally the raw scores."""
        assert self.humans is not None, "set_humans() first"
        N = visual_feature.shape[0]
        dev, K, Sn, HD = self.device, self.K, self.n_subjects, self.decoder.heads
        table = torch.empty((HD, K, N, 11), dtype=torch.float32, device=dev)
        reward = torch.empty((HD, K, N), dtype=torch.float64, device=dev)
        scores_all = torch.empty((HD, K, N, Sn, 4), dtype=torch.float64, device=dev) if keep_scores else None
        acc = torch.zeros((HD, 12), dtype=torch.float64, device=dev)   # sum, sumsq (4 each), best sums/sumsq (2+2)
        paths = [] if keep_paths else None
        host_in = not visual_feature.is_cuda
        # host input: the next wave's features are copied on a side stream while this wave computes
        copy_stream = self._copy_stream() if host_in else None
        main = torch.cuda.current_stream(dev)
        starts = list(range(0, N, self.wave))

        def fetch(n0):
            n1 = min(N, n0 + self.wave)
            with torch.cuda.stream(copy_stream):
                t = visual_feature[n0:n1].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return t, ev

        pending = fetch(starts[0]) if host_in else None
        for wi, n0 in enumerate(starts):
            n1 = min(N, n0 + self.wave)
            n = n1 - n0
            if host_in:
                vf, ev = pending
                main.wait_event(ev)
                vf.record_stream(main)
                pending = fetch(starts[wi + 1]) if wi + 1 < len(starts) else None
            else:
                vf = visual_feature[n0:n1]
            att = None if attention_maps is None else attention_maps[n0:n1].to(dev, non_blocking=True)
            tk = None if tasks is None else tasks[n0:n1]
            probs, mu, s2, _ = self.decoder.decode(vf, att, tk)
            ph, ps = self._pair_map(n, n0)
            for hd in range(HD):
                smp = self.sampler.sample_paths(probs[hd], mu[hd], s2[hd], K)
                pp = S.prep_paths(smp["xyd"], smp["len"], self.cfg)
                sc = S.score_pairs(self.humans, pp, ph, ps, self.cfg, workspace=self._ws, check=False)
                valid = None
                if valid_min_len > 0:
                    valid = ((self.humans.len[ph.long()] >= valid_min_len) &
                             (pp.len[ps.long()] >= valid_min_len)).to(torch.uint8)
                tab, rew = S.reduce_pairs_eval(sc, Sn, valid)
                table[hd, :, n0:n1] = tab.view(K, n, 11)
                reward[hd, :, n0:n1] = rew.view(K, n)
                g = sc.view(K * n, Sn, 4)
                sed_best, stde_best = g[:, :, 2].min(1)[0], g[:, :, 3].max(1)[0]
                acc[hd, 0:4] += sc.sum(0)
                acc[hd, 4:8] += (sc * sc).sum(0)
                acc[hd, 8] += sed_best.sum(); acc[hd, 9] += stde_best.sum()
                acc[hd, 10] += (sed_best * sed_best).sum(); acc[hd, 11] += (stde_best * stde_best).sum()
                if keep_scores:
                    scores_all[hd, :, n0:n1] = sc.view(K, n, Sn, 4)
                if keep_paths:
                    paths.append((hd, n0, n1, smp))
        out = {"table": table, "reward": reward, "acc": acc, "n_pairs": K * N * Sn, "n_groups": K * N}
        if keep_scores:
            out["scores"] = scores_all
        if keep_paths:
            out["paths"] = paths
        return out

    @staticmethod
    def metrics(out, head=0):
        """Host-side view of the accumulators in the shape of `evaluation`'s return (one D2H read)."""
        a = out["acc"][head].cpu().numpy()
        P, G = out["n_pairs"], out["n_groups"]
        mean = a[0:4] / P
        std = np.sqrt(np.maximum(a[4:8] / P - mean ** 2, 0.0))
        bmean = a[8:10] / G
        bstd = np.sqrt(np.maximum(a[10:12] / G - bmean ** 2, 0.0))
        m = {"ScanMatch": {"w/o duration": mean[1], "with duration": mean[0]},
             "VAME": {"SED": mean[2], "STDE": mean[3], "SED_best": bmean[0], "STDE_best": bmean[1]}}
        s = {"ScanMatch": {"w/o duration": std[1], "with duration": std[0]},
             "VAME": {"SED": std[2], "STDE": std[3], "SED_best": bstd[0], "STDE_best": bstd[1]}}
        return m, s
