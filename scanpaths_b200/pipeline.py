"""The whole hot path as one call: decode -> sample K scanpaths per image -> score every
(sample, subject) pair -> reduce.  This is what test.py:111-149 / validation() /
the SCST reward loop (train.py:216-254) do per batch, in waves of images that stay on
the device from the feature map to the reduced score table.

Host side only: every stage is one C-ABI call (spb_decode, spb_sample_paths,
spb_prep_paths, spb_score_pairs, spb_reduce_pairs); no ATen arithmetic runs between them.
"""
from __future__ import annotations

import numpy as np
import torch

from . import scoring as S
from .models.baseline_attention import CudaDecoder
from .models.sampling import Sampling

MIN_LEN_VALID = 3    # multimatch_gaze's NaN rule, which pairs_eval's row elimination follows


class ScanpathPipeline:
    def __init__(self, state_dict, task="OSIE", steps=16, k_samples=10, min_length=1, device="cuda", wave=256,
                 seed=0, use_tensor_cores=True, overlap_tail=True):
        self.device = torch.device(device)
        self.task, self.K, self.steps, self.wave = task, int(k_samples), int(steps), int(wave)
        self.decoder = CudaDecoder(state_dict, task, steps, self.device, wave, use_tensor_cores)
        self.sampler = Sampling(convLSTM_length=steps, min_length=min_length, seed=seed)
        self.cfg = S.ScoreConfig.evaluation(device=self.device, dur_scale=1000.0)
        self.humans = None
        self.n_subjects = 0
        self.subject_count = None
        self._pairs = {}
        self._ws = None
        self._cs = None
        # sampling + scoring + reduction of wave w run on a side stream while wave w+1 decodes: those kernels use
        # neither tensor cores nor much bandwidth (latency / issue bound), so they fill the gaps of the decode
        self.overlap_tail = bool(overlap_tail)
        self._ts = None

    # ---- human scanpaths: packed once, resident on the device
    def set_humans(self, xyd, lens, n_subjects=None):
        """xyd [N, S, Lmax, 3] f64 (x, y, seconds), lens [N, S] i32, n_subjects [N] i32 or None (numpy or torch,
        host or device).  `n_subjects` is the real subject count per image when the lists were padded to a common
        S (scoring.pack_subject_lists): padded subjects are never scored into any result."""
        xyd = torch.as_tensor(xyd, dtype=torch.float64)
        lens = torch.as_tensor(lens, dtype=torch.int32)
        N, Sn, L, _ = xyd.shape
        self.n_subjects = Sn
        self._pairs.clear()                     # pair maps depend on the subject count of THIS set of humans
        self.humans = S.prep_paths(xyd.reshape(N * Sn, L, 3).to(self.device, non_blocking=True),
                                   lens.reshape(N * Sn).to(self.device, non_blocking=True), self.cfg)
        self.subject_count = None
        if n_subjects is not None:
            self.subject_count = torch.as_tensor(n_subjects, dtype=torch.int32).to(self.device, non_blocking=True)
            assert self.subject_count.numel() == N
        self._ws = S.Workspace(int(self.humans.nwd.max().item()), self.device)
        return self.humans

    def _copy_stream(self):
        if self._cs is None:
            self._cs = torch.cuda.Stream(device=self.device)
        return self._cs

    def _tail_stream(self):
        if self._ts is None:
            self._ts = torch.cuda.Stream(device=self.device)
        return self._ts

    def _pair_map(self, n, n0, sort_subjects=False):
        """Pair map of a wave: pair (k, image, s) scores sample k of the image against its subject s.
        sort_subjects: within every image the subjects are visited in order of decreasing with-duration length.
        The four pairs a warp of the scoring kernel works on are consecutive pairs, i.e. consecutive subjects of
        one sample: sorted, their DP tables have similar heights and the warp's step count (the maximum over
        its four pairs) drops by ~15 %.  Groups stay contiguous and padded subjects (length 0) stay last, so the
        reduction is unaffected; only the order of the raw score rows inside a group changes."""
        key = (n, n0, self.n_subjects, self.K, sort_subjects)
        if key not in self._pairs:
            Sn = self.n_subjects
            ph, ps = S.grid_pairs(n, self.K, Sn, self.device)
            ph = ph + n0 * Sn
            if sort_subjects:
                nwd = self.humans.nwd[n0 * Sn:(n0 + n) * Sn].view(n, Sn)
                order = torch.argsort(nwd, dim=1, descending=True, stable=True).to(torch.int32)     # [n, Sn]
                base = (torch.arange(n, device=self.device, dtype=torch.int32) * Sn + n0 * Sn).view(n, 1)
                ph = (base + order).view(1, n * Sn).expand(self.K, -1).reshape(-1)
            self._pairs[key] = (ph.contiguous(), ps)
            if len(self._pairs) > 64:
                self._pairs.pop(next(iter(self._pairs)))
        return self._pairs[key]

    def _tail(self, hd, n0, n1, N, probs, mu, s2, ph, ps, cnt, valid_min_len, table, reward, gvalid, scores_all, acc,
              paths, paths_host, stream, copy_stream):
        """sample -> prep -> score -> reduce for one head of one wave, on the current stream."""
        K, Sn, T, n = self.K, self.n_subjects, self.steps, n1 - n0
        full = n == N
        smp = self.sampler.sample_paths(probs, mu, s2, K)
        pp = S.prep_paths(smp["xyd"], smp["len"], self.cfg)
        sc = scores_all[hd].view(-1, 4) if (scores_all is not None and full) else None
        sc = S.score_pairs(self.humans, pp, ph, ps, self.cfg, workspace=self._ws, check=False, out=sc)
        # the reduced tables of a wave are dense [K, n, ...] blocks; written in place when one wave covers the
        # batch, else copied into the [K, N, ...] results
        tab_w = table[hd].view(K * N, 11) if full else None
        rew_w = reward[hd].view(K * N) if full else None
        gv_w = gvalid[hd].view(K * N) if full else None
        tab_w, rew_w, gv_w = S.reduce_pairs(sc, Sn, n_images=n, group_count=cnt, pair_h=ph, pair_s=ps,
                                            len_h=self.humans.len, len_s=pp.len, min_len_valid=valid_min_len,
                                            acc=acc, out=tab_w, reward=rew_w, group_valid=gv_w)
        if not full:
            table[hd, :, n0:n1] = tab_w.view(K, n, 11)
            reward[hd, :, n0:n1] = rew_w.view(K, n)
            gvalid[hd, :, n0:n1] = gv_w.view(K, n)
            if scores_all is not None:
                scores_all[hd, :, n0:n1] = sc.view(K, n, Sn, 4)
        if paths is not None:
            paths.append((hd, n0, n1, smp))
        if paths_host is not None:
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                paths_host[0][hd, :, n0:n1].copy_(smp["xyd"].view(K, n, T, 3), non_blocking=True)
                paths_host[1][hd, :, n0:n1].copy_(smp["len"].view(K, n), non_blocking=True)
            smp["xyd"].record_stream(copy_stream)
            smp["len"].record_stream(copy_stream)

    def run(self, visual_feature, attention_maps=None, tasks=None, keep_scores=False, valid_min_len=MIN_LEN_VALID,
            keep_paths=False, paths_host=None):
        """visual_feature [N,512,30,40] f32 on the device or in (pinned) host memory.
        Returns dict: table [HD, K, N, 11] f32 (pairs_eval layout per sample; rows follow the reference's
        elimination rule with `valid_min_len`), reward [HD, K, N] f64, group_valid [HD, K, N] u8 (0 where
        train.py:237 would reject the trial), acc (the running sums behind `evaluation`'s aggregate over all
        real pairs) and optionally the raw scores / the sampled scanpaths.
        paths_host: optional (xyd [HD, K, N, T, 3] f64, len [HD, K, N] i32) pinned host tensors; every wave's
        sampled scanpaths are copied into them on the side stream (what test.py:135-152 dumps)."""
        assert self.humans is not None, "set_humans() first"
        N = visual_feature.shape[0]
        assert self.humans.n == N * self.n_subjects, "set_humans() was called for %d images, run() got %d" % (
            self.humans.n // max(self.n_subjects, 1), N)
        dev, K, Sn, HD, T = self.device, self.K, self.n_subjects, self.decoder.heads, self.steps
        table = torch.empty((HD, K, N, 11), dtype=torch.float32, device=dev)
        reward = torch.empty((HD, K, N), dtype=torch.float64, device=dev)
        gvalid = torch.empty((HD, K, N), dtype=torch.uint8, device=dev)
        scores_all = torch.empty((HD, K, N, Sn, 4), dtype=torch.float64, device=dev) if keep_scores else None
        accs = [S.new_accumulator(dev) for _ in range(HD)]
        paths = [] if keep_paths else None
        host_in = not visual_feature.is_cuda
        # host input: the next wave's features are copied on a side stream while this wave computes
        copy_stream = self._copy_stream() if (host_in or paths_host is not None) else None
        main = torch.cuda.current_stream(dev)
        starts = list(range(0, N, self.wave))
        full = len(starts) == 1
        tail = self._tail_stream() if (self.overlap_tail and not full) else main

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
            ph, ps = self._pair_map(n, n0, sort_subjects=not keep_scores)     # raw rows are returned in (k, image, s) order
            cnt = None if self.subject_count is None else self.subject_count[n0:n1]
            if tail is not main:
                decoded = torch.cuda.Event()
                decoded.record(main)
                tail.wait_event(decoded)
                for t_ in (probs, mu, s2):
                    t_.record_stream(tail)
            with torch.cuda.stream(tail):
                for hd in range(HD):
                    self._tail(hd, n0, n1, N, probs[hd], mu[hd], s2[hd], ph, ps, cnt, valid_min_len, table, reward,
                               gvalid, scores_all, accs[hd], paths, paths_host, tail, copy_stream)
        if tail is not main:
            main.wait_stream(tail)
        if paths_host is not None:
            main.wait_stream(copy_stream)
        acc = torch.stack([a[:16] for a in accs], 0)
        out = {"table": table, "reward": reward, "group_valid": gvalid, "acc": acc}
        if keep_scores:
            out["scores"] = scores_all
        if keep_paths:
            out["paths"] = paths
        return out

    @staticmethod
    def metrics(out, head=0):
        """Host-side view of the accumulators in the shape of `evaluation`'s return (one D2H read)."""
        a = out["acc"][head].cpu().numpy()
        P, G = max(a[12], 1.0), max(a[13], 1.0)
        mean = a[0:4] / P
        std = np.sqrt(np.maximum(a[4:8] / P - mean ** 2, 0.0))
        bmean = a[8:10] / G
        bstd = np.sqrt(np.maximum(a[10:12] / G - bmean ** 2, 0.0))
        m = {"ScanMatch": {"w/o duration": mean[1], "with duration": mean[0]},
             "VAME": {"SED": mean[2], "STDE": mean[3], "SED_best": bmean[0], "STDE_best": bmean[1]}}
        s = {"ScanMatch": {"w/o duration": std[1], "with duration": std[0]},
             "VAME": {"SED": std[2], "STDE": std[3], "SED_best": bstd[0], "STDE_best": bstd[1]}}
        return m, s
