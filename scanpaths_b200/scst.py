"""The self-critical (SCST) reward / loss tail of the training loops on the device
(OSIE/train.py:223-258; COCO_Search18/train.py:255-287; AiR/train.py:253-342).

Per batch the reference runs, K = rl_sample_number times on the host: random_sample ->
generate_scanpath (2N device->host syncs) -> pairs_eval (N*S pure-Python pair scorings) -> reject the
trial on NaN -> LogAction / LogDuration, then scipy hmean, the mean-over-trials baseline and the loss.
Here the same batch is: ONE sampling launch for all trials (a few spare ones for the rejection rule),
ONE prep + ONE score launch for all K*N*S pairs, ONE reduction (reward + per-image validity), and the
fused loss (2 launches forward, 2 backward) -- nothing leaves the device until the caller reads the loss.

``ScstLoss`` is a ``torch.autograd.Function``: gradients flow to all_actions_prob, log_normal_mu and
log_normal_sigma2, i.e. into the PyTorch forward-with-grad of the model (which stays in PyTorch).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import scoring as S

MIN_LEN_VALID = 3


def _scst_args(probs, mu, s2, actions, dur, am, dm, reward, group_valid, extra_adv, bufs, k_use):
    K, N, T = actions.shape
    a = _lib.ScstArgs()
    p = lambda t: None if t is None else t.data_ptr()
    a.d_probs, a.d_mu, a.d_sigma2 = p(probs), p(mu), p(s2)
    a.d_actions, a.d_dur, a.d_action_mask, a.d_duration_mask = p(actions), p(dur), p(am), p(dm)
    a.d_reward, a.d_group_valid, a.d_extra_adv = p(reward), p(group_valid), p(extra_adv)
    a.d_loss, a.d_adv, a.d_log_actions, a.d_log_durations = p(bufs["loss"]), p(bufs["adv"]), p(bufs["la"]), p(bufs["ld"])
    a.d_mask_sums, a.d_trial_used = p(bufs["msum"]), p(bufs["used"])
    a.N, a.T, a.A, a.K, a.k_use = N, T, probs.shape[-1], K, int(k_use)
    return a


class ScstLoss(torch.autograd.Function):
    """loss = sum_{k used, n} (-LogAction[k,n] - LogDuration[k,n]) * (reward[k,n] - mean_k reward[.,n])
    (train.py:242-258) for K sampled trials, of which the first `k_use` accepted ones count.
    Returns (loss, aux) with aux = dict(loss_actions, loss_duration, advantage [K,N], trial_used [K],
    n_used, neg_log_actions [K,N], neg_log_durations [K,N]) -- all device tensors."""

    @staticmethod
    def forward(ctx, all_actions_prob, log_normal_mu, log_normal_sigma2, actions, durations, action_masks,
                duration_masks, reward, group_valid, k_use, extra_adv):
        lib = _lib.load()
        if not all_actions_prob.is_cuda:
            raise _lib.SpbError("scanpaths_b200 has no CPU path: tensors must be on the GPU")
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        probs, mu, s2 = f32(all_actions_prob), f32(log_normal_mu), f32(log_normal_sigma2)
        acts = actions.detach().to(torch.int32).contiguous()
        dur, am, dm = f32(durations), f32(action_masks), f32(duration_masks)
        rew = reward.detach().to(torch.float64).contiguous()
        gv = None if group_valid is None else group_valid.detach().to(torch.uint8).contiguous()
        ex = None if extra_adv is None else f32(extra_adv)
        K, N, T = acts.shape
        dev = probs.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        bufs = {"loss": f(3), "adv": f(K, N), "la": f(K, N), "ld": f(K, N), "msum": f(K, 2),
                "used": torch.empty((K + 1,), dtype=torch.int32, device=dev)}
        args = _scst_args(probs, mu, s2, acts, dur, am, dm, rew, gv, ex, bufs, k_use or K)
        with torch.cuda.device(dev):
            _lib.check(lib.spb_scst_loss(C.byref(args), _lib.current_stream()), "spb_scst_loss")
        ctx.keep = (probs, mu, s2, acts, dur, am, dm, rew, gv, ex, bufs, k_use or K)
        ctx.mark_non_differentiable(bufs["adv"], bufs["used"], bufs["la"], bufs["ld"])
        return bufs["loss"], bufs["adv"], bufs["used"], bufs["la"], bufs["ld"]

    @staticmethod
    def backward(ctx, g_loss3, *_):
        lib = _lib.load()
        probs, mu, s2, acts, dur, am, dm, rew, gv, ex, bufs, k_use = ctx.keep
        g = g_loss3.detach().to(torch.float32).reshape(3).contiguous()     # on (loss, loss_actions, loss_duration)
        gp, gm, gs = torch.empty_like(probs), torch.empty_like(mu), torch.empty_like(s2)
        args = _scst_args(probs, mu, s2, acts, dur, am, dm, rew, gv, ex, bufs, k_use)
        with torch.cuda.device(probs.device):
            _lib.check(lib.spb_scst_loss_backward(C.byref(args), _lib.ptr(g), _lib.ptr(gp), _lib.ptr(gm), _lib.ptr(gs),
                                                  _lib.current_stream()), "spb_scst_loss_backward")
        return (gp, gm, gs) + (None,) * 8


def scst_loss(all_actions_prob, log_normal_mu, log_normal_sigma2, samples, reward, group_valid=None, k_use=None,
              extra_adv=None):
    """samples: the dict of Sampling.sample_paths ([K,N,T] tensors).  Returns (loss, aux dict)."""
    out = ScstLoss.apply(all_actions_prob, log_normal_mu, log_normal_sigma2, samples["selected_actions"],
                         samples["durations"], samples["action_masks"], samples["duration_masks"], reward, group_valid,
                         k_use, extra_adv)
    loss3, adv, used, la, ld = out
    K = adv.shape[0]
    return loss3[0], {"loss_actions": loss3[1], "loss_duration": loss3[2], "advantage": adv, "trial_used": used[:K],
                      "n_used": used[K], "neg_log_actions": -la, "neg_log_durations": -ld}


class ScstRewardStep:
    """One SCST batch on the device: sample -> score -> reward -> fused loss.

        step = ScstRewardStep(sampler, device, rl_sample_number=5)
        step.set_humans(batch["fix_vectors"])            # list per image of per-subject structured arrays
        loss, aux = step(predict["all_actions_prob"], predict["log_normal_mu"], predict["log_normal_sigma2"])
        loss.backward()

    `spare` extra trials are sampled in the same launch so that the rejection rule (train.py:237-238: a trial
    whose pairs_eval table has a NaN is re-drawn) normally needs no second round; if fewer than K trials are
    accepted, another round is sampled -- like the reference's `while True` loop, on device-side flags."""

    def __init__(self, sampler, device, rl_sample_number=5, spare=3, max_rounds=50, scanmatch_only=False):
        self.sampler, self.device = sampler, torch.device(device)
        self.K, self.spare, self.max_rounds = int(rl_sample_number), int(spare), int(max_rounds)
        self.cfg = S.ScoreConfig.evaluation(device=self.device, dur_scale=1000.0)
        # COCO-Search18's pairs_eval_scanmatch has no MultiMatch call, hence no < 3 fixations rule
        self.min_len_valid = 0 if scanmatch_only else MIN_LEN_VALID
        self.humans = self.count = None
        self.n_subjects = 0
        self._pairs = {}
        self._ws = None

    def set_humans(self, fix_vectors=None, packed=None):
        xyd, lens, nsub = packed if packed is not None else S.pack_subject_lists(fix_vectors)
        N, Sn, L, _ = xyd.shape
        self.n_subjects = Sn
        self.humans = S.prep_paths(xyd.reshape(N * Sn, L, 3).to(self.device, non_blocking=True),
                                   lens.reshape(N * Sn).to(self.device, non_blocking=True), self.cfg)
        self.count = nsub.to(self.device, non_blocking=True)
        self._ws = S.Workspace(int(self.humans.nwd.max().item()), self.device)

    def _pair_map(self, n, k):
        key = (n, k, self.n_subjects)
        if key not in self._pairs:
            self._pairs[key] = S.grid_pairs(n, k, self.n_subjects, self.device)
        return self._pairs[key]

    def sample_and_score(self, probs, mu, s2, k_total):
        N = probs.shape[0]
        with torch.no_grad():
            smp = self.sampler.sample_paths(probs, mu, s2, k_total)
            pp = S.prep_paths(smp["xyd"], smp["len"], self.cfg)
            ph, ps = self._pair_map(N, k_total)
            sc = S.score_pairs(self.humans, pp, ph, ps, self.cfg, workspace=self._ws, check=False)
            table, reward, gvalid = S.reduce_pairs(sc, self.n_subjects, n_images=N, group_count=self.count, pair_h=ph,
                                                   pair_s=ps, len_h=self.humans.len, len_s=pp.len,
                                                   min_len_valid=self.min_len_valid)
        return smp, table.view(k_total, N, 11), reward.view(k_total, N), gvalid.view(k_total, N)

    def __call__(self, all_actions_prob, log_normal_mu, log_normal_sigma2, extra_adv=None):
        assert self.humans is not None, "set_humans() first"
        k_total = self.K + self.spare
        for _ in range(self.max_rounds):
            smp, table, reward, gvalid = self.sample_and_score(all_actions_prob, log_normal_mu, log_normal_sigma2,
                                                               k_total)
            loss, aux = scst_loss(all_actions_prob, log_normal_mu, log_normal_sigma2, smp, reward, gvalid, self.K,
                                  extra_adv)
            aux["table"], aux["reward"], aux["samples"] = table, reward, smp
            if self.spare == 0 and self.min_len_valid == 0:
                return loss, aux                       # nothing can be rejected: stay asynchronous
            if int(aux["n_used"].item()) >= self.K:    # the one host read of the step (the loss is read anyway)
                return loss, aux
        raise _lib.SpbError("SCST: fewer than %d accepted trials after %d rounds (every sampled scanpath shorter "
                            "than 3 fixations?)" % (self.K, self.max_rounds))


class AirScstStep:
    """The AiR SCST batch (AiR/train.py:219-342) on the device: for the correct-answer head (`good_*`, scored
    against the subjects who answered correctly) and the incorrect-answer head (`poor_*`), K trials each:

        same[k,n] = hmean over (SM w/o, SM with duration) of the mean over the subjects whose performance equals
                    the head's, diff[k,n] likewise over the other subjects     (pairs_eval_scanmatch_performance_
                    related, AiR/utils/evaluation.py:361-420; groups without a subject count 0, train.py:283-284)
        loss      = sum over both heads of ScstLoss(reward = same, baseline = mean over the head's K trials)

    That is the loss the reference ACTUALLY optimises: in AiR/train.py:330-338 the `+ args.lambda_5 * (...)`
    lines are separate expression statements, so the Consistency-Divergence term never reaches `loss`.  With
    `lambda_5 != 0` this class adds the term as it was evidently meant (through ScstLoss's `extra_adv`):
        difference_reward = |(same - diff) - (gt_same - gt_diff)| * usable        (train.py:322-327)
    with the human-vs-human scores of gtpairs_eval_scanmatch_performance_related (:227-229, :311-321)."""

    def __init__(self, sampler, device, rl_sample_number=5, lambda_5=0.0):
        self.sampler, self.device, self.K, self.lambda_5 = sampler, torch.device(device), int(rl_sample_number), float(lambda_5)
        self.cfg = S.ScoreConfig.evaluation(device=self.device, dur_scale=1000.0)
        self.humans = None

    def set_humans(self, fix_vectors, performances):
        xyd, lens, nsub = S.pack_subject_lists(fix_vectors)
        N, Sn, L, _ = xyd.shape
        self.N, self.Sn = N, Sn
        self.humans = S.prep_paths(xyd.reshape(N * Sn, L, 3).to(self.device, non_blocking=True),
                                   lens.reshape(N * Sn).to(self.device, non_blocking=True), self.cfg)
        perf = torch.zeros((N, Sn), dtype=torch.bool)
        real = torch.zeros((N, Sn), dtype=torch.bool)
        for i, p in enumerate(performances):
            for j, v in enumerate(p):
                perf[i, j] = bool(v == True)        # noqa: E712 (the reference compares with ==)
                real[i, j] = True
        self.perf, self.real = perf.to(self.device), real.to(self.device)
        self._ws = S.Workspace(int(self.humans.nwd.max().item()), self.device)
        self._pairs = S.grid_pairs(N, self.K, Sn, self.device)
        self.gt_scores = None
        if self.lambda_5 != 0.0:
            from .utils.evaluation import gtpairs_eval_scanmatch_performance_related
            g, p, d = gtpairs_eval_scanmatch_performance_related(fix_vectors, None, None, performances)
            hm = lambda a: _hmean2(torch.nan_to_num(torch.from_numpy(a).to(self.device)))
            self.gt_scores = (hm(g), hm(p), hm(d))             # good-good, poor-poor, good-poor  [N]

    def _head(self, probs, mu, s2, given):
        K, N, Sn = self.K, self.N, self.Sn
        with torch.no_grad():
            smp = self.sampler.sample_paths(probs, mu, s2, K)
            pp = S.prep_paths(smp["xyd"], smp["len"], self.cfg)
            ph, ps = self._pairs
            sc = S.score_pairs(self.humans, pp, ph, ps, self.cfg, workspace=self._ws, check=False)
            same_m = (self.real & (self.perf == given)).to(torch.uint8)
            diff_m = (self.real & (self.perf != given)).to(torch.uint8)
            out = []
            for m in (same_m, diff_m):
                valid = m.unsqueeze(0).expand(K, N, Sn).contiguous().view(-1)
                tab, rew, _ = S.reduce_pairs(sc, Sn, n_images=N, valid=valid, mean_over_kept=True)
                out.append((tab.view(K, N, 11)[..., 5:7], torch.nan_to_num(rew.view(K, N))))   # NaN (empty group) -> 0
        return smp, out[0], out[1]

    def __call__(self, predict):
        """predict: the AiR model's output dict (good_* / poor_* tensors).  Returns (loss, aux)."""
        assert self.humans is not None, "set_humans() first"
        total, aux = None, {}
        for name, given in (("good", True), ("poor", False)):
            probs, mu, s2 = (predict[name + "_all_actions_prob"], predict[name + "_log_normal_mu"],
                             predict[name + "_log_normal_sigma2"])
            smp, (same_tab, same), (diff_tab, diff) = self._head(probs, mu, s2, given)
            extra = None
            if self.lambda_5 != 0.0:
                gt_same = self.gt_scores[0 if given else 1].unsqueeze(0)
                gt_diff = self.gt_scores[2].unsqueeze(0)
                usable = ((gt_same != 0) & (gt_diff != 0)).to(torch.float64)
                dr = ((same - diff) - (gt_same - gt_diff)).abs() * usable
                extra = (self.lambda_5 * (dr - dr.mean(0, keepdim=True))).float()
            loss, a = scst_loss(probs, mu, s2, smp, same, None, self.K, extra)
            total = loss if total is None else total + loss
            aux[name] = dict(a, same=same, diff=diff, same_table=same_tab, diff_table=diff_tab, samples=smp)
        return total, aux


def _hmean2(t):
    """scipy.stats.hmean over the last axis of [..., 2] non-negative values (0 if either is 0)."""
    a, b = t[..., 0].double(), t[..., 1].double()
    return torch.where((a > 0) & (b > 0), 2.0 / (1.0 / a + 1.0 / b), torch.zeros_like(a))
