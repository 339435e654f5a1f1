"""Parity of the CUDA sampler (csrc/sample.cu, through the C ABI) with the
goldens recorded from the reference under injected draws, and with the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Sampling():
    from scanpaths_b200 import build
    build.build_library()
    from scanpaths_b200.models.sampling import Sampling
    return Sampling


@pytest.mark.parametrize("min_len", [1, 2])
def test_injected_draws_reproduce_reference(Sampling, golden_dir, min_len):
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    probs, mu, s2 = t(g["probs"]), t(g["mu"]), t(g["sigma2"])
    sampler = Sampling(convLSTM_length=16, min_length=min_len)
    # all three trials in one batched call (K = 3)
    q = torch.stack([t(g["m%d_t%d_q" % (min_len, k)]) for k in range(3)])
    z = torch.stack([t(g["m%d_t%d_z" % (min_len, k)]) for k in range(3)])
    out = sampler.sample_paths(probs, mu, s2, K=3, q=q, z=z)
    for k in range(3):
        tag = "m%d_t%d_" % (min_len, k)
        assert np.array_equal(out["selected_actions"][k].cpu().numpy(), g[tag + "actions"])       # exact
        assert np.array_equal(out["selected_actions_probs"][k].cpu().numpy(), g[tag + "sel_prob"])
        np.testing.assert_allclose(out["durations"][k].cpu().numpy(), g[tag + "dur"], rtol=1e-6)
        assert np.array_equal(out["scanpath_length"][k].cpu().numpy(), g[tag + "length"].reshape(-1))
        assert np.array_equal(out["action_masks"][k].cpu().numpy(), g[tag + "action_mask"])
        assert np.array_equal(out["duration_masks"][k].cpu().numpy(), g[tag + "duration_mask"])
        N = probs.shape[0]
        lens = out["len"][k * N:(k + 1) * N].cpu().numpy()
        assert np.array_equal(lens, g[tag + "fix_len"])
        xyd = out["xyd"][k * N:(k + 1) * N].cpu().numpy()
        for n in range(N):
            assert np.array_equal(xyd[n, :lens[n], :2], g[tag + "fix"][n, :lens[n], :2])
            np.testing.assert_allclose(xyd[n, :lens[n], 2], g[tag + "fix"][n, :lens[n], 2], rtol=1e-6)
        # reference-style API, one trial
        r = sampler.random_sample(probs, mu, s2, q=q[k], z=z[k])
        assert r["selected_actions"].dtype == torch.int64
        assert np.array_equal(r["selected_actions"].cpu().numpy(), g[tag + "actions"])
        assert r["scanpath_length"].shape == (N, 1)
        fix, am, dm = sampler.generate_scanpath(torch.zeros(N, 3, 4, 4, device=dev), r["selected_actions_probs"],
                                                r["durations"], r["selected_actions"])
        assert np.array_equal(am.cpu().numpy(), g[tag + "action_mask"])
        assert [len(f) for f in fix] == list(g[tag + "fix_len"])
        assert fix[0].dtype.names == ("start_x", "start_y", "duration")
        # log-likelihoods of the samples (loss.py:34-45)
        from scanpaths_b200.models.loss import LogAction, LogDuration
        la = LogAction(r["selected_actions_probs"], am)
        ld = LogDuration(r["durations"], mu, s2, dm)
        np.testing.assert_allclose(la.cpu().numpy(), g[tag + "log_action"], rtol=1e-5)
        np.testing.assert_allclose(ld.cpu().numpy(), g[tag + "log_duration"], rtol=1e-5)


def test_philox_sampler_statistics(Sampling):
    """Without injection: the empirical action distribution matches the masked,
    renormalised probabilities; durations follow exp(z*sigma2+mu); streams differ per call."""
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(5)
    N, T, A, K = 2, 16, 1201, 4096
    logits = torch.randn(N, T, A, generator=gen, device=dev)
    logits[:, :, :8] += 5.0
    probs = torch.softmax(logits, -1)
    mu = torch.full((N, T), -1.4, device=dev); s2 = torch.full((N, T), 0.15, device=dev)
    sampler = Sampling(convLSTM_length=T, min_length=2, seed=1234)
    out = sampler.sample_paths(probs, mu, s2, K=K)
    acts = out["selected_actions"]
    assert int((acts[:, :, :2] == 0).sum()) == 0                      # min_length masking
    p = probs[0, 5].double()
    freq = torch.bincount(acts[:, 0, 5].long(), minlength=A).double() / K
    top = torch.topk(p, 8).indices
    assert torch.allclose(freq[top], p[top], atol=4 * float((p[top] * (1 - p[top]) / K).sqrt().max()) + 1e-3)
    logd = out["durations"].double().log()
    assert abs(float(logd.mean()) + 1.4) < 0.01 and abs(float(logd.std()) - 0.15) < 0.01
    out2 = sampler.sample_paths(probs, mu, s2, K=K)
    assert not torch.equal(out2["selected_actions"], acts)
    again = Sampling(convLSTM_length=T, min_length=2, seed=1234).sample_paths(probs, mu, s2, K=K)
    assert torch.equal(again["selected_actions"], acts)                # reproducible from the seed


def test_sampled_paths_score_like_oracle(Sampling):
    """Sampler output -> scoring kernels, checked end to end against the C oracle."""
    from oracle import c_scoring as CO
    from scanpaths_b200 import scoring as S
    from golden.make_goldens import human_paths
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(9)
    N, T, A, K, Sn = 8, 16, 1201, 10, 5
    logits = torch.randn(N, T, A, generator=gen, device=dev); logits[:, :, 0] += 4.5
    probs = torch.softmax(logits, -1)
    mu = torch.full((N, T), -1.4, device=dev); s2 = torch.full((N, T), 0.15, device=dev)
    out = Sampling(convLSTM_length=T, min_length=1, seed=3).sample_paths(probs, mu, s2, K=K)
    cfg = S.ScoreConfig.evaluation()                                   # seconds in, x1000 inside
    pp = S.prep_paths(out["xyd"], out["len"], cfg)
    rng = np.random.default_rng(4)
    H = human_paths(rng, N * Sn)
    hp = S.pack_paths(H, cfg)
    ph, ps = S.grid_pairs(N, K, Sn, dev)
    got = S.score_pairs(hp, pp, ph, ps, cfg).cpu().numpy()
    ha, hl = S.pad_paths([h * [1, 1, 1000.0] for h in H])
    pa = out["xyd"].cpu().numpy().copy(); pa[..., 2] *= 1000.0
    ref = CO.score_pairs(ha, hl, pa, out["len"].cpu().numpy(), ph.cpu().numpy(), ps.cpu().numpy())
    assert np.array_equal(got[:, :3], ref[:, :3], equal_nan=True)
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=1e-12)
