#!/usr/bin/env python
"""bench.py -- scored scanpaths/sec of the decode + sample + ScanMatch/SED/STDE hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--task OSIE|AiR|COCO_Search18]
                  [--scaling weak|strong] [--no-extras]

One "step" = one pass of the hot path over the OSIE-shaped workload of BASELINE.json configs[1]:
4096 images x 64 sampled scanpaths x 15 human subjects per GPU (weak scaling: every rank owns its own
4096-image shard; `--scaling strong`: ONE 4096-image job sharded over the ranks); the reduced score tables
are exchanged with one NCCL all-gather.  Prints ONE JSON line (rank 0).  At N = 1 the line also carries
`extra`: the AiR-shaped workload (north_star's target configuration: two streams, two heads, 128 samples per
image) with its own CPU baseline, roofline and e2e, the COCO-Search18-shaped workload, the SCST reward step
(N = 4 images, K = 5 trials, 15 subjects: a latency regime) and `human_evaluation` (S x (S-1) ordered pairs per
image: scoring only).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scored scanpaths/sec (decode+ScanMatch/SED/STDE)"
UNIT = "scanpaths/s"
T_STEPS, A = 16, 1201
TAGS = {1: "conv3x3_x", 2: "winograd_gemm_h", 3: "conv5x5", 4: "lstm_cell", 5: "head", 6: "feedback",
        7: "rank1", 8: "prep", 9: "wino_input", 10: "score_pairs", 11: "sample"}

_STDOUT_FD = None


def _quiet_stdout():
    """Everything libraries print to stdout during the run (NCCL's version banner, ...) goes to stderr, so that
    the JSON line is the only thing on stdout."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    sys.stdout.write(json.dumps(line) + "\n")
    sys.stdout.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=4096, help="images per GPU (weak) or in total (strong)")
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--subjects", type=int, default=15)
    ap.add_argument("--wave", type=int, default=256)
    ap.add_argument("--task", default="OSIE", choices=["OSIE", "AiR", "COCO_Search18"],
                    help="model variant of the workload (default: the OSIE-shaped configs[1] the metric is quoted on)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the AiR / COCO / SCST / human_evaluation records")
    ap.add_argument("--cpu-images", type=int, default=16,
                    help="images in the bounded CPU-baseline sample (about 10-15 s on 16 cores)")
    return ap.parse_args()


def synth_humans(n_images, n_subjects, seed, lo=6, hi=14):
    """Lh ~ U{6..14}, x ~ U(0,320), y ~ U(0,240), dur = exp(N(log 0.25, 0.4)) s."""
    rng = np.random.default_rng(seed)
    L = rng.integers(lo, hi + 1, (n_images, n_subjects)).astype(np.int32)
    xyd = np.zeros((n_images, n_subjects, hi, 3), dtype=np.float64)
    xyd[..., 0] = rng.uniform(0, 320, (n_images, n_subjects, hi))
    xyd[..., 1] = rng.uniform(0, 240, (n_images, n_subjects, hi))
    xyd[..., 2] = np.exp(rng.normal(np.log(0.25), 0.4, (n_images, n_subjects, hi)))
    xyd *= (np.arange(hi)[None, None, :, None] < L[..., None, None])
    return xyd, L


def human_len_range(task):
    return (2, 6) if task == "COCO_Search18" else (6, 14)


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's own CPU path (the reference is pure Python;
# oracle/ restates it and is pinned to it by tests/golden) on a bounded sample.
# ---------------------------------------------------------------------------------------
def _score_chunk(args):
    from oracle import scoring as O
    humans, preds = args
    out = []
    for gts, p in zip(humans, preds):
        for g in gts:
            out.append(O.score_pair(g, p))
    return out


def _pool_score(humans, preds, pool, cores):
    chunks = max(1, min(len(preds), cores * 4))
    idx = np.array_split(np.arange(len(preds)), chunks)
    jobs = [([humans[i] for i in ix], [preds[i] for i in ix]) for ix in idx if len(ix)]
    res = pool.map(_score_chunk, jobs) if pool is not None else [_score_chunk(j) for j in jobs]
    return sum(len(r) for r in res)


def cpu_port_step(task, n_images, K, S, seed, pool, decode=True):
    """decode (torch CPU fp32, all threads: baseline.inference of `task`) + sample (K per head) + score
    (Python/numpy, one process per core).  Returns (scored scanpaths, seconds)."""
    import torch
    from oracle import decoder as OD
    from oracle import sampling as OSm
    from scanpaths_b200.weights import random_state_dict, synthetic_features
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict(task, 0)
    lo, hi = human_len_range(task)
    hx, hl = synth_humans(n_images, S, seed, lo, hi)
    rng = np.random.default_rng(seed)
    att = tasks = None
    if task == "OSIE":
        vf = synthetic_features(n_images, seed)
    else:
        vf, att = synthetic_features(n_images, seed, attention=True)
        if task == "COCO_Search18":
            tasks = torch.from_numpy(rng.integers(0, 18, n_images))
    t0 = time.perf_counter()
    with torch.no_grad():
        out = OD.decode(sd, vf, task, attention_maps=att, tasks=tasks, steps=T_STEPS)
    prefixes = ["good_", "poor_"] if task == "AiR" else [""]
    humans, preds = [], []
    for pre in prefixes:
        probs = out[pre + "all_actions_prob"].numpy()
        mu, s2 = out[pre + "log_normal_mu"].numpy(), out[pre + "log_normal_sigma2"].numpy()
        for k in range(K):
            q = rng.exponential(1.0, probs.shape).astype(np.float32)
            z = rng.standard_normal(mu.shape).astype(np.float32)
            s = OSm.random_sample(probs, mu, s2, q, z, 1)
            fix, _, _ = OSm.generate_scanpath(s["selected_actions"], s["durations"])
            for n in range(n_images):
                humans.append([hx[n, j, :hl[n, j]] * [1, 1, 1000.0] for j in range(S)])
                preds.append(fix[n] * [1, 1, 1000.0])
    n_pairs = _pool_score(humans, preds, pool, cores)
    return n_pairs // S, time.perf_counter() - t0


def cpu_human_eval(n_images, S, seed, pool):
    """human_evaluation (OSIE/utils/evaluation.py:11-148): S*(S-1) ordered pairs per image.  (pairs, seconds)."""
    cores = os.cpu_count() or 1
    hx, hl = synth_humans(n_images, S, seed)
    humans, preds = [], []
    for n in range(n_images):
        for j in range(S):
            humans.append([hx[n, i, :hl[n, i]] * [1, 1, 1000.0] for i in range(S) if i != j])
            preds.append(hx[n, j, :hl[n, j]] * [1, 1, 1000.0])
    t0 = time.perf_counter()
    n_pairs = _pool_score(humans, preds, pool, cores)
    return n_pairs, time.perf_counter() - t0


def cpu_scst_reward(n_images, K, S, seed, pool):
    """The host half of one SCST batch (train.py:223-258): K x (random_sample, generate_scanpath, pairs_eval) on
    given decoder outputs + the loss tail.  (scored scanpaths, seconds); the reference runs it on ONE thread --
    the pool makes this a generous baseline."""
    from oracle import sampling as OSm
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((n_images, T_STEPS, A)).astype(np.float32) * 2
    logits[:, :, 0] += 3.0
    probs = np.exp(logits - logits.max(-1, keepdims=True)); probs /= probs.sum(-1, keepdims=True)
    mu = np.full((n_images, T_STEPS), np.log(0.25), np.float32)
    s2 = np.full((n_images, T_STEPS), 0.15, np.float32)
    hx, hl = synth_humans(n_images, S, seed)
    t0 = time.perf_counter()
    humans, preds = [], []
    for k in range(K):
        q = rng.exponential(1.0, probs.shape).astype(np.float32)
        z = rng.standard_normal(mu.shape).astype(np.float32)
        s = OSm.random_sample(probs, mu, s2, q, z, 1)
        fix, am, dm = OSm.generate_scanpath(s["selected_actions"], s["durations"])
        OSm.log_action(s["selected_actions_probs"], am); OSm.log_duration(s["durations"], mu, s2, dm)
        for n in range(n_images):
            humans.append([hx[n, j, :hl[n, j]] * [1, 1, 1000.0] for j in range(S)])
            preds.append(fix[n] * [1, 1, 1000.0])
    n_pairs = _pool_score(humans, preds, pool, cores)
    return n_pairs // S, time.perf_counter() - t0


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_img = 8                                             # bounded sample per step: ~5 s on 16 cores
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup):
            cpu_port_step(args.task, n_img, args.samples, args.subjects, 100 + i, pool)
        total, secs = 0, 0.0
        for i in range(args.steps):
            n, s = cpu_port_step(args.task, n_img, args.samples, args.subjects, 200 + i, pool)
            total += n; secs += s
    v = total / secs
    sample = "%s: %d image(s) x %d samples%s x %d subjects per step (decode + sample + score), %d steps" % (
        args.task, n_img, args.samples, " x 2 heads" if args.task == "AiR" else "", args.subjects, args.steps)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(args, world, task=None, images=None, subjects=None):
    task = task or args.task
    images = images or args.images
    subjects = subjects or args.subjects
    per = images if args.scaling == "weak" else (images + world - 1) // world
    return {"workload": "%s-shaped decode+sample+score: %d images x %d samples%s x %d subjects %s, T=16, A=1201, "
                        "random-init weights with the SURVEY 8d bias calibration" %
                        (task, images, args.samples, " x 2 heads" if task == "AiR" else "", subjects,
                         "per GPU" if args.scaling == "weak" else "in total, sharded by image over the GPUs"),
            "task": task, "images_per_gpu": per, "samples": args.samples, "subjects": subjects, "wave": args.wave,
            "l2": "inputs (%.1f GB of feature maps per GPU and step) are larger than L2" % (per * 512 * 1200 * 4 / 1e9)}


# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
class Workload:
    """One task-shaped workload resident on one GPU: pipeline + synthetic features / humans."""

    def __init__(self, args, task, n_local, S, dev, rank, vf_dev=None, calibrated=True):
        import torch
        from scanpaths_b200.pipeline import ScanpathPipeline
        from scanpaths_b200.weights import random_state_dict
        self.task, self.N, self.K, self.S, self.dev, self.args = task, n_local, args.samples, S, dev, args
        self.pipe = ScanpathPipeline(random_state_dict(task, 0, calibrated=calibrated), task, T_STEPS, self.K, 1, dev,
                                     args.wave, seed=1234 + rank)
        self.heads = self.pipe.decoder.heads
        gen = torch.Generator(device=dev).manual_seed(1000 + rank)
        if vf_dev is None:                                # features relu(N(0,1)) generated on the device in chunks
            vf_dev = torch.empty((n_local, 512, 30, 40), dtype=torch.float32, device=dev)
            for n0 in range(0, n_local, 256):
                n1 = min(n_local, n0 + 256)
                vf_dev[n0:n1] = torch.randn((n1 - n0, 512, 30, 40), generator=gen, device=dev).clamp_min_(0)
        self.vf_dev = vf_dev
        lo, hi = human_len_range(task)
        self.hx, self.hl = synth_humans(n_local, S, 50 + rank, lo, hi)
        self.pipe.set_humans(self.hx, self.hl)
        self.att_dev = self.tasks_dev = None
        if task != "OSIE":                                # attention map ~ U(0,1)/max, task ~ U{0..17}
            self.att_dev = torch.rand((n_local, 1, 30, 40), generator=gen, device=dev)
            self.att_dev /= self.att_dev.amax(dim=(1, 2, 3), keepdim=True)
            if task == "COCO_Search18":
                self.tasks_dev = torch.randint(0, 18, (n_local,), generator=gen, device=dev)
        self.pinned = None
        torch.cuda.synchronize()

    def step(self, host=False):
        if host:
            vf_pin, hx_pin, hl_pin, paths_host = self.pinned
            self.pipe.set_humans(hx_pin, hl_pin)
            return self.pipe.run(vf_pin, self.att_dev, self.tasks_dev, paths_host=paths_host)
        return self.pipe.run(self.vf_dev, self.att_dev, self.tasks_dev)

    def pin(self):
        import torch
        if self.pinned is None:
            vf_pin = torch.empty(self.vf_dev.shape, dtype=torch.float32, pin_memory=True)
            vf_pin.copy_(self.vf_dev)
            hx_pin, hl_pin = torch.from_numpy(self.hx).pin_memory(), torch.from_numpy(self.hl).pin_memory()
            paths_host = (torch.empty((self.heads, self.K, self.N, T_STEPS, 3), dtype=torch.float64, pin_memory=True),
                          torch.empty((self.heads, self.K, self.N), dtype=torch.int32, pin_memory=True))
            self.tab_host = torch.empty((self.heads, self.K, self.N, 11), dtype=torch.float32, pin_memory=True)
            self.acc_host = torch.empty((self.heads, 16), dtype=torch.float64, pin_memory=True)
            self.pinned = (vf_pin, hx_pin, hl_pin, paths_host)
        return self.pinned


def timed_steps(fn, steps, barrier, dev):
    """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events on the launching stream."""
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    out = None
    for _ in range(steps):
        out = fn()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1), out


def profile_pass(lib, w, steps=1):
    """A separate pass with the library's per-kernel CUDA-event brackets on (kept out of the timed region)."""
    import ctypes as C
    import torch
    from scanpaths_b200 import _lib
    cap = 200 * (w.N // w.args.wave + 1) * steps * max(1, w.heads)
    _lib.check(lib.spb_profile_enable(cap), "spb_profile_enable")
    overlap, w.pipe.overlap_tail = w.pipe.overlap_tail, False      # one stream: every bracket times its kernels alone
    for _ in range(steps):
        w.step()
    torch.cuda.synchronize()
    w.pipe.overlap_tail = overlap
    ms_buf = np.zeros(cap, dtype=np.float32); tag_buf = np.zeros(cap, dtype=np.int32)
    n_out = C.c_int32(0)
    _lib.check(lib.spb_profile_collect(_lib.ptr(ms_buf), _lib.ptr(tag_buf), cap, C.byref(n_out)), "spb_profile_collect")
    lib.spb_profile_enable(0)
    return ms_buf[:n_out.value], tag_buf[:n_out.value]


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def gemm_roofline(ms_buf, tag_buf, wave_imgs, peaks):
    """Dominant kernel: wino_gemm_tc_kernel = the 24 per-position GEMMs of the Winograd F(2x4,3x3) form of the
    3x3 gate convolution of h, [wave*150 tiles x 512] x [512 x 2048] each."""
    conv_h = ms_buf[tag_buf == 2]
    if not len(conv_h):
        return None
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: the kernel runs inside a long step)" if peaks \
        else "fallback 1.4 PFLOP/s sustained"
    share = {TAGS[k]: float(ms_buf[tag_buf == k].sum()) for k in TAGS if (tag_buf == k).any()}
    tot = sum(share.values()) or 1.0
    flop_gemm = 24 * 2.0 * wave_imgs * 150 * 2048 * 512
    flop_direct = 2.0 * wave_imgs * 1200 * 2048 * 4608
    avg_ms = float(conv_h.mean())
    ach = flop_gemm / (avg_ms * 1e-3) / 1e12
    return {"kernel": "wino_gemm_tc_kernel (Winograd F(2x4,3x3) gate convolution: 24 per-position GEMMs, %d images "
                      "per launch)" % wave_imgs,
            "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            # DRAM bytes per launch from the committed ncu --set full capture (256 images: dram__bytes_read 2.02 GB +
            # dram__bytes_write 3.74 GB, profiles/r01_step_kernels_ncu_summary.csv); algorithmic: 1.89 GB of U
            # operands + 0.10 GB of weights + 3.77 GB of results
            "traffic": 5.762e9 * wave_imgs / 256.0, "traffic_unit": "B per launch (ncu)",
            "peak_source": which, "avg_launch_ms": avg_ms, "launches": int(len(conv_h)),
            "issued_tflops": 3 * ach, "issued_frac": 3 * ach / peak_tf,
            "direct_conv_equivalent_tflops": flop_direct / (avg_ms * 1e-3) / 1e12,
            "survey_8d_frac": flop_direct / (avg_ms * 1e-3) / 1e12 / peak_tf,
            "note": "achieved counts the ALGORITHMIC flops of this kernel (2*M*N*K of its 24 GEMMs = %.2f TFLOP per "
                    "launch); it issues 3 fp16 MMA flops per algorithmic flop (hi*hi, hi*lo, lo*hi) for fp32-equivalent "
                    "results, so frac tops out at 1/3; SURVEY 8(d) counts the same convolution done directly (%.2f "
                    "TFLOP): survey_8d_frac" % (flop_gemm / 1e12, flop_direct / 1e12),
            "time_share_of_tagged_kernels": {k: v / tot for k, v in share.items()}}


def pair_work(lh, lp, nh, npd):
    """Per-pair DP work (torch tensors, f64): NW cells (with + without duration), SED cells, STDE window updates."""
    import torch
    mn = torch.minimum(lh, lp)
    stde = mn * (lp + 1) * (lh + 1) - (lp + lh + 2) * mn * (mn + 1) / 2 + mn * (mn + 1) * (2 * mn + 1) / 6
    return nh * npd + lh * lp, lh * lp, stde


def scoring_roofline(cells_nw, cells_sed, cells_stde, n_pairs, ms, nbytes, peaks):
    """north_star: achieved DP cell-updates/s and HBM GB/s against the chip's peak.  The peak is kernel-independent:
    the fewest arithmetic instructions the recurrences need -- NW cell: 1 f64 add + 2 f64 max; STDE window update:
    1 f64 add + 1 f64 min; SED cell: 3 int32 ops -- on B200's FP64 pipe (64 lanes/clk/SM) and INT32/FP32 pipe
    (128 lanes/clk/SM) at the maximum SM clock."""
    sm_hz, sms = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6, 148
    t_min = (3 * cells_nw + 2 * cells_stde) / (sms * 64 * sm_hz) + 3 * cells_sed / (sms * 128 * sm_hz)
    cells = cells_nw + cells_sed + cells_stde
    ach = cells / (ms * 1e-3)
    peak = cells / t_min
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "score_pairs_g8_kernel (4 pairs per warp)", "bound": "fp64 pipe", "achieved": ach, "peak": peak,
            "unit": "DP cell-updates/s", "frac": ach / peak, "avg_launch_ms": ms, "pairs_per_s": n_pairs / (ms * 1e-3),
            "cell_updates_per_pair": cells / max(n_pairs, 1), "hbm_gb_s": gbs,
            "hbm_frac": gbs / float(peaks.get("hbm_gbs", 6541.5)),
            "note": "cell updates = n_wd*m_wd + 2*Lh*Lp + sum_k (Lp-k+1)(Lh-k+1) per pair; peak = the same cells at the "
                    "minimum instruction count (3 f64 per NW cell, 2 f64 per STDE update, 3 int32 per SED cell) on the "
                    "FP64 (64/clk/SM) and INT32 (128/clk/SM) pipes at %.0f MHz -- a kernel-independent bound; the "
                    "kernel is latency / issue bound (wavefront dependencies), not HBM bound" % (sm_hz / 1e6)}


def scoring_record(w, ms_buf, tag_buf, peaks):
    import torch
    from scanpaths_b200 import scoring as SC
    score_ms = ms_buf[tag_buf == 10]
    if not len(score_ms):
        return None
    pipe, K, S = w.pipe, w.K, w.S
    n = min(w.args.wave, w.N)
    # the DP work of one launch, counted on the first wave (its own humans, so that the pair map lines up)
    pipe.set_humans(w.hx[:n], w.hl[:n])
    o2 = pipe.run(w.vf_dev[:n], None if w.att_dev is None else w.att_dev[:n],
                  None if w.tasks_dev is None else w.tasks_dev[:n], keep_paths=True)
    nw = sed = stde = nbytes = 0.0
    for (hd, n0, n1, smp) in o2["paths"]:
        pp = SC.prep_paths(smp["xyd"], smp["len"], pipe.cfg)
        ph, ps = pipe._pair_map(n1 - n0, n0)
        lh, lp = pipe.humans.len[ph.long()].double(), pp.len[ps.long()].double()
        a, b, c = pair_work(lh, lp, pipe.humans.nwd[ph.long()].double(), pp.nwd[ps.long()].double())
        nw += float(a.sum()); sed += float(b.sum()); stde += float(c.sum())
        # unique bytes: both symbol packs (25 B per fixation slot + 8 B per path), pair map, scores
        nbytes += (pp.len.numel() + (n1 - n0) * S) * (16 * 25 + 8) + ph.numel() * (8 + 32)
    L = len(o2["paths"])
    pipe.set_humans(w.hx, w.hl)
    return scoring_roofline(nw / L, sed / L, stde / L, n * K * S, float(score_ms.mean()), nbytes / L, peaks)


def e2e_record(w, steps, barrier, world, n_total_per_step, dist):
    """Same metric through host buffers: every step copies the features + human scanpaths host -> device (pinned
    memory, overlapped with compute on a side stream) and reads back the reduced table, the aggregate AND every
    sampled scanpath (what test.py:135-152 dumps), all inside the timed region."""
    import torch
    vf_pin, hx_pin, hl_pin, paths_host = w.pin()

    def one():
        o = w.step(host=True)
        w.tab_host.copy_(o["table"], non_blocking=True)
        w.acc_host.copy_(o["acc"], non_blocking=True)
        torch.cuda.synchronize()
        return o

    one()
    ms, _ = timed_steps(one, steps, barrier, w.dev)
    tt = torch.tensor([ms], dtype=torch.float64, device=w.dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    h2d = int(vf_pin.numel() * 4 + hx_pin.numel() * 8 + hl_pin.numel() * 4)
    d2h = int(w.tab_host.numel() * 4 + w.acc_host.numel() * 8 + paths_host[0].numel() * 8 + paths_host[1].numel() * 4)
    # the upload on its own (nothing to hide behind): PCIe rate and what it would add to a step if it were serial
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chunk = vf_pin[:min(w.N, 1024)]
    torch.cuda.synchronize()
    t0.record()
    chunk.to(w.dev, non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    gbs = chunk.numel() * 4 / (t0.elapsed_time(t1) * 1e-3) / 1e9
    return {"value": n_total_per_step * steps / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": steps, "ms_per_step": ms / steps,
            "h2d_gb_s_alone": gbs, "h2d_seconds_per_step_if_serial": h2d / (gbs * 1e9),
            "note": "d2h includes every sampled scanpath (x, y, duration f64 + length), the reduced [heads,K,N,11] table "
                    "and the aggregate; the feature upload overlaps the previous wave's compute"}


def scst_record(args, dev, lib, cpu):
    """SCST reward step, per rank N = 4 images, K = 5 trials (+3 spare for the rejection rule), S = 15
    (OSIE/train.py:216-258): decode is inference here (the reference's forward-with-grad stays in PyTorch)."""
    import torch
    from scanpaths_b200.models.baseline_attention import CudaDecoder
    from scanpaths_b200.models.sampling import Sampling
    from scanpaths_b200.scst import ScstRewardStep
    from scanpaths_b200.weights import random_state_dict, synthetic_features
    N, K, S = 4, 5, 15
    dec = CudaDecoder(random_state_dict("OSIE", 0), "OSIE", T_STEPS, dev, wave=N)
    vf = synthetic_features(N, 3).to(dev)
    hx, hl = synth_humans(N, S, 77)
    step = ScstRewardStep(Sampling(convLSTM_length=T_STEPS, min_length=1, seed=9), dev, rl_sample_number=K, spare=3)
    step.set_humans(packed=(torch.from_numpy(hx), torch.from_numpy(hl), torch.full((N,), S, dtype=torch.int32)))
    probs, mu, s2, _ = dec.decode(vf)
    p0 = probs[0].clone().requires_grad_(True)
    m0, v0 = mu[0].clone().requires_grad_(True), s2[0].clone().requires_grad_(True)

    def reward_only():
        loss, aux = step(p0, m0, v0)
        loss.backward()
        return loss

    def with_decode():
        pr, mm, ss, _ = dec.decode(vf)
        loss, aux = step(pr[0], mm[0], ss[0])
        return loss

    def with_graphed_decode():                            # the 16-step rollout replayed from one CUDA graph
        pr, mm, ss, _ = dec.decode_graphed(vf)
        loss, aux = step(pr[0], mm[0], ss[0])
        return loss

    res = {}
    l0 = lib.spb_kernel_launches()
    reward_only(); torch.cuda.synchronize()
    res["launches_reward_step"] = int(lib.spb_kernel_launches() - l0)
    for name, fn, reps in (("reward_sample_score_loss_backward", reward_only, 50), ("decode_plus_reward", with_decode, 10),
                           ("graphed_decode_plus_reward", with_graphed_decode, 20)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        res[name] = {"ms_per_step": ms, "scanpaths_per_s": N * K / (ms * 1e-3)}
    rec = {"workload": "SCST reward step: N=4 images, K=5 trials (+3 spare), S=15 subjects, 300 scored pairs "
                       "(+180 spare); wall-clock per step incl. the host read of the accept count", **res}
    if cpu is not None:
        rec["cpu_baseline"] = cpu
        rec["speedup_reward_step_vs_cpu"] = res["reward_sample_score_loss_backward"]["scanpaths_per_s"] / cpu["value"]
    return rec


def human_eval_record(args, dev, peaks, cpu):
    """human_evaluation (OSIE/utils/evaluation.py:11-148): S*(S-1) ordered pairs per image, scoring only."""
    import torch
    from scanpaths_b200 import scoring as SC
    N, S = args.images, args.subjects
    cfg = SC.ScoreConfig.evaluation(device=dev, dur_scale=1000.0)
    hx, hl = synth_humans(N, S, 91)
    L = hx.shape[2]
    pack = SC.prep_paths(torch.from_numpy(hx.reshape(N * S, L, 3)).to(dev), torch.from_numpy(hl.reshape(-1)).to(dev), cfg)
    i, j = np.nonzero(~np.eye(S, dtype=bool))
    base = (np.arange(N) * S)[:, None]
    ph = torch.from_numpy((base + i[None]).reshape(-1).astype(np.int32)).to(dev)
    ps = torch.from_numpy((base + j[None]).reshape(-1).astype(np.int32)).to(dev)
    ws = SC.Workspace(int(pack.nwd.max().item()), dev)
    out = torch.empty((ph.numel(), 4), dtype=torch.float64, device=dev)
    for _ in range(2):
        SC.score_pairs(pack, pack, ph, ps, cfg, workspace=ws, out=out, check=False)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    reps = 5
    for _ in range(reps):
        SC.score_pairs(pack, pack, ph, ps, cfg, workspace=ws, out=out, check=False)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    lh, lp = pack.len[ph.long()].double(), pack.len[ps.long()].double()
    a, b, c = pair_work(lh, lp, pack.nwd[ph.long()].double(), pack.nwd[ps.long()].double())
    nbytes = pack.len.numel() * (L * 25 + 8) + ph.numel() * (8 + 32)
    rec = scoring_roofline(float(a.sum()), float(b.sum()), float(c.sum()), ph.numel(), ms, nbytes, peaks)
    rec["workload"] = "human_evaluation: %d images x %d subjects -> %d ordered pairs in one launch" % (N, S, ph.numel())
    rec["mean_scores"] = [float(x) for x in out.mean(0).cpu()]
    if cpu is not None:
        rec["cpu_baseline"] = cpu
        rec["speedup_vs_cpu"] = rec["pairs_per_s"] / cpu["value"]
    return rec


def stress_record(args, dev, vf_dev, barrier):
    """SURVEY 8d's long-string stress case: RAW random-init weights (no bias calibration) -- the model almost never
    stops (16 fixations per scanpath) and draws durations around exp(N(0, 1)) s, so the with-duration strings of
    the predictions are ~500 symbols instead of ~50 and every pair takes the warp-per-pair multi-panel kernel.
    One wave of images; decode + sample + score like the headline."""
    import torch
    n = min(args.wave, args.images)
    w = Workload(args, "OSIE", n, args.subjects, dev, 0, vf_dev=vf_dev[:n], calibrated=False)
    w.step()
    steps = 2
    ms, out = timed_steps(w.step, steps, barrier, dev)
    from scanpaths_b200.pipeline import ScanpathPipeline
    m, _ = ScanpathPipeline.metrics(out)
    res = w.pipe.run(w.vf_dev, keep_paths=True)
    smp = res["paths"][0][3]
    from scanpaths_b200 import scoring as SC
    pp = SC.prep_paths(smp["xyd"], smp["len"], w.pipe.cfg)
    rec = {"workload": "OSIE-shaped, raw random-init weights: %d images x %d samples x %d subjects" % (n, w.K, w.S),
           "value": n * w.K * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
           "pred_fixations_mean": float(smp["len"].float().mean()),
           "pred_wd_string_mean": float(pp.nwd.float().mean()), "pred_wd_string_max": int(pp.nwd.max()),
           "human_wd_string_mean": float(w.pipe.humans.nwd.float().mean()),
           "scores": {"ScanMatch_wd": m["ScanMatch"]["with duration"], "ScanMatch_wod": m["ScanMatch"]["w/o duration"],
                      "SED": m["VAME"]["SED"], "STDE": m["VAME"]["STDE"]}}
    del w
    torch.cuda.empty_cache()
    return rec


def extra_task_record(args, task, dev, lib, peaks, cpu, vf_dev, do_e2e, barrier):
    """A task-shaped 1-GPU record (AiR: two streams / two heads -> 2K samples per image; COCO: per-image 5x5 weights)."""
    import torch
    S = 10 if task == "COCO_Search18" else args.subjects       # S = 10
    w = Workload(args, task, args.images, S, dev, 0, vf_dev=vf_dev)
    n_per_step = w.N * w.K * w.heads
    w.step(); w.step()
    steps = 2
    ms, out = timed_steps(w.step, steps, barrier, dev)
    from scanpaths_b200.pipeline import ScanpathPipeline
    m, _ = ScanpathPipeline.metrics(out)
    ms_buf, tag_buf = profile_pass(lib, w)
    rec = {"metric": METRIC, "value": n_per_step * steps / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": steps,
           "warmup": 2, "ms_per_step": ms / steps, "config": workload_config(args, 1, task, args.images, S),
           "roofline": gemm_roofline(ms_buf, tag_buf, min(args.wave, w.N), peaks),
           "scores": {"ScanMatch_wd": m["ScanMatch"]["with duration"], "ScanMatch_wod": m["ScanMatch"]["w/o duration"],
                      "SED": m["VAME"]["SED"], "STDE": m["VAME"]["STDE"]}}
    if do_e2e:
        rec["e2e"] = e2e_record(w, steps, barrier, 1, n_per_step, None)
    if cpu is not None:
        rec["cpu_baseline"] = cpu
        rec["speedup_vs_cpu"] = (rec["e2e"]["value"] if do_e2e else rec["value"]) / cpu["value"]
    del w
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    extras_on = world == 1 and not args.no_extras and args.task == "OSIE"

    # ---- CPU baselines first (rank 0, N=1 only; forks worker processes before CUDA is touched)
    cpu = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1

        def rec(value, unit, sample, secs):
            return {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": "%s, %.1f s" % (sample, secs)}
        with mp.get_context("fork").Pool(cores) as pool:
            n, s = cpu_port_step(args.task, args.cpu_images, args.samples, args.subjects, 7, pool)
            cpu["main"] = rec(n / s, UNIT, "%s: %d images x %d samples x %d subjects (decode + sample + score)" %
                              (args.task, args.cpu_images, args.samples, args.subjects), s)
            if extras_on:
                n, s = cpu_port_step("AiR", max(2, args.cpu_images // 4), args.samples, args.subjects, 8, pool)
                cpu["AiR"] = rec(n / s, UNIT, "AiR: %d images x %d samples x 2 heads x %d subjects (decode + sample + "
                                 "score)" % (max(2, args.cpu_images // 4), args.samples, args.subjects), s)
                n, s = cpu_port_step("COCO_Search18", max(2, args.cpu_images // 2), args.samples, 10, 9, pool)
                cpu["COCO_Search18"] = rec(n / s, UNIT, "COCO-Search18: %d images x %d samples x 10 subjects (decode + "
                                           "sample + score)" % (max(2, args.cpu_images // 2), args.samples), s)
                n, s = cpu_scst_reward(4, 5, 15, 10, pool)
                cpu["scst"] = rec(n / s, UNIT, "SCST reward step: 4 images x 5 trials x 15 subjects (sample + pairs_eval "
                                  "+ log-likelihoods; one process per core, the reference uses one)", s)
                n, s = cpu_human_eval(4, args.subjects, 11, pool)
                cpu["human_eval"] = rec(n / s, "pairs/s", "human_evaluation: 4 images x %d ordered pairs" %
                                        (args.subjects * (args.subjects - 1)), s)

    import torch
    import torch.distributed as dist
    from scanpaths_b200 import _lib, build
    if build.needs_build():
        if local_rank == 0:
            build.build_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    lib = _lib.load()
    from scanpaths_b200.dist import allgather_tables, shard_range
    from scanpaths_b200.pipeline import ScanpathPipeline

    K, S = args.samples, args.subjects
    if args.scaling == "strong":
        lo, hi = shard_range(args.images, rank, world)
        n_local, n_total = hi - lo, args.images
    else:
        n_local, n_total = args.images, args.images * world
    w = Workload(args, args.task, n_local, S, dev, rank)
    heads = w.heads
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [None]

    def step():
        out = w.step()
        if world > 1:                                     # the one collective of the path: score tables
            gathered[0] = allgather_tables(out["table"], n_total, image_dim=2)
        return out

    for _ in range(args.warmup):
        out = step()
    barrier()

    # ---- timed region: K steps, device-resident inputs, no profiler brackets
    launches0 = lib.spb_kernel_launches()
    clocks = ClockSampler(local_rank)
    total_ms, out = timed_steps(step, args.steps, barrier, dev)
    clock_info = clocks.stop()
    launches = lib.spb_kernel_launches() - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_total * K * heads * args.steps / (total_ms / 1e3)
    m, s_ = ScanpathPipeline.metrics(out)

    # ---- per-kernel live timings in a separate pass: roofline of the dominant kernel + the scoring kernel
    ms_buf, tag_buf = profile_pass(lib, w)
    roofline = gemm_roofline(ms_buf, tag_buf, min(args.wave, n_local), peaks)
    scoring = scoring_record(w, ms_buf, tag_buf, peaks)

    # ---- e2e: host buffers in, host results out, copies inside the timed region, all --steps
    e2e = None
    if not args.no_e2e:
        e2e = e2e_record(w, max(1, args.steps), barrier, world, n_total * K * heads, dist)

    extra = None
    if extras_on and rank == 0:
        extra = {}
        vf_dev = w.vf_dev
        del w
        torch.cuda.empty_cache()
        for name, fn in (("AiR", lambda: extra_task_record(args, "AiR", dev, lib, peaks, cpu.get("AiR"), vf_dev, True, barrier)),
                         ("COCO_Search18", lambda: extra_task_record(args, "COCO_Search18", dev, lib, peaks,
                                                                     cpu.get("COCO_Search18"), vf_dev, False, barrier)),
                         ("scst_reward_step", lambda: scst_record(args, dev, lib, cpu.get("scst"))),
                         ("human_evaluation", lambda: human_eval_record(args, dev, peaks, cpu.get("human_eval"))),
                         ("raw_random_init_stress", lambda: stress_record(args, dev, vf_dev, barrier))):
            try:
                extra[name] = fn()
            except Exception as e:                        # an extra record must never cost the headline line
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": workload_config(args, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info,
                "roofline": roofline, "scoring_roofline": scoring, "cpu_baseline": cpu.get("main"),
                "scores": {"ScanMatch_wd": m["ScanMatch"]["with duration"], "ScanMatch_wod": m["ScanMatch"]["w/o duration"],
                           "SED": m["VAME"]["SED"], "STDE": m["VAME"]["STDE"]},
                "extra": extra}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    _quiet_stdout()
    main()
