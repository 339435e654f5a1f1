#!/usr/bin/env python
"""bench.py -- scored scanpaths/sec of the decode + sample + ScanMatch/SED/STDE hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over the OSIE-shaped workload of BASELINE.json
configs[1]: 4096 images x 64 sampled scanpaths x 15 human subjects per GPU (weak scaling:
every rank owns its own 4096-image shard; the reduced score tables are exchanged with one
NCCL all-gather).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import math
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
    ap.add_argument("--images", type=int, default=4096, help="images per GPU (default = BASELINE configs[1])")
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--subjects", type=int, default=15)
    ap.add_argument("--wave", type=int, default=256)
    ap.add_argument("--task", default="OSIE", choices=["OSIE", "AiR", "COCO_Search18"],
                    help="model variant of the workload (default: the OSIE-shaped configs[1] the metric is quoted on)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_port_step(n_images, K, S, seed, pool):
    """decode (torch CPU fp32, all threads) + sample + score (Python/numpy, one process per core).
    Returns (scored scanpaths, seconds)."""
    import torch
    from oracle import decoder as OD
    from oracle import sampling as OSm
    from scanpaths_b200.weights import random_state_dict, synthetic_features
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict("OSIE", 0)
    vf = synthetic_features(n_images, seed)
    hx, hl = synth_humans(n_images, S, seed)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        out = OD.decode(sd, vf, "OSIE", steps=T_STEPS)
    probs = out["all_actions_prob"].numpy()
    mu, s2 = out["log_normal_mu"].numpy(), out["log_normal_sigma2"].numpy()
    humans, preds = [], []
    for k in range(K):
        q = rng.exponential(1.0, probs.shape).astype(np.float32)
        z = rng.standard_normal(mu.shape).astype(np.float32)
        s = OSm.random_sample(probs, mu, s2, q, z, 1)
        fix, _, _ = OSm.generate_scanpath(s["selected_actions"], s["durations"])
        for n in range(n_images):
            humans.append([hx[n, j, :hl[n, j]] * [1, 1, 1000.0] for j in range(S)])
            preds.append(fix[n] * [1, 1, 1000.0])
    chunks = max(1, min(len(preds), cores * 4))
    idx = np.array_split(np.arange(len(preds)), chunks)
    jobs = [([humans[i] for i in ix], [preds[i] for i in ix]) for ix in idx if len(ix)]
    res = pool.map(_score_chunk, jobs) if pool is not None else [_score_chunk(j) for j in jobs]
    n_scored = sum(len(r) for r in res) // S
    return n_scored, time.perf_counter() - t0


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_img = 8                                             # bounded sample per step: ~5 s on 16 cores
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup):
            cpu_port_step(n_img, args.samples, args.subjects, 100 + i, pool)
        total, secs = 0, 0.0
        for i in range(args.steps):
            n, s = cpu_port_step(n_img, args.samples, args.subjects, 200 + i, pool)
            total += n; secs += s
    v = total / secs
    sample = "%d image(s) x %d samples x %d subjects per step (decode + sample + score), %d steps" % (
        n_img, args.samples, args.subjects, args.steps)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def workload_config(args):
    return {"workload": "%s-shaped decode+sample+score: %d images x %d samples%s x %d subjects per GPU, T=16, "
                        "A=1201, random-init weights with the SURVEY 8d bias calibration" %
                        (args.task, args.images, args.samples, " x 2 heads" if args.task == "AiR" else "", args.subjects),
            "task": args.task,
            "images_per_gpu": args.images, "samples": args.samples, "subjects": args.subjects,
            "wave": args.wave, "l2": "inputs (%.1f GB of feature maps per step) are larger than L2" %
                                     (args.images * 512 * 1200 * 4 / 1e9)}


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    # ---- CPU baseline first (rank 0, N=1 only; forks worker processes before CUDA is touched)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            n, s = cpu_port_step(args.cpu_images, args.samples, args.subjects, 7, pool)
        cpu_baseline = {"value": n / s, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d images x %d samples x %d subjects (decode + sample + score), %.1f s" %
                                  (args.cpu_images, args.samples, args.subjects, s)}

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
    from scanpaths_b200.dist import allgather_tables
    from scanpaths_b200.pipeline import ScanpathPipeline
    from scanpaths_b200.weights import random_state_dict

    N, K, S = args.images, args.samples, args.subjects
    pipe = ScanpathPipeline(random_state_dict(args.task, 0), args.task, T_STEPS, K, 1, dev, args.wave, seed=1234 + rank)
    heads = pipe.decoder.heads                            # AiR: good + poor head -> 2K samples per image
    # synthetic inputs: features relu(N(0,1)) generated on the device in chunks, a pinned host copy for e2e
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    vf_dev = torch.empty((N, 512, 30, 40), dtype=torch.float32, device=dev)
    for n0 in range(0, N, 256):
        n1 = min(N, n0 + 256)
        vf_dev[n0:n1] = torch.randn((n1 - n0, 512, 30, 40), generator=gen, device=dev).clamp_min_(0)
    if args.task == "COCO_Search18":
        hx, hl = synth_humans(N, S, 50 + rank, lo=2, hi=6)
    else:
        hx, hl = synth_humans(N, S, 50 + rank)
    pipe.set_humans(hx, hl)
    att_dev = tasks_dev = None
    if args.task != "OSIE":                               # attention map ~ U(0,1)/max, task ~ U{0..17}
        att_dev = torch.rand((N, 1, 30, 40), generator=gen, device=dev)
        att_dev /= att_dev.amax(dim=(1, 2, 3), keepdim=True)
        if args.task == "COCO_Search18":
            tasks_dev = torch.randint(0, 18, (N,), generator=gen, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [None]

    def step(vf, repack_humans=False):
        if repack_humans:
            pipe.set_humans(hx_pin, hl_pin)
        out = pipe.run(vf, att_dev, tasks_dev)
        if world > 1:                                     # the one collective of the path: score tables
            gathered[0] = allgather_tables(out["table"], world * N, image_dim=2)
        return out

    for _ in range(args.warmup):
        out = step(vf_dev)
    barrier()

    # ---- timed region: K steps, device-resident inputs
    _lib.check(lib.spb_profile_enable(200 * (N // args.wave + 1) * max(args.steps, 1)), "spb_profile_enable")
    launches0 = lib.spb_kernel_launches()
    clocks = ClockSampler(local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        out = step(vf_dev)
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    clock_info = clocks.stop()
    launches = lib.spb_kernel_launches() - launches0
    # live per-launch timings of the tagged kernels
    cap = 200 * (N // args.wave + 1) * max(args.steps, 1)
    ms_buf = np.zeros(cap, dtype=np.float32); tag_buf = np.zeros(cap, dtype=np.int32)
    import ctypes as C
    n_out = C.c_int32(0)
    _lib.check(lib.spb_profile_collect(_lib.ptr(ms_buf), _lib.ptr(tag_buf), cap, C.byref(n_out)), "spb_profile_collect")
    lib.spb_profile_enable(0)
    ms_buf, tag_buf = ms_buf[:n_out.value], tag_buf[:n_out.value]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * N * K * heads * args.steps / (total_ms / 1e3)
    m, s_ = ScanpathPipeline.metrics(out)

    # ---- roofline of the dominant kernel: the 3x3 gate convolution (tag 2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    names = {1: "conv3x3_x", 2: "winograd_gemm_h", 3: "conv5x5", 4: "lstm_cell", 5: "head", 6: "feedback",
             7: "rank1", 8: "prep", 9: "wino_input", 10: "score_pairs", 11: "sample"}
    share = {names[k]: float(ms_buf[tag_buf == k].sum()) for k in names if (tag_buf == k).any()}
    tot_tagged = sum(share.values()) or 1.0
    conv_h = ms_buf[tag_buf == 2]
    wave_imgs = min(args.wave, N)
    # dominant kernel: wino_gemm_tc_kernel = the 24 per-position GEMMs of the Winograd F(2x4,3x3) form of the
    # 3x3 gate convolution, [wave*150 tiles x 512] x [512 x 2048] each
    flop_gemm = 24 * 2.0 * wave_imgs * 150 * 2048 * 512
    flop_direct = 2.0 * wave_imgs * 1200 * 2048 * 4608
    roofline = None
    if len(conv_h):
        avg_ms = float(conv_h.mean())
        ach = flop_gemm / (avg_ms * 1e-3) / 1e12
        roofline = {"kernel": "wino_gemm_tc_kernel (Winograd F(2x4,3x3) gate convolution: 24 per-position GEMMs, "
                              "%d images per launch)" % wave_imgs,
                    "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    # DRAM bytes per launch from the committed ncu --set full capture (256 images:
                    # dram__bytes_read 2.02 GB + dram__bytes_write 3.74 GB, profiles/r01_step_kernels_ncu_summary.csv);
                    # algorithmic: 1.89 GB of U operands + 0.10 GB of weights + 3.77 GB of results
                    "traffic": 5.762e9 * wave_imgs / 256.0, "traffic_unit": "B per launch (ncu, round 1)",
                    "peak_source": which, "avg_launch_ms": avg_ms, "launches": int(len(conv_h)),
                    "issued_tflops": 3 * ach, "issued_frac": 3 * ach / peak_tf,
                    "direct_conv_equivalent_tflops": flop_direct / (avg_ms * 1e-3) / 1e12,
                    "note": "achieved counts the ALGORITHMIC flops of this kernel (2*M*N*K of its 24 GEMMs = %.2f "
                            "TFLOP per launch); it issues 3 fp16 MMA flops per algorithmic flop (hi*hi, hi*lo, lo*hi) "
                            "for fp32-equivalent results, so frac tops out at 1/3; the same convolution done directly "
                            "would need %.2f TFLOP" % (flop_gemm / 1e12, flop_direct / 1e12),
                    "time_share_of_tagged_kernels": {k: v / tot_tagged for k, v in share.items()}}

    # ---- scoring kernel: DP cell-updates/s and HBM GB/s (north_star), from the live brackets of score_pairs_kernel
    scoring = None
    score_ms = ms_buf[tag_buf == 10]
    if len(score_ms):
        from scanpaths_b200 import scoring as SC
        w_n = min(args.wave, N)
        o2 = pipe.run(vf_dev[:w_n], None if att_dev is None else att_dev[:w_n],
                      None if tasks_dev is None else tasks_dev[:w_n], keep_paths=True)
        cells = 0.0
        nbytes = 0.0
        for (hd, n0, n1, smp) in o2["paths"]:
            pp = SC.prep_paths(smp["xyd"], smp["len"], pipe.cfg)
            ph, ps = pipe._pair_map(n1 - n0, n0)
            lh, lp = pipe.humans.len[ph.long()].double(), pp.len[ps.long()].double()
            mn = torch.minimum(lh, lp)
            stde = mn * (lp + 1) * (lh + 1) - (lp + lh + 2) * mn * (mn + 1) / 2 + mn * (mn + 1) * (2 * mn + 1) / 6
            cells += float((pipe.humans.nwd[ph.long()].double() * pp.nwd[ps.long()].double() + 2 * lh * lp + stde).sum())
            # unique bytes: both symbol packs (25 B per fixation slot + 8 B per path), pair map, scores
            nbytes += (pp.len.numel() + (n1 - n0) * S) * (16 * 25 + 8) + ph.numel() * (8 + 32)
        per_launch = cells / len(o2["paths"])
        avg_ms = float(score_ms.mean())
        ach = per_launch / (avg_ms * 1e-3)
        # ncu (profiles/r01_score_pairs_ncu.txt): 4.06 warp instructions per cell update, 16 of 32 lanes active
        # (wavefront ramps of ~50-symbol strings); issue peak = 148 SMs x 4 warp-instr/clk x 1.965 GHz / 4.06
        peak = 148 * 4 * 1.965e9 / 4.06
        scoring = {"kernel": "score_pairs_kernel", "bound": "sm_issue", "achieved": ach, "peak": peak,
                   "unit": "DP cell-updates/s", "frac": ach / peak, "avg_launch_ms": avg_ms,
                   "cell_updates_per_pair": per_launch / (w_n * K * S),
                   "hbm_gb_s": nbytes / len(o2["paths"]) / (avg_ms * 1e-3) / 1e9,
                   "hbm_frac": nbytes / len(o2["paths"]) / (avg_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6541.5)),
                   "note": "cell updates = n_wd*m_wd + 2*Lh*Lp + sum_k (Lp-k+1)(Lh-k+1) per pair, counted on one wave"}

    # ---- e2e: host buffers in, host results out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        vf_pin = torch.empty((N, 512, 30, 40), dtype=torch.float32, pin_memory=True)
        vf_pin.copy_(vf_dev)
        hx_pin = torch.from_numpy(hx).pin_memory(); hl_pin = torch.from_numpy(hl).pin_memory()
        tab_host = torch.empty((pipe.decoder.heads, K, N, 11), dtype=torch.float32, pin_memory=True)
        acc_host = torch.empty((pipe.decoder.heads, 12), dtype=torch.float64, pin_memory=True)
        e_steps = max(1, min(2, args.steps))

        def e2e_step():
            o = step(vf_pin, repack_humans=True)
            tab_host.copy_(o["table"], non_blocking=True)
            acc_host.copy_(o["acc"], non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e_steps):
            e2e_step()
        t1.record()
        barrier()
        tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * K * heads * e_steps / (float(tt.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(vf_pin.numel() * 4 + hx_pin.numel() * 8 + hl_pin.numel() * 4),
               "d2h_bytes_per_step": int(tab_host.numel() * 4 + acc_host.numel() * 8), "steps": e_steps}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": workload_config(args), "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info,
                "roofline": roofline, "scoring_roofline": scoring, "cpu_baseline": cpu_baseline,
                "scores": {"ScanMatch_wd": m["ScanMatch"]["with duration"], "ScanMatch_wod": m["ScanMatch"]["w/o duration"],
                           "SED": m["VAME"]["SED"], "STDE": m["VAME"]["STDE"]}}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    _quiet_stdout()
    main()
