"""Drop-in for the reference's model entry points on the GPU.

``baseline`` keeps the constructor / ``forward`` signatures and the state_dict key
names of the three reference models

    OSIE/models/baseline_attention.py            baseline(images)
    AiR/models/baseline_attention.py             baseline(images, attention_maps[, performances])
    COCO_Search18/models/baseline_attention_multihead.py  baseline(images, attention_maps, tasks)

so a reference checkpoint loads unchanged (utils/checkpointing.py:79-110).  The
once-per-image encoder (dilated ResNet-50 + sal_conv, :191-194, :327-328) stays in
PyTorch; everything from ``state = init_hidden`` on (:333-396) runs in the CUDA
decoder (csrc/decode.cu + csrc/conv_tc.cu) through ``spb_decode``.  ``decode()``
takes ``visual_feature`` directly (what bench.py and the tests feed).

Inference only: the reference's SCST loop back-propagates through this forward
(train.py:216, 256); that differentiable path stays with PyTorch and is out of
scope here -- calling with autograd enabled raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ..weights import COCO_OBJECTS

E, HW, A = 512, 1200, 1201
GATES_H = ("input_h", "forget_h", "output_h", "memory_h")
GATES_X = ("input_x", "forget_x", "output_x", "memory_x")


DecoderWeights, DecoderIO = _lib.DecoderWeights, _lib.DecoderIO


def split_pair(w: torch.Tensor):
    """fp32 tensor -> (hi, lo) fp16 with w * scale = hi + lo / 2^11, scale a power of two that puts
    max|w| in [32, 64).  Returns (hi, lo, 1/scale)."""
    mx = float(w.abs().max())
    scale = 2.0 ** (5 - math.floor(math.log2(mx))) if mx > 0 else 1.0
    ws = w.double() * scale
    hi = ws.to(torch.float16)
    lo = ((ws - hi.double()) * 2048.0).to(torch.float16)
    return hi.contiguous(), lo.contiguous(), 1.0 / scale


def _conv_to_gemm(w):
    """[co, ci, kh, kw] -> [co, (ky*kw + kx)*ci_count + ci]"""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _interleave_gates(mats):
    """4 x [512, K] (i, f, o, g) -> [2048, K] rows ordered [64-channel block][32-channel half][gate][32]
    (csrc/decoder.cuh::gate_col): one epilogue thread of the tensor-core kernel then owns all four
    gates of 32 channels of its pixel."""
    g = torch.stack(mats, 0)                                  # [4, 512, K]
    return g.view(4, 8, 2, 32, -1).permute(1, 2, 0, 3, 4).reshape(2048, -1).contiguous()


def prepare_weights(sd, task: str, device):
    """Reference state_dict (fp32, any device) -> prepared device tensors + the C struct.
    All of the once-per-checkpoint algebra (gate interleaving, Winograd weight transform, the composed head,
    fp16-pair splits) runs in float64 on the HOST and the results are copied to the device: no library GEMM /
    convolution kernel (cuBLAS, cuDNN) is ever launched by this package, the only kernels on the device are its own."""
    target = torch.device(device)
    device = torch.device("cpu")
    f = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    streams = ["_pos", "_neg"] if task == "AiR" else [""]
    if task == "AiR":
        sets = ["performance_sal_layer.True", "performance_sal_layer.False"]      # head 0 = good, 1 = poor
    elif task == "COCO_Search18":
        sets = ["object_sal_layer.%s" % o for o in COCO_OBJECTS]
    else:
        sets = ["performance_sal_layer"]
    t = {}
    wx = _interleave_gates([_conv_to_gemm(f("lstm.%s.weight" % g)) for g in GATES_X])
    wh = _interleave_gates([_conv_to_gemm(f("lstm.%s.weight" % g)) for g in GATES_H])
    wp = torch.cat([_conv_to_gemm(f(s + ".weight")) for s in sets], 0).contiguous()
    # Winograd F(2x4,3x3) weights of the h-gates: W'[pos = 4j+i] = (G2 g G4^T)[i][j] (i: row position of F(2,3),
    # j: column position of F(4,3) on the points 0, +-1, +-2, inf), composed in float64, rows gate-interleaved
    # like wh, position-major: [24 * 2048, 512]
    G2 = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64, device=device)
    G4 = torch.tensor([[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
                       [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]], dtype=torch.float64, device=device)
    def wino_rows(gates):
        mats = [torch.einsum("ia,ocab,jb->ojic", G2, f("lstm.%s.weight" % g).double(), G4).reshape(512, 24 * 512)
                for g in gates]
        return _interleave_gates(mats).view(2048, 24, 512).permute(1, 0, 2).reshape(24 * 2048, 512)
    t["ww_hi"], t["ww_lo"], isw = split_pair(wino_rows(GATES_H))
    t["wwx_hi"], t["wwx_lo"], iswx = split_pair(wino_rows(GATES_X))
    # Winograd F(2x2,3x3) weights of the x-gates (the product path's loop-invariant convolution): G2 g G2^T,
    # 16 positions 4j+i, [16 * 2048, 512]
    mats22 = [torch.einsum("ia,ocab,jb->ojic", G2, f("lstm.%s.weight" % g).double(), G2).reshape(512, 16 * 512)
              for g in GATES_X]
    t["wwx2_hi"], t["wwx2_lo"], iswx2 = split_pair(
        _interleave_gates(mats22).view(2048, 16, 512).permute(1, 0, 2).reshape(16 * 2048, 512))
    t["wx_hi"], t["wx_lo"], isx = split_pair(wx)
    t["wh_hi"], t["wh_lo"], ish = split_pair(wh)
    t["wp_hi"], t["wp_lo"], isp = split_pair(wp)
    biases = []
    for gi, (gx, gh) in enumerate(zip(GATES_X, GATES_H)):
        b = f("lstm.%s.bias" % gx) + f("lstm.%s.bias" % gh)
        if gi < 3:
            for s in streams:
                b = b + f("lstm.%s%s.bias" % (gx[:-2], s))
        biases.append(b.view(512, 1))
    t["bias_gate"] = _interleave_gates(biases).view(-1).contiguous()
    t["bias_p"] = torch.cat([f(s + ".bias") for s in sets], 0).contiguous()
    wm = [f("lstm.%s%s.weight" % (g, s)).permute(0, 2, 3, 1).reshape(512, 9, 512)
          for s in streams for g in ("input", "forget", "output")]
    t["wm"] = torch.stack(wm, 0).reshape(-1, 512).contiguous()
    t["wm_hi"], t["wm_lo"], ism = split_pair(t["wm"])
    t["w2"] = f("object_head.sal_layer_2.weight").reshape(512).contiguous()
    t["w3"] = f("object_head.sal_layer_3.weight").reshape(512).contiguous()
    t["wd1"] = f("object_head.drt_layer_1.weight")[0].permute(1, 2, 0).reshape(49, 512).contiguous()
    t["wd2"] = f("object_head.drt_layer_2.weight").reshape(2, 48).contiguous()
    t["w_spatial_embed"] = f("spatial_embed.weight").contiguous()
    t["b_spatial_embed"] = f("spatial_embed.bias").contiguous()
    t["w_semantic_embed"] = f("semantic_embed.weight").contiguous()
    t["wse_hi"], t["wse_lo"], isse = split_pair(t["w_semantic_embed"])
    t["b_semantic_embed"] = f("semantic_embed.bias").contiguous()
    # composed head: feat = conv5x5(h) + bp is consumed only by linear maps (sal_layer_2, sal_layer_3 1x1,
    # drt_layer_1 7x7 stride 5) before any nonlinearity (predict_head.forward :144-150), so
    #   y2/y3[p] = sum_{tap,ci} (sum_c w[c] Wp[c,ci,tap]) h[ci,p+tap] + (w.bp + b)
    #   drt[oy,ox] = sum_{ci,11x11} (wd (*) Wp)[ci,dy,dx] h[ci,5oy-4+dy,5ox-4+dx] + const, where drt taps that
    #   land on the zero padding of feat (top row / left column windows) are masked out -> 4 variants.
    # Composed in float64, stored float32.
    wp64 = torch.stack([f(s_ + ".weight").double() for s_ in sets], 0)           # [sets, c, ci, 5, 5]
    bp64 = torch.stack([f(s_ + ".bias").double() for s_ in sets], 0)             # [sets, c]
    w2_64 = f("object_head.sal_layer_2.weight").double().reshape(512)
    w3_64 = f("object_head.sal_layer_3.weight").double().reshape(512)
    wd_64 = f("object_head.drt_layer_1.weight").double()[0]                      # [c, 7, 7]
    w23 = torch.stack([torch.einsum("c,scikl->skli", w2_64, wp64), torch.einsum("c,scikl->skli", w3_64, wp64)], -1)
    b23 = torch.stack([bp64 @ w2_64 + f("object_head.sal_layer_2.bias").double()[0],
                       bp64 @ w3_64 + f("object_head.sal_layer_3.bias").double()[0]], -1)
    t["b23_eff"] = b23.float().contiguous()
    wde, bde = [], []
    for top in (0, 1):
        for left in (0, 1):
            m = wd_64.clone()
            if top:
                m[:, :2, :] = 0
            if left:
                m[:, :, :2] = 0
            # full 2-D convolution of the masked 7x7 with each 5x5, summed over c -> [sets, ci, 11, 11]
            comp = torch.stack([F.conv_transpose2d(m[None], wp64[i])[0] for i in range(len(sets))], 0)
            wde.append(comp.permute(0, 2, 3, 1).reshape(len(sets), 121, 512))
            bde.append(bp64 @ m.sum(dim=(1, 2)) + f("object_head.drt_layer_1.bias").double()[0])
    t["wd_eff"] = torch.stack(wde, 1).float().contiguous()                       # [sets, 4, 121, 512]
    # the head as rows of a per-pixel GEMM, 256 rows per set: [0, 50) = (tap*2 + map) of the 5x5 -> 2 maps,
    # [128, 249) = the 121 taps of the interior variant of the duration convolution, the rest zero
    w23g = torch.zeros((len(sets), 256, 512), dtype=torch.float64, device=device)
    w23g[:, :50] = w23.reshape(len(sets), 25, 512, 2).permute(0, 1, 3, 2).reshape(len(sets), 50, 512)
    w23g[:, 128:249] = wde[0]
    t["w23_hi"], t["w23_lo"], is23 = split_pair(w23g.reshape(len(sets) * 256, 512))
    t["bd_eff"] = torch.stack(bde, 1).float().contiguous()                       # [sets, 4]
    # spatial_att: score_j = <spatial_attention, conv3x3(spatial_lists, list_j)> + const
    #            = <w_eff, list_j> + const  (adjoint of the 3x3 correlation applied to spatial_attention)
    watt = f("spatial_att.spatial_attention.weight").double().view(1, 1, 30, 40)
    kl = f("spatial_att.spatial_lists.weight").double().view(1, 1, 3, 3)
    t["w_eff_spatial"] = F.conv_transpose2d(watt, kl, padding=1).reshape(1200).float().contiguous()
    # semantic_att: score_j = semantic_attention . (semantic_lists list_j) + const = <u, list_j> + const
    t["u_semantic"] = (f("semantic_att.semantic_attention.weight").double().view(1, 512)
                       @ f("semantic_att.semantic_lists.weight").double()).reshape(512).float().contiguous()
    sc = lambda k: float(sd[k].detach().reshape(-1)[0])
    t = {k: v.contiguous().to(target) for k, v in t.items()}
    w = DecoderWeights()
    for k, v in t.items():
        setattr(w, k, v.data_ptr())
    w.b2, w.b3, w.bd1 = sc("object_head.sal_layer_2.bias"), sc("object_head.sal_layer_3.bias"), sc(
        "object_head.drt_layer_1.bias")
    bd2 = sd["object_head.drt_layer_2.bias"].detach().reshape(-1)
    w.bd2_mu, w.bd2_sigma = float(bd2[0]), float(bd2[1])
    w.inv_scale_x, w.inv_scale_h, w.inv_scale_p, w.inv_scale_w, w.inv_scale_wx, w.inv_scale_23 = isx, ish, isp, isw, iswx, is23
    w.inv_scale_m, w.inv_scale_se, w.inv_scale_wx2 = ism, isse, iswx2
    w.n_streams = w.n_heads = len(streams)
    w.n_weight_sets = len(sets)
    return t, w


_ACC_FIX_DONE = {}


def measure_acc_trunc_bias(device, one_signed=False, seed=12345, fine=False):
    """Relative shrink of the tensor-core accumulation, measured with a probe through spb_wino_gemm (fix = 0):
    24 GEMMs [128 x 512] x [512 x 128] on seeded operands, compared with the exact float64 value of what the
    kernel sums (hi*hi + (hi*lo + lo*hi) / 2^11 of the same fp16 pairs; numpy on the host: no GPU library call).
    Returns b with  device result ~ (1 - b) * exact.  `one_signed`: all operands >= 0 (the worst case for a
    truncating accumulator)."""
    import numpy as np
    lib = _lib.load()
    dev = torch.device(device)
    rng = np.random.default_rng(seed)
    rows = cols = 128
    u = rng.standard_normal((24, rows, 512)).astype(np.float32)
    w = (rng.standard_normal((24 * cols, 512)) * 0.05).astype(np.float32)
    if one_signed:
        u, w = np.abs(u), np.abs(w)
    mx = float(np.abs(w).max())
    scale = 2.0 ** (5 - math.floor(math.log2(mx)))

    def pair(x):
        hi = x.astype(np.float16)
        lo = ((x.astype(np.float64) - hi.astype(np.float64)) * 2048.0).astype(np.float16)
        return hi, lo
    u_hi, u_lo = pair(u)
    w_hi, w_lo = pair((w.astype(np.float64) * scale).astype(np.float32))
    f = lambda a: a.astype(np.float64)
    wh, wl = f(w_hi).reshape(24, cols, 512), f(w_lo).reshape(24, cols, 512)
    m = (np.einsum("prk,pck->prc", f(u_hi), wh)
         + (np.einsum("prk,pck->prc", f(u_hi), wl) + np.einsum("prk,pck->prc", f(u_lo), wh)) / 2048.0) / scale
    m = m.reshape(6, 4, rows, cols)                                           # [j][i]
    ref = np.stack([m[:, 0] + m[:, 1] + m[:, 2], m[:, 1] - m[:, 2] - m[:, 3]], 1).reshape(12, rows, cols)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = [t(x) for x in (u_hi, u_lo, w_hi, w_lo)]
    out = torch.empty((12, cols // 128, rows, 128), dtype=torch.float32, device=dev)
    getter, setter = (lib.spb_get_acc_trunc_fix_fine, lib.spb_set_acc_trunc_fix_fine) if fine else \
        (lib.spb_get_acc_trunc_fix, lib.spb_set_acc_trunc_fix)
    old = getter()
    _lib.check(setter(0.0), "spb_set_acc_trunc_fix")
    try:
        with torch.cuda.device(dev):
            _lib.check(lib.spb_wino_gemm(_lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), _lib.ptr(out),
                                         rows, cols, 1.0 / scale, 1 if fine else 0, _lib.current_stream()), "spb_wino_gemm")
        got = out.permute(0, 2, 1, 3).reshape(12, rows, cols).double().cpu().numpy()
    finally:
        setter(old)
    return 1.0 - float((got * ref).sum() / (ref * ref).sum())


def calibrate_acc_trunc_fix(device):
    """Installs the measured compensation factor (csrc/decoder.cuh) once per process and GPU model.
    SPB_ACC_TRUNC_FIX=default keeps the built-in 5.5e-7, SPB_ACC_TRUNC_FIX=<number> forces a value."""
    import os
    import warnings
    lib = _lib.load()
    key = torch.cuda.get_device_name(device)
    if key in _ACC_FIX_DONE:
        return _ACC_FIX_DONE[key]
    env = os.environ.get("SPB_ACC_TRUNC_FIX", "")
    if env == "default":
        fix = lib.spb_get_acc_trunc_fix()
    elif env:
        fix = float(env)
    else:
        b = measure_acc_trunc_bias(device)
        if not (0.0 <= b < 5e-6):
            warnings.warn("tensor-core accumulation probe measured a relative bias of %.3e (expected ~5.5e-7 on B200); "
                          "keeping the default compensation" % b)
            fix = lib.spb_get_acc_trunc_fix()
        else:
            fix = b
    _lib.check(lib.spb_set_acc_trunc_fix(fix), "spb_set_acc_trunc_fix")
    if not env:                                              # the 8-k-step accumulators of the fine-drain GEMM
        bf = measure_acc_trunc_bias(device, fine=True)
        if 0.0 <= bf < 5e-6:
            _lib.check(lib.spb_set_acc_trunc_fix_fine(bf), "spb_set_acc_trunc_fix_fine")
    _ACC_FIX_DONE[key] = fix
    return fix


class CudaDecoder:
    """Owns the prepared weights and the workspace of one device; runs spb_decode in waves."""

    def __init__(self, state_dict, task="OSIE", steps=16, device="cuda", wave=256, use_tensor_cores=True):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.task, self.steps, self.wave = task, int(steps), int(wave)
        self.device = torch.device(device)
        # 0 = SIMT check path (explicit 5x5 layer); 1 = product path: tcgen05, Winograd F(2x4) h-gates + Winograd
        # F(2x2) x-gates + composed head; 2 = tcgen05 direct 3x3 for both; 3 / 4 = F(2x4) for both; 5 = F(2x4)
        # h-gates + direct x-gates (see csrc/decode.cu)
        self.use_tensor_cores = int(use_tensor_cores)
        if self.use_tensor_cores:
            self.acc_trunc_fix = calibrate_acc_trunc_fix(self.device)
        self.tensors, self.w = prepare_weights(state_dict, task, self.device)
        self.heads = int(self.w.n_heads)
        self._ws, self._ws_n = None, 0
        self._graphs = {}

    def _workspace(self, n):
        if self._ws is None or self._ws_n < n:
            nbytes = self.lib.spb_decoder_workspace_bytes(n, self.w.n_streams, self.w.n_heads, self.steps)
            self._ws = torch.empty((nbytes + 1024,), dtype=torch.uint8, device=self.device)
            self._ws_n = n
        off = (-self._ws.data_ptr()) % 1024
        return self._ws.data_ptr() + off, self._ws.numel() - off

    def decode(self, visual_feature, attention_maps=None, tasks=None):
        """visual_feature [N,512,30,40] f32 (device).  Returns probs [HD,N,T,1201], mu, sigma2 [HD,N,T],
        action_map [HD,N,T,30,40] (HD = 2 for AiR: good, poor)."""
        vf = visual_feature.detach().to(self.device, torch.float32).contiguous()
        N, T, HD = vf.shape[0], self.steps, self.heads
        assert vf.shape[1:] == (E, 30, 40), "visual_feature must be [N,512,30,40]"
        dev = self.device
        probs = torch.empty((HD, N, T, A), dtype=torch.float32, device=dev)
        mu = torch.empty((HD, N, T), dtype=torch.float32, device=dev)
        s2 = torch.empty((HD, N, T), dtype=torch.float32, device=dev)
        amap = torch.empty((HD, N, T, HW), dtype=torch.float32, device=dev)
        att = None
        if attention_maps is not None and self.task != "OSIE":
            att = attention_maps.detach().to(dev, torch.float32).reshape(N, HW).contiguous()
        rows = None
        if self.task == "COCO_Search18":
            assert tasks is not None, "COCO-Search18 decoding needs the task ids"
            rows = (torch.as_tensor(tasks).to(dev).to(torch.int32) * E).contiguous()
        for n0 in range(0, N, self.wave):
            n1 = min(N, n0 + self.wave)
            n = n1 - n0
            ws_ptr, ws_bytes = self._workspace(n if N <= self.wave else self.wave)
            # per-wave outputs are strided views of [HD, N, ...]; the kernels want dense [HD, n, ...]
            dense = (n == N)
            p_w = probs if dense else torch.empty((HD, n, T, A), dtype=torch.float32, device=dev)
            m_w = mu if dense else torch.empty((HD, n, T), dtype=torch.float32, device=dev)
            s_w = s2 if dense else torch.empty((HD, n, T), dtype=torch.float32, device=dev)
            a_w = amap if dense else torch.empty((HD, n, T, HW), dtype=torch.float32, device=dev)
            io = DecoderIO(n, T, self.use_tensor_cores, 0, vf[n0:n1].data_ptr(),
                           att[n0:n1].data_ptr() if att is not None else None,
                           rows[n0:n1].data_ptr() if rows is not None else None, ws_ptr, ws_bytes, p_w.data_ptr(),
                           m_w.data_ptr(), s_w.data_ptr(), a_w.data_ptr())
            with torch.cuda.device(dev):
                _lib.check(self.lib.spb_decode(C.byref(self.w), C.byref(io), _lib.current_stream()), "spb_decode")
            if not dense:
                probs[:, n0:n1], mu[:, n0:n1], s2[:, n0:n1], amap[:, n0:n1] = p_w, m_w, s_w, a_w
        return probs, mu, s2, amap.view(HD, N, T, 30, 40)


def _decode_graphed(self, visual_feature, attention_maps=None, tasks=None):
    """decode() replayed from a CUDA graph: the ~230 launches of a 16-step rollout (each with its tensor-map
    encodes) become one cudaGraphLaunch -- what matters when the wave is small and the rollout is launch-bound
    (the SCST step: 4 images).  Inputs are copied into static buffers; the returned tensors are the graph's
    static outputs and are overwritten by the next call with the same shape."""
    vf = visual_feature.detach().to(self.device, torch.float32)
    N = vf.shape[0]
    assert N <= self.wave, "graphed decode handles one wave"
    key = (N, attention_maps is not None and self.task != "OSIE", tasks is not None)
    g = self._graphs.get(key)
    if g is None:
        st = {"vf": torch.empty_like(vf).contiguous(),
              "att": torch.empty((N, 1, 30, 40), dtype=torch.float32, device=self.device) if key[1] else None,
              "tasks": torch.empty((N,), dtype=torch.int64, device=self.device) if key[2] else None}
        st["vf"].copy_(vf)
        if key[1]:
            st["att"].copy_(attention_maps.reshape(N, 1, 30, 40))
        if key[2]:
            st["tasks"].copy_(torch.as_tensor(tasks).to(self.device))
        self.decode(st["vf"], st["att"], st["tasks"])                     # warm-up: workspace, attributes, modules
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["out"] = self.decode(st["vf"], st["att"], st["tasks"])
        g = self._graphs[key] = (graph, st)
    graph, st = g
    st["vf"].copy_(vf, non_blocking=True)
    if key[1]:
        st["att"].copy_(attention_maps.reshape(N, 1, 30, 40), non_blocking=True)
    if key[2]:
        st["tasks"].copy_(torch.as_tensor(tasks).to(self.device), non_blocking=True)
    graph.replay()
    return st["out"]


CudaDecoder.decode_graphed = _decode_graphed


def _result_dict(task, probs, mu, s2, amap):
    if task == "AiR":
        out = {}
        for i, pre in enumerate(("good_", "poor_")):
            out[pre + "all_actions_prob"], out[pre + "log_normal_mu"] = probs[i], mu[i]
            out[pre + "log_normal_sigma2"], out[pre + "action_map"] = s2[i], amap[i]
        return out
    return {"all_actions_prob": probs[0], "log_normal_mu": mu[0], "log_normal_sigma2": s2[0], "action_map": amap[0]}


class _ConvLSTMParams(nn.Module):
    def __init__(self, task):
        super().__init__()
        names = list(GATES_X) + list(GATES_H)
        names += ["input_pos", "forget_pos", "output_pos", "input_neg", "forget_neg", "output_neg"] \
            if task == "AiR" else ["input", "forget", "output"]
        for n in names:
            setattr(self, n, nn.Conv2d(E, E, kernel_size=3, padding=1, stride=1, bias=True))


class _SemanticAtt(nn.Module):
    def __init__(self):
        super().__init__()
        self.semantic_lists = nn.Linear(E, E, bias=True)
        self.semantic_cur = nn.Linear(E, E, bias=True)
        self.semantic_attention = nn.Linear(E, 1, bias=True)


class _SpatialAtt(nn.Module):
    def __init__(self):
        super().__init__()
        self.spatial_lists = nn.Conv2d(1, 1, kernel_size=3, padding=1, stride=1, bias=True)
        self.spatial_cur = nn.Conv2d(1, 1, kernel_size=3, padding=1, stride=1, bias=True)
        self.spatial_attention = nn.Conv2d(1, 1, kernel_size=(30, 40), padding=0, stride=1, bias=True)


class _PredictHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.sal_layer_2 = nn.Conv2d(512, 1, kernel_size=1, padding=0, stride=1, bias=True)
        self.sal_layer_3 = nn.Conv2d(512, 1, kernel_size=1, padding=0, stride=1, bias=True)
        self.drt_layer_1 = nn.Conv2d(512, 1, kernel_size=7, padding=2, stride=5, bias=True)
        self.drt_layer_2 = nn.Conv2d(1, 2, kernel_size=(6, 8), padding=0, stride=1, bias=True)


class baseline(nn.Module):
    """Same parameters (names, shapes) as the reference ``baseline``; ``task`` selects which of the
    three reference variants is mirrored.  The encoder is attached lazily by ``attach_encoder``."""

    def __init__(self, embed_size=512, convLSTM_length=16, min_length=1, ratio=4, map_width=40, map_height=30,
                 projected_label_length=18, task="OSIE", wave=256, use_tensor_cores=True):
        super().__init__()
        assert embed_size == 512 and map_width == 40 and map_height == 30, "the reference hard-codes 512 x 30 x 40"
        self.task = task
        self.embed_size, self.convLSTM_length, self.min_length = embed_size, convLSTM_length, min_length
        self.map_width, self.map_height = map_width, map_height
        self.wave, self.use_tensor_cores = wave, use_tensor_cores
        self.lstm = _ConvLSTMParams(task)
        self.semantic_embed = nn.Linear(512, embed_size)
        self.spatial_embed = nn.Linear(1200, 1200, bias=True)
        self.semantic_att = _SemanticAtt()
        self.spatial_att = _SpatialAtt()
        conv5 = lambda: nn.Conv2d(512, 512, kernel_size=5, padding=2, stride=1, bias=True)
        if task == "AiR":
            self.performance_sal_layer = nn.ModuleDict({"False": conv5(), "True": conv5()})
        elif task == "COCO_Search18":
            self.object_sal_layer = nn.ModuleDict({o: conv5() for o in COCO_OBJECTS})
        else:
            self.performance_sal_layer = conv5()
        self.object_head = _PredictHead()
        self.resnet = None          # encoder: PyTorch, out of scope of the CUDA path
        self.sal_conv = None
        self._decoder = None
        self.eval()

    def attach_encoder(self):
        """Adds the once-per-image encoder (stays in PyTorch, out of scope of the CUDA path):
        ResNet-50 with the stride on conv1 of each stage's first block and a ceil-mode max-pool
        (models/resnet.py:55-103), dilated as in dilate_resnet (baseline_attention.py:212-224),
        children()[:-2], then sal_conv 3x3 2048 -> 512 (:191-194).  Parameter names match the
        reference's `resnet.*` / `sal_conv.*` keys."""
        import torchvision
        net = torchvision.models.resnet50(weights=None)
        net.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=0, ceil_mode=True)
        for layer in (net.layer2, net.layer3, net.layer4):      # torchvision strides conv2, the reference conv1
            layer[0].conv1.stride, layer[0].conv2.stride = layer[0].conv2.stride, (1, 1)
        for layer in (net.layer2, net.layer4):
            layer[0].conv1.stride = (1, 1)
            layer[0].downsample[0].stride = (1, 1)
        for block in net.layer3:
            block.conv2.dilation, block.conv2.padding = (2, 2), (2, 2)
        for block in net.layer4:
            block.conv2.dilation, block.conv2.padding = (4, 4), (4, 4)
        self.resnet = nn.Sequential(*list(net.children())[:-2])
        self.sal_conv = nn.Conv2d(2048, 512, kernel_size=3, padding=1, stride=1, bias=True)
        dev = next(self.lstm.parameters()).device
        self.resnet.to(dev); self.sal_conv.to(dev)
        self.eval()
        return self

    def encode(self, images):
        """The once-per-image encoder (baseline_attention.py:327-328): the dilated ResNet-50 stays in PyTorch;
        its last layer, relu(sal_conv(.)) -- 22.65 GFLOP per image, as much as the whole x-gate convolution --
        runs on the tensor cores through spb_sal_conv when the model lives on a CUDA device."""
        x = self.resnet(images)
        if not x.is_cuda:
            return F.relu(self.sal_conv(x))                # CPU tensors: plain PyTorch (tests/test_encoder_reference.py)
        return self.sal_conv_cuda(x)

    def sal_conv_cuda(self, x):
        """relu(sal_conv(x)) for x [N,2048,30,40] f32 on the device -> [N,512,30,40] f32 (csrc/conv_tc.cu)."""
        lib = _lib.load()
        x = x.detach().to(torch.float32).contiguous()
        N = x.shape[0]
        assert x.shape[1:] == (2048, 30, 40), "sal_conv input must be [N,2048,30,40]"
        dev = x.device
        if getattr(self, "_sal_prepared", None) is None or self._sal_prepared[0] != dev:
            if self.use_tensor_cores:
                calibrate_acc_trunc_fix(dev)
            w = self.sal_conv.weight.detach().to("cpu", torch.float32)
            hi, lo, inv = split_pair(_conv_to_gemm(w))     # [512, 9*2048], K index (ky*3+kx)*2048 + ci
            self._sal_prepared = (dev, hi.to(dev), lo.to(dev), self.sal_conv.bias.detach().to(dev, torch.float32).contiguous(), inv)
        _, w_hi, w_lo, bias, inv = self._sal_prepared
        out = torch.empty((N, E, 30, 40), dtype=torch.float32, device=dev)
        for n0 in range(0, N, self.wave):                  # waves bound the operand workspace (9.8 MB per image)
            n = min(N, n0 + self.wave) - n0
            nbytes = lib.spb_sal_conv_workspace_bytes(n)
            if getattr(self, "_sal_ws", None) is None or self._sal_ws.numel() < nbytes + 1024 or self._sal_ws.device != dev:
                self._sal_ws = torch.empty((nbytes + 1024,), dtype=torch.uint8, device=dev)
            off = (-self._sal_ws.data_ptr()) % 1024
            with torch.cuda.device(dev):
                _lib.check(lib.spb_sal_conv(x[n0:n0 + n].data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), bias.data_ptr(), inv, n,
                                            self._sal_ws.data_ptr() + off, self._sal_ws.numel() - off,
                                            out[n0:n0 + n].data_ptr(), _lib.current_stream()), "spb_sal_conv")
        return out

    def load_state_dict(self, state_dict, strict=False, **kw):
        self._decoder = None
        self._sal_prepared = None
        own = {k: v for k, v in state_dict.items() if not k.startswith(("resnet.", "sal_conv."))} \
            if self.resnet is None else state_dict
        return super().load_state_dict(own, strict=strict, **kw)

    def decoder(self):
        if self._decoder is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise _lib.SpbError("scanpaths_b200 has no CPU path: move the model to a CUDA device")
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith(("resnet.", "sal_conv."))}
            self._decoder = CudaDecoder(sd, self.task, self.convLSTM_length, dev, self.wave, self.use_tensor_cores)
        return self._decoder

    def decode(self, visual_feature, attention_maps=None, tasks=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("training_process stays in PyTorch (out of scope); call under eval()")
        return _result_dict(self.task, *self.decoder().decode(visual_feature, attention_maps, tasks))

    def forward(self, images, attention_maps=None, tasks=None, performances=None):
        """OSIE: forward(images); COCO-Search18: forward(images, attention_maps, tasks); AiR:
        forward(images, attention_maps, performances=None) -- the third positional of the AiR reference is
        `performances` (AiR/models/baseline_attention.py:253), which only its training branch reads."""
        if self.task == "AiR":
            tasks = None                                   # a positional `performances` lands here: inference ignores it
        if self.training:
            raise NotImplementedError("training_process (with gradients) is out of scope of the CUDA path")
        if images.shape[1] == E:                       # already visual_feature [N,512,30,40]
            vf = images
        else:
            if self.resnet is None:
                raise _lib.SpbError("no encoder attached: pass visual_feature [N,512,30,40] or attach_encoder()")
            vf = self.encode(images)
        return self.decode(vf, attention_maps, tasks)


def smoke_decode(n, device):
    """Tiny decode used by __graft_entry__.smoke(): random-init OSIE weights, synthetic features."""
    from ..weights import random_state_dict, synthetic_features
    dec = CudaDecoder(random_state_dict("OSIE", 0), "OSIE", 16, device, wave=n)
    probs, mu, s2, _ = dec.decode(synthetic_features(n, 0).to(device))
    return probs[0], mu[0], s2[0]
