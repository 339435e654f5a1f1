"""CPU oracle: the 16-step ConvLSTM rollout + prediction head (TEST INFRASTRUCTURE ONLY).

Functional torch-CPU restatement (any dtype; the parity truth is float64) of
``baseline.inference`` from ``state = init_hidden`` on, i.e. everything after
the once-per-image encoder, operating on a state_dict with the reference's key
names.  Follows, per task tree under /root/reference:

  OSIE   models/baseline_attention.py            ConvLSTM :33-48, semantic_att :69-80,
         spatial_att :103-116, predict_head :141-166, feedback :226-236, driver :323-396
  AiR    models/baseline_attention.py            pos/neg streams :37-56, two heads :423-429
  COCO   models/baseline_attention_multihead.py  per-task 5x5 layer :359-362

Pinned by tests/golden/decoder_*.npz: outputs of the reference modules themselves
(float64 and float32) on seeded weights / features.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

COCO_OBJECTS = ["bottle", "bowl", "car", "chair", "clock", "cup", "fork", "keyboard", "knife",
                "laptop", "microwave", "mouse", "oven", "potted plant", "sink", "stop sign",
                "toilet", "tv"]


def _conv(sd, name, x, padding=0, stride=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _convlstm(sd, x, h, c, streams):
    """streams: list of (suffix, spatial [N,30,40], semantic [N,512]); OSIE/COCO use
    one stream with suffix '' (lstm.input/forget/output), AiR two ('_pos', '_neg')."""
    pre = {g: _conv(sd, "lstm.%s_x" % g, x, 1) + _conv(sd, "lstm.%s_h" % g, h, 1)
           for g in ("input", "forget", "output", "memory")}
    for suffix, spatial, semantic in streams:
        ss = spatial.unsqueeze(1) * semantic.unsqueeze(-1).unsqueeze(-1)
        for g in ("input", "forget", "output"):
            pre[g] = pre[g] + _conv(sd, "lstm.%s%s" % (g, suffix), ss, 1)
    i, f, o = torch.sigmoid(pre["input"]), torch.sigmoid(pre["forget"]), torch.sigmoid(pre["output"])
    g = torch.tanh(pre["memory"])
    c = f * c + i * g
    return o * c, c                                   # h = o * c (no tanh), :45


def _head(sd, feat):
    n = feat.shape[0]
    stop = _conv(sd, "object_head.sal_layer_2", feat).squeeze(1).mean(dim=(1, 2)).view(n, 1, 1)
    t = _conv(sd, "object_head.drt_layer_2", F.relu(_conv(sd, "object_head.drt_layer_1", feat, 2, 5)))
    mu = t[:, 0].reshape(n, -1)
    sigma2 = torch.exp(t[:, 1]).reshape(n, -1)
    amap = F.relu(_conv(sd, "object_head.sal_layer_3", feat))
    z = F.softmax(torch.cat([stop, amap.view(n, 1, -1)], dim=-1), -1)      # eval mode
    return z, mu, sigma2, amap


def _spatial_att(sd, lists, cur):
    n, t, hh, ww = lists.shape
    a = _conv(sd, "spatial_att.spatial_lists", lists.reshape(-1, 1, hh, ww), 1).view(n, t, hh, ww)
    b = _conv(sd, "spatial_att.spatial_cur", cur, 1)
    score = _conv(sd, "spatial_att.spatial_attention", (a + b).reshape(-1, 1, hh, ww)).view(n, t, 1, 1)
    return (lists * F.softmax(score, 1)).sum(1)


def _semantic_att(sd, lists, cur):
    a = _lin(sd, "semantic_att.semantic_lists", lists)
    b = _lin(sd, "semantic_att.semantic_cur", cur)
    w = F.softmax(_lin(sd, "semantic_att.semantic_attention", a + b.unsqueeze(1)), 1)
    return (lists * w).sum(1)


def _feedback(sd, amap, vf):
    n = vf.shape[0]
    prod = amap.expand_as(vf) * vf
    sp = F.relu(prod.mean(1, keepdim=True))
    sp = _lin(sd, "spatial_embed", sp.view(n, 1, -1)).view(n, 1, vf.shape[2], vf.shape[3])
    se = _lin(sd, "semantic_embed", F.relu(prod.view(n, vf.shape[1], -1).mean(-1)))
    return sp, se


def decode(sd, visual_feature, task="OSIE", attention_maps=None, tasks=None, steps=16):
    """Returns a dict like ``baseline.inference``: for OSIE/COCO the keys
    all_actions_prob [N,T,1201], log_normal_mu [N,T], log_normal_sigma2 [N,T],
    action_map [N,T,30,40]; for AiR the same with good_/poor_ prefixes."""
    vf = visual_feature
    n = vf.shape[0]
    sd = {k: v.to(vf.dtype) for k, v in sd.items()}
    if task == "OSIE" or attention_maps is None:
        attention_maps = vf.new_zeros((n, 1, vf.shape[2], vf.shape[3]))
    n_streams = 2 if task == "AiR" else 1
    suffixes = ["_pos", "_neg"] if task == "AiR" else [""]
    heads = ["True", "False"] if task == "AiR" else [None]

    sp_lists, se_lists, sp_mem, se_mem = [], [], [], []
    for _ in range(n_streams):
        sp, se = _feedback(sd, attention_maps, vf)
        sp_lists.append([sp]); se_lists.append([se])
        sp_mem.append(_spatial_att(sd, sp, sp))
        se_mem.append(_semantic_att(sd, se.unsqueeze(1), se))
    h, c = torch.zeros_like(vf), torch.zeros_like(vf)
    outs = [[] for _ in heads]
    for _ in range(steps):
        h, c = _convlstm(sd, vf, h, c, [(suffixes[s], sp_mem[s], se_mem[s]) for s in range(n_streams)])
        for hi, head in enumerate(heads):
            if task == "AiR":
                feat = _conv(sd, "performance_sal_layer.%s" % head, h, 2)
            elif task == "COCO_Search18":
                feat = torch.cat([_conv(sd, "object_sal_layer.%s" % COCO_OBJECTS[int(tasks[i])], h[i:i + 1], 2)
                                  for i in range(n)], 0)
            else:
                feat = _conv(sd, "performance_sal_layer", h, 2)
            z, mu, s2, amap = _head(sd, feat)
            outs[hi].append((z, mu, s2, amap))
            s = hi if task == "AiR" else 0           # good head feeds the pos stream, poor the neg
            sp, se = _feedback(sd, amap, vf)
            sp_lists[s].append(sp); se_lists[s].append(se)
            sp_mem[s] = _spatial_att(sd, torch.cat(sp_lists[s], 1), sp)
            se_mem[s] = _semantic_att(sd, torch.stack(se_lists[s], 1), se)
    res = {}
    for hi, head in enumerate(heads):
        pre = "" if head is None else ("good_" if head == "True" else "poor_")
        res[pre + "all_actions_prob"] = torch.cat([o[0] for o in outs[hi]], 1)
        res[pre + "log_normal_mu"] = torch.cat([o[1] for o in outs[hi]], 1)
        res[pre + "log_normal_sigma2"] = torch.cat([o[2] for o in outs[hi]], 1)
        res[pre + "action_map"] = torch.cat([o[3] for o in outs[hi]], 1)
    return res
