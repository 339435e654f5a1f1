"""Import the UNMODIFIED reference (chenxy99/Scanpaths) from /root/reference.

Only usable where the reference is mounted (the authoring container); the GPU
box never has it, so nothing marked ``gpu`` may call this.  The reference is
three script trees with clashing top-level module names (``models``, ``utils``)
and a few imports that are absent here; we add import-only stubs:

  * ``matplotlib.pyplot``   (visual_attention_metrics.py:19, never called)
  * ``multimatch_gaze``     (utils/evaluation.py:7) -> ``docomparison`` returning
    NaN x5 when either scanpath has < 3 fixations (the upstream rule, see
    SURVEY.md section 8c) else zeros; MultiMatch values themselves are out of scope
  * ``mmcv.cnn``            (baseline_attention.py:9) -> the four init helpers
  * ``Tensor.get_device``   shim so ``Sampling.random_sample`` runs on CPU
    (sampling.py:26 does ``.to(t.get_device())`` which is -1 on CPU)
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SCANPATHS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "OSIE", "utils", "evaltools"))


def _install_stubs():
    import torch
    import torch.nn as nn

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt

    if "multimatch_gaze" not in sys.modules:
        mm = types.ModuleType("multimatch_gaze")

        def docomparison(fix1, fix2, screensize=None, **kw):
            if len(fix1) < 3 or len(fix2) < 3:
                return [np.nan] * 5
            return [0.0] * 5

        mm.docomparison = docomparison
        sys.modules["multimatch_gaze"] = mm

    if "mmcv" not in sys.modules:
        mmcv = types.ModuleType("mmcv")
        cnn = types.ModuleType("mmcv.cnn")

        def xavier_init(module, gain=1, bias=0, distribution="normal"):
            if getattr(module, "weight", None) is not None:
                if distribution == "uniform":
                    nn.init.xavier_uniform_(module.weight, gain=gain)
                else:
                    nn.init.xavier_normal_(module.weight, gain=gain)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        def constant_init(module, val, bias=0):
            if getattr(module, "weight", None) is not None:
                nn.init.constant_(module.weight, val)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        def normal_init(module, mean=0, std=1, bias=0):
            if getattr(module, "weight", None) is not None:
                nn.init.normal_(module.weight, mean, std)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
            if getattr(module, "weight", None) is not None:
                nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        cnn.xavier_init, cnn.constant_init = xavier_init, constant_init
        cnn.normal_init, cnn.kaiming_init = normal_init, kaiming_init
        mmcv.cnn = cnn
        sys.modules["mmcv"] = mmcv
        sys.modules["mmcv.cnn"] = cnn

    if not getattr(torch.Tensor, "_spb_get_device_shim", False):
        orig = torch.Tensor.get_device

        def get_device(self):
            d = orig(self)
            return self.device if d < 0 else d

        torch.Tensor.get_device = get_device
        torch.Tensor._spb_get_device_shim = True


_CLASH = ("models", "utils", "dataset", "opts")


def load_reference(task: str = "OSIE"):
    """Returns a namespace with the reference modules of one task tree."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    _install_stubs()
    for name in list(sys.modules):
        if name.split(".")[0] in _CLASH:
            del sys.modules[name]
    tree = os.path.join(REF_ROOT, task)
    sys.path[:] = [p for p in sys.path if not p.startswith(REF_ROOT)]
    sys.path.insert(0, tree)
    ns = types.SimpleNamespace()
    ns.scanmatch = importlib.import_module("utils.evaltools.scanmatch")
    ns.vame = importlib.import_module("utils.evaltools.visual_attention_metrics")
    ns.evaluation = importlib.import_module("utils.evaluation")
    ns.sampling = importlib.import_module("models.sampling")
    ns.loss = importlib.import_module("models.loss")
    ns.tree = tree
    return ns


def load_reference_model(task: str = "OSIE"):
    """Imports the task's model module with the pretrained-ResNet download
    neutralised (resnet.py:187 calls model_zoo.load_url; no network here)."""
    ns = load_reference(task)
    import torch.utils.model_zoo as model_zoo

    model_zoo.load_url = lambda *a, **k: {}
    resnet = importlib.import_module("models.resnet")
    import torch.nn as nn

    orig_load = nn.Module.load_state_dict

    def _lenient(self, sd, *a, **k):
        if len(sd) == 0:
            return None
        return orig_load(self, sd, *a, **k)

    resnet.ResNet.load_state_dict = _lenient
    modname = {"OSIE": "models.baseline_attention", "AiR": "models.baseline_attention",
               "COCO_Search18": "models.baseline_attention_multihead"}[task]
    ns.model = importlib.import_module(modname)
    return ns
