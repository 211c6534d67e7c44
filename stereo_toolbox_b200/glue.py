"""Inference-only shadow of a drop-in model with every eval-mode BatchNorm2d folded into the Conv2d that feeds it.

The iterative models' 2-D networks (MobileNetV2 trunk, context / feature encoders, stems, upsampling heads) are torch glue
outside the hot path, but at BASELINE config 5 cuDNN's NCHW BatchNorm-inference passes alone are 21 ms of IGEV-Stereo's
forward (tools/model_profile.py) -- pure memory traffic that a folded convolution does not have.  The fold is done on a
deep copy kept beside the model (never registered as a submodule, so the state dict and the parameters a user sees are
untouched) and rebuilt whenever a parameter or running statistic changes.

Which BatchNorm follows which convolution is not read off attribute names: one small recording forward runs on the copy
with hooks, and a BatchNorm2d is folded exactly when its input IS the output tensor of a Conv2d / ConvTranspose2d (object
identity) and that pairing is one-to-one over the whole forward.  InstanceNorm / GroupNorm and all 3-D modules are left alone.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
from torch.nn.utils.fusion import fuse_conv_bn_eval

_CACHE_KEYS = ("_graph_cache", "_umma_update", "_fold_shadow", "_fe_umma")


def _state_version(model: nn.Module):
    return tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers()))


def _fold(shadow: nn.Module, record) -> int:
    produced, feeds = {}, {}
    hooks = []

    def conv_hook(mod, inp, out):
        produced[id(out)] = (mod, out)                  # the tensor is kept alive: its id cannot be reused during the recording

    def bn_hook(mod, inp):
        src = produced.get(id(inp[0]))
        feeds.setdefault(mod, set()).add(None if src is None else src[0])

    for m in shadow.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.BatchNorm2d):
            hooks.append(m.register_forward_pre_hook(bn_hook))
    try:
        with torch.no_grad():
            record(shadow)
    finally:
        for h in hooks:
            h.remove()
        produced.clear()
    conv_use = {}
    for bn, srcs in feeds.items():
        for c in srcs:
            conv_use.setdefault(c, set()).add(bn)
    folded = set()
    for bn, srcs in feeds.items():
        if len(srcs) != 1:
            continue
        conv = next(iter(srcs))
        if conv is None or len(conv_use[conv]) != 1 or bn.running_mean is None or bn.training:
            continue
        fused = fuse_conv_bn_eval(conv.eval(), bn, transpose=isinstance(conv, nn.ConvTranspose2d))
        conv.weight = fused.weight
        conv.bias = fused.bias
        folded.add(bn)
    for parent in shadow.modules():
        for name, child in list(parent.named_children()):
            if child in folded:
                setattr(parent, name, nn.Identity())
    return len(folded)


def inference_shadow(model: nn.Module, record):
    """The folded copy of ``model`` (cached in ``model.__dict__``; rebuilt when the model's state changes).
    ``record(shadow)`` must run one small forward of the copy."""
    ver = _state_version(model)
    hit = model.__dict__.get("_fold_shadow")
    if hit is not None and hit[0] == ver:
        return hit[1]
    stash = {k: model.__dict__.pop(k) for k in _CACHE_KEYS if k in model.__dict__}
    try:
        shadow = copy.deepcopy(model)
    finally:
        model.__dict__.update({k: v for k, v in stash.items() if k != "_fold_shadow"})
    shadow.eval()
    shadow.__dict__["fold_bn"] = False                   # the copy runs its own forward
    if hasattr(shadow, "_be") and hasattr(shadow, "precision"):
        # kernel-plan caches are keyed by module identity: start empty (make_backend as the model's own module sees it)
        import sys
        mk = next((getattr(sys.modules[c.__module__], "make_backend") for c in type(shadow).__mro__
                   if hasattr(sys.modules.get(c.__module__), "make_backend")), None)
        if mk is not None:
            shadow._be = mk(shadow.precision)
    shadow.__dict__["_folded_bn_count"] = _fold(shadow, record)
    model.__dict__["_fold_shadow"] = (ver, shadow)
    return shadow
