"""Test infrastructure: run the drop-in models' HOST code on CPU with every hot-path call answered by the oracle.

The product has no CPU path (ops raise ``StbError`` on CPU tensors).  To pin the host-side mirrors of the reference --
constructors, state-dict layouts, 2-D glue, samplers / warps, layer order, residual wiring, return conventions --
without a GPU, ``oracle_hot_path`` swaps, for the duration of one test,

* ``make_backend`` in every model module for ``OracleBackend`` (the ``aggregation.Fp32Backend`` interface, each method
  answered by ``oracle/ref_ops.py``), and
* the handful of ``stereo_toolbox_b200.ops`` functions that model code calls directly.

Each swapped model is then compared with the fixture the REFERENCE produced (tests/golden/*.npz), so these tests pin
mirror + oracle together; the CUDA kernels behind the same calls are pinned by the ``-m gpu`` tests.
"""
import contextlib
import importlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import ref_ops as R


def _split(layer):
    if isinstance(layer, nn.Sequential):
        bn = layer[1] if len(layer) > 1 and isinstance(layer[1], nn.modules.batchnorm._BatchNorm) else None
        return layer[0], bn
    return layer, None


class _NoProf:
    enabled = False

    def bracket(self, *a, **k):
        return contextlib.nullcontext()


class OracleBackend:
    name = "oracle"

    def __init__(self):
        self.prof = _NoProf()

    # --- volumes
    def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
        vol = R.build_gwc_volume(gwc_l, gwc_r, maxdisp4, groups)
        if cat_l is None:
            return vol
        return torch.cat((vol, R.build_concat_volume(cat_l, cat_r, maxdisp4, True)), 1)

    def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
        vol = R.build_concat_volume(l, r, maxdisp4, mask_left)
        return vol if att_prob is None else vol * att_prob

    # --- conv family
    def conv(self, layer, x, act="none", residual=None):
        conv, bn = _split(layer)
        tr = isinstance(conv, nn.ConvTranspose3d)
        bnd = None if bn is None else dict(weight=bn.weight, bias=bn.bias, running_mean=bn.running_mean,
                                           running_var=bn.running_var)
        if conv.bias is None:
            return R.conv3d_bn_act(x, conv.weight, bnd, conv.stride[0], conv.padding[0], act, residual, tr,
                                   conv.output_padding[0] if tr else 0)
        assert bn is None and not tr                      # biased convs on the path: ACVNet qkv / final1x1
        y = F.conv3d(x, conv.weight, conv.bias, conv.stride, conv.padding)
        return R.activation(y if residual is None else y + residual, act)

    def block_attention(self, qkv, bias, heads, block):
        return R.block_attention_core(qkv, bias, heads, block)

    def gate(self, x, gate_logits):
        return x * torch.sigmoid(gate_logits).unsqueeze(2)

    def cat(self, xs):
        return torch.cat(list(xs), dim=1)

    # --- layout boundary: the oracle works in the reference's NCDHW fp32 layout throughout
    def from_ncdhw(self, x):
        return x

    def to_ncdhw(self, x, channels=None):
        return x

    def cost_ncdhw(self, cost):
        return cost

    def cost_native(self, cost):
        return cost

    def head(self, cost, maxdisp, H, W, align_corners=False):
        return R.upsample_softargmin(cost, maxdisp, H, W, align_corners, False)


class OracleTrainBackend(OracleBackend):
    """aggregation.TrainBackend's semantics (raw conv -> the layer's own BatchNorm3d module, i.e. batch statistics in
    train mode -> residual -> activation) with torch's convolution instead of the CUDA kernels."""
    name = "oracle-train"
    training_path = True

    def conv(self, layer, x, act="none", residual=None):
        conv, bn = _split(layer)
        y = conv(x)
        if bn is not None:
            y = bn(y)
        return R.activation(y if residual is None else y + residual, act)


def _patch_dw(x, weight, dilation, out=None, c_off=0):
    C = weight.shape[0]
    y = R.depthwise_patch(x[:, c_off:c_off + C], weight, dilation)
    if out is None:
        assert C == x.shape[1]
        return y
    out[:, c_off:c_off + C] = y
    return out


def _concat_volume(left, right, maxdisp, mask_left=True, att_prob=None, out=None, c_off=0):
    assert out is None
    vol = R.build_concat_volume(left, right, maxdisp, mask_left)
    return vol if att_prob is None else vol * att_prob


def _softmax_d(x):
    return torch.softmax(x, dim=2 if x.dim() == 5 else 1)


_OPS = dict(
    gwc_volume=lambda l, r, d, g, out=None, c_off=0: R.build_gwc_volume(l, r, d, g),
    concat_volume=_concat_volume,
    patch_dw=_patch_dw,
    softmax_d=_softmax_d,
    disparity_regression=lambda prob, maxdisp, keepdim=False: R.disparity_regression(prob, maxdisp, keepdim),
    disparity_variance=lambda prob, maxdisp, disp: R.disparity_variance(prob, maxdisp, disp.view(prob.shape[0], 1, *prob.shape[2:])),
    upsample_softargmin=lambda cost, maxdisp, H, W, ac=False: R.upsample_softargmin(cost, maxdisp, H, W, ac, False),
    # iterative models: functional.CorrBlock1D / Combined_Geo_Encoding_Volume keep their own host code
    corr1d=lambda f1, f2, scale=True: R.corr1d(f1, f2, scale),
    avgpool_last=R.avg_pool_last,
    corr1d_lookup=lambda pyr, coords, radius, levels: R.corr_lookup(pyr, coords[:, 0], radius, levels),
    geo_permute=lambda geo: geo.permute(0, 3, 4, 1, 2).contiguous(),
    geo_lookup=lambda geos, corrs, disp, coords, radius: R.geo_lookup(
        geos, corrs, disp.reshape(geos[0].shape[:3]), coords.reshape(geos[0].shape[:3]), radius),
)

_MODEL_MODULES = ("gwcnet", "psmnet", "acvnet", "cfnet", "pcwnet", "igev")


@contextlib.contextmanager
def oracle_hot_path():
    """Swap the hot path for the oracle (see module docstring).  Everything is restored on exit."""
    from stereo_toolbox_b200 import ops
    undo = []
    for name in _MODEL_MODULES:
        mod = importlib.import_module(f"stereo_toolbox_b200.{name}")
        if hasattr(mod, "make_backend"):
            undo.append((mod, "make_backend", mod.make_backend))
            mod.make_backend = lambda precision: OracleBackend()
        if hasattr(mod, "TrainBackend"):
            undo.append((mod, "TrainBackend", mod.TrainBackend))
            mod.TrainBackend = OracleTrainBackend
    for name, fn in _OPS.items():
        undo.append((ops, name, getattr(ops, name)))
        setattr(ops, name, fn)
    try:
        yield
    finally:
        for obj, name, old in undo:
            setattr(obj, name, old)
