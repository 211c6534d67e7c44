"""Host side of the 3-D aggregation hot path: parameter containers with the reference's names and
the backends that run them through libstb200.so.

A *backend* owns the activation layout between kernels so model code never touches it:
  Fp32Backend  : NCDHW fp32, CUDA-core tap-list convolution (bit-faithful path, <=1e-3 px EPE)
  TrainBackend : the same, differentiable (autograd.py), batch-statistic BatchNorm
  UmmaBackend  : NDHWC fp16/bf16, tcgen05 implicit-GEMM convolution (aggregation_umma.py, conv3d_umma.cu)
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops


def convbn_3d(cin, cout, k, stride, pad):
    """Parameter container only: Conv3d(bias=False)+BatchNorm3d (PSMNet/submodule.py:16-19)."""
    return nn.Sequential(nn.Conv3d(cin, cout, k, stride, pad, bias=False), nn.BatchNorm3d(cout))


def deconvbn_3d(cin, cout):
    """ConvTranspose3d(k3,s2,p1,op1,bias=False)+BatchNorm3d (PSMNet/stackhourglass.py:25-29)."""
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
                         nn.BatchNorm3d(cout))


def _split(layer) -> Tuple[nn.Module, Optional[nn.Module]]:
    if isinstance(layer, nn.Sequential):
        conv = layer[0]
        bn = layer[1] if len(layer) > 1 and isinstance(layer[1], nn.modules.batchnorm._BatchNorm) else None
        return conv, bn
    return layer, None


def _bias_ver(conv):
    b = getattr(conv, "bias", None)
    return () if b is None else (b.data_ptr(), b._version)


class _NoProf:
    enabled = False

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def bracket(self, *a, **k):
        return self._Null()


def conv_work(plan, x_shape, out_shape, itemsize, has_res):
    """(algorithmic flops, compulsory bytes) of one conv layer: every tap of every output position,
    each tensor read/written once."""
    B = x_shape[0]
    if plan.transposed:
        pos = B * x_shape[2] * x_shape[3] * x_shape[4]
    else:
        pos = B * out_shape[2] * out_shape[3] * out_shape[4]
    flops = 2.0 * plan.k ** 3 * plan.cin * plan.cout * pos
    n_in = 1
    for d in x_shape:
        n_in *= d
    n_out = 1
    for d in out_shape:
        n_out *= d
    nbytes = itemsize * (n_in + n_out * (2 if has_res else 1)) + 4 * plan.k ** 3 * plan.cin * plan.cout
    return flops, nbytes


class Fp32Backend:
    """Exact path: fp32 NCDHW tensors, stb_conv3d_taps_f32."""
    name = "fp32"

    def __init__(self):
        self._plans: Dict[int, tuple] = {}
        self.prof = _NoProf()

    def _plan(self, layer) -> ops.ConvPlan:
        conv, bn = _split(layer)
        ver = (conv.weight.data_ptr(), conv.weight._version) + _bias_ver(conv) + \
              (() if bn is None else (bn.weight._version, bn.bias._version, bn.running_mean._version,
                                      bn.running_var._version, bn.running_mean.data_ptr()))
        hit = self._plans.get(id(conv))
        if hit is not None and hit[0] == ver:
            return hit[1]
        tr = isinstance(conv, nn.ConvTranspose3d)
        bnp = None if bn is None else (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        eps = 1e-5 if bn is None else bn.eps
        plan = ops.ConvPlan(conv.weight, bnp, conv.stride[0], conv.padding[0], tr,
                            conv.output_padding[0] if tr else 0, eps, bias=conv.bias)
        self._plans[id(conv)] = (ver, plan)
        return plan

    # --- layout boundary
    def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
        """gwc (+ variant-A concat) volume built straight into one [B,G+2C,D,H,W] buffer
        (replaces the two builders + torch.cat of GwcNet/gwcnet.py:175-182)."""
        B, _, H, W = gwc_l.shape
        ct = groups + (0 if cat_l is None else 2 * cat_l.shape[1])
        vol = torch.empty(B, ct, maxdisp4, H, W, device=gwc_l.device, dtype=torch.float32)
        with self.prof.bracket("gwc_volume", 0.0, 4.0 * (2 * gwc_l.numel() + B * groups * maxdisp4 * H * W)):
            ops.gwc_volume(gwc_l, gwc_r, maxdisp4, groups, out=vol, c_off=0)
        if cat_l is not None:
            with self.prof.bracket("concat_volume", 0.0, 4.0 * (2 * cat_l.numel() + B * 2 * cat_l.shape[1] * maxdisp4 * H * W)):
                ops.concat_volume(cat_l, cat_r, maxdisp4, True, out=vol, c_off=groups)
        return vol

    def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
        B, C, H, W = l.shape
        with self.prof.bracket("concat_volume", 0.0, 4.0 * (2 * l.numel() + B * 2 * C * maxdisp4 * H * W)):
            return ops.concat_volume(l, r, maxdisp4, mask_left, att_prob)

    def conv(self, layer, x, act="none", residual=None):
        plan = self._plan(layer)
        if not self.prof.enabled:
            return ops.conv3d_plan_apply(plan, x, act, residual)
        oshape = (x.shape[0], plan.cout) + tuple(plan.out_size(n) for n in x.shape[2:])
        fl, by = conv_work(plan, tuple(x.shape), oshape, 4, residual is not None)
        with self.prof.bracket("conv3d_taps_f32", fl, by):
            return ops.conv3d_plan_apply(plan, x, act, residual)

    # --- ACVNet helpers (layout boundary + the block attention core)
    def from_ncdhw(self, x):
        return ops._f32c(x)

    def cost_ncdhw(self, cost):
        """classifier output -> [B,1,D,H,W] fp32."""
        return cost

    def cost_native(self, cost):
        return cost

    def block_attention(self, qkv, bias, heads, block):
        B, C3, D, H, W = qkv.shape
        with self.prof.bracket("block_attention", 4.0 * B * D * H * W * (C3 // 3) * block[0] * block[1] * block[2],
                               4.0 * (qkv.numel() + qkv.numel() // 3)):
            return ops.block_attention(qkv, bias, heads, block, channels_last=False)

    # --- IGEV / CFNet helpers
    def gate(self, x, gate_logits):
        """FeatureAtt gate: x * sigmoid(logits)[:, :, None]."""
        return ops.feature_gate(x, gate_logits, channels_last=False)

    def cat(self, xs):
        """channel concatenation (torch.cat(dim=1) of the reference)."""
        return torch.cat(list(xs), dim=1)

    def to_ncdhw(self, x, channels=None):
        return x

    def head(self, cost, maxdisp, H, W, align_corners=False):
        B = cost.shape[0]
        with self.prof.bracket("upsample_softargmin", 0.0, 4.0 * (cost.numel() + B * H * W)):
            return ops.upsample_softargmin(cost, maxdisp, H, W, align_corners)


class TrainBackend:
    """Training path (model.train()): same interface as Fp32Backend, every call differentiable.

    Convolutions, volume builders and the head run forward AND backward in libstb200.so (autograd.py: the data gradient
    is the adjoint convolution through the same tap-list kernel, the weight gradient stb_conv3d_wgrad_f32, the volume
    and head adjoints their own kernels).  BatchNorm3d uses batch statistics and updates its running statistics exactly
    like the reference -- it is the layer's own torch module applied to the raw convolution output -- and the residual
    add / activation are torch elementwise ops, so their autograd is torch's.  fp32, reference layouts (NCDHW)."""
    name = "fp32-train"
    training_path = True          # every call differentiable; model code picks its torch-autograd forms where no adjoint kernel exists

    def __init__(self):
        self.prof = _NoProf()

    def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
        from . import autograd as A
        vol = A.gwc_volume(gwc_l, gwc_r, maxdisp4, groups)
        if cat_l is None:
            return vol
        return torch.cat((vol, A.concat_volume(cat_l, cat_r, maxdisp4, True)), 1)       # GwcNet/gwcnet.py:175-180

    def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
        from . import autograd as A
        vol = A.concat_volume(l, r, maxdisp4, mask_left)
        return vol if att_prob is None else vol * att_prob

    def conv(self, layer, x, act="none", residual=None):
        from . import autograd as A
        conv, bn = _split(layer)
        y = A.conv3d(x, conv)
        if bn is not None:
            y = bn(y)
        if residual is not None:
            y = y + residual
        if act == "relu":
            return torch.relu(y)
        if act == "leaky":
            return torch.nn.functional.leaky_relu(y, 0.01)
        if act == "mish":
            return y * torch.tanh(torch.nn.functional.softplus(y))
        return y

    def from_ncdhw(self, x):
        return x

    def to_ncdhw(self, x, channels=None):
        return x

    def cost_ncdhw(self, cost):
        return cost

    def cost_native(self, cost):
        return cost

    def cat(self, xs):
        return torch.cat(list(xs), dim=1)

    def gate(self, x, gate_logits):
        return x * torch.sigmoid(gate_logits).unsqueeze(2)

    def head(self, cost, maxdisp, H, W, align_corners=False):
        from . import autograd as A
        return A.upsample_softargmin(cost, maxdisp, H, W, align_corners)


def train_backend_for(model):
    """Backend of ``model.train()``: ``model.train_precision`` = "fp32" (default) -> TrainBackend, the exact path;
    "bf16" / "fp16" -> train16.Umma16TrainBackend (forward and data gradient of the 3-D convs on the tcgen05 kernel), kept
    on the model so that its per-layer adjoint modules and kernel plans survive across steps."""
    prec = getattr(model, "train_precision", "fp32")
    if prec == "fp32":
        return TrainBackend()
    hit = model.__dict__.get("_train16")
    if hit is None or hit.precision != prec:
        from .train16 import Umma16TrainBackend
        hit = Umma16TrainBackend(prec)
        model.__dict__["_train16"] = hit
    return hit


def make_backend(precision: str):
    if precision == "fp32":
        return Fp32Backend()
    if precision in ("bf16", "fp16", "fp16x2"):
        from .aggregation_umma import UmmaBackend
        return UmmaBackend(precision)
    raise ValueError(f"unknown precision {precision!r} (use 'fp32', 'fp16x2', 'fp16' or 'bf16')")
