"""Tensor-level wrappers over the C ABI (include/stb200.h) + ``torch.ops.stb200.*`` registration.

Every wrapper validates its inputs the way the reference function does (same asserts), then
passes raw device pointers, sizes and the current CUDA stream to libstb200.so.  Inputs must be
CUDA tensors: there is deliberately no CPU / eager fallback (the CPU path is the oracle, which
lives outside the product).
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

ACT = {"none": 0, "relu": 1, "leaky": 2, "mish": 3, "sigmoid": 4, "tanh": 5}


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.StbError("stereo_toolbox_b200 ops need CUDA tensors (no CPU fallback exists)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _p(t: Optional[torch.Tensor]):
    """Raw device pointer of a tensor argument; notes the tensor's device for the launch (see _lib.note_device)."""
    if t is None:
        return ctypes.c_void_p(0)
    if t.is_cuda:
        _lib.note_device(t.device.index)
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    """Current stream of the device the call's tensors live on (every wrapper passes it as the LAST argument, after the
    tensors), not of the process-wide current device."""
    dev = _lib.call_device()
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ------------------------------------------------------------------------------------ volumes
def gwc_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, num_groups: int,
               out: Optional[torch.Tensor] = None, c_off: int = 0) -> torch.Tensor:
    """build_gwc_volume (GwcNet/submodule.py:53-63). ``out``/``c_off`` let a caller build into a
    slice of a wider [B,Ct,D,H,W] volume."""
    _need_cuda(left, right)
    B, C, H, W = left.shape
    assert right.shape == left.shape
    assert C % num_groups == 0                       # GwcNet/submodule.py:46
    left, right = _f32c(left), _f32c(right)
    if out is None:
        out = torch.empty(B, num_groups, maxdisp, H, W, device=left.device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[2:] == (maxdisp, H, W)
    _lib.call("stb_gwc_volume_f32", _p(left), _p(right), _p(out), B, C, H, W, maxdisp, num_groups,
              out.shape[1], c_off, _stream())
    return out


def concat_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, mask_left: bool = True,
                  att_prob: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  c_off: int = 0) -> torch.Tensor:
    """build_concat_volume, variant A (mask_left) GwcNet/submodule.py:30-41 / variant B
    ACVNet/submodule.py:180-191; optional fused ACVNet attention multiply (acv.py:196)."""
    _need_cuda(left, right, att_prob)
    B, C, H, W = left.shape
    assert right.shape == left.shape
    left, right = _f32c(left), _f32c(right)
    if att_prob is not None:
        att_prob = _f32c(att_prob)
        assert att_prob.numel() == B * maxdisp * H * W
    if out is None:
        out = torch.empty(B, 2 * C, maxdisp, H, W, device=left.device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[2:] == (maxdisp, H, W)
    _lib.call("stb_concat_volume_f32", _p(left), _p(right), _p(att_prob), _p(out), B, C, H, W, maxdisp,
              int(mask_left), out.shape[1], c_off, _stream())
    return out


def sampled_volume(gw_left, gw_right, cat_left, cat_right, samples, num_groups: int) -> torch.Tensor:
    """CFNet cascade-stage volume [gwc(groups) | left | warped right | samples] in one launch (CFNet/cfnet.py:472-496,
    545-550; SpatialTransformer CFNet/submodule.py:302-349).  gw_* [B,Cg,H,W], cat_* [B,Cc,H,W], samples [B,S,H,W] (integer
    valued) -> [B, groups + 2*Cc + 1, S, H, W] fp32."""
    _need_cuda(gw_left, gw_right, cat_left, cat_right, samples)
    gw_left, gw_right, cat_left, cat_right, samples = (_f32c(t) for t in (gw_left, gw_right, cat_left, cat_right, samples))
    B, Cg, H, W = gw_left.shape
    Cc, S = cat_left.shape[1], samples.shape[1]
    assert Cg % num_groups == 0                              # CFNet/submodule.py:164
    assert tuple(samples.shape) == (B, S, H, W) and tuple(cat_left.shape) == (B, Cc, H, W)
    vol = torch.empty(B, num_groups + 2 * Cc + 1, S, H, W, device=gw_left.device, dtype=torch.float32)
    _lib.call("stb_sampled_volume_f32", _p(gw_left), _p(gw_right), _p(cat_left), _p(cat_right), _p(samples), _p(vol),
              B, Cg, num_groups, Cc, S, H, W, _stream())
    return vol


def softmax_d(x: torch.Tensor) -> torch.Tensor:
    """F.softmax(x, dim=2) of [B,1,D,H,W] (or dim=1 of [B,D,H,W])."""
    _need_cuda(x)
    x = _f32c(x)
    if x.dim() == 5:
        assert x.shape[1] == 1
        B, D, plane = x.shape[0], x.shape[2], x.shape[3] * x.shape[4]
    else:
        B, D, plane = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    y = torch.empty_like(x)
    _lib.call("stb_softmax_d_f32", _p(x), _p(y), B, D, plane, _stream())
    return y


# ------------------------------------------------------------------------------------ head
def upsample_softargmin(cost: torch.Tensor, maxdisp: int, out_h: int, out_w: int,
                        align_corners: bool = False) -> torch.Tensor:
    """F.upsample(cost,[maxdisp,H,W],'trilinear') + softmax(dim=1) + disparity_regression, fused.
    cost [B,1,D,h,w] or [B,D,h,w] -> [B,out_h,out_w]."""
    _need_cuda(cost)
    if cost.dim() == 5:
        assert cost.shape[1] == 1
        cost = cost[:, 0]
    cost = _f32c(cost)
    B, D, H, W = cost.shape
    disp = torch.empty(B, out_h, out_w, device=cost.device, dtype=torch.float32)
    _lib.call("stb_upsample_softargmin_f32", _p(cost), _p(disp), B, D, H, W, maxdisp, out_h, out_w,
              int(align_corners), _stream())
    return disp


def disparity_regression(prob: torch.Tensor, maxdisp: int, keepdim: bool = False) -> torch.Tensor:
    """sum_d d * prob[:, d] (GwcNet/submodule.py:23-27; keepdim=True: PSMNet/submodule.py:46-54)."""
    _need_cuda(prob)
    assert len(prob.shape) == 4                      # GwcNet/submodule.py:24
    assert prob.shape[1] == maxdisp
    prob = _f32c(prob)
    B, D, H, W = prob.shape
    disp = torch.empty(B, H, W, device=prob.device, dtype=torch.float32)
    _lib.call("stb_disparity_regression_f32", _p(prob), _p(disp), B, D, H * W, _stream())
    return disp.unsqueeze(1) if keepdim else disp


def disparity_variance(prob: torch.Tensor, maxdisp: int, disparity: torch.Tensor) -> torch.Tensor:
    """sum_d prob[:, d] * (d - disparity)^2, CFNet/submodule.py:127-133.  prob [B,D,H,W], disparity [B,1,H,W]
    (or [B,H,W]) -> [B,1,H,W]."""
    _need_cuda(prob, disparity)
    assert len(prob.shape) == 4 and prob.shape[1] == maxdisp
    prob, disparity = _f32c(prob), _f32c(disparity)
    B, D, H, W = prob.shape
    assert disparity.numel() == B * H * W
    var = torch.empty(B, 1, H, W, device=prob.device, dtype=torch.float32)
    _lib.call("stb_disparity_variance_f32", _p(prob), _p(disparity), _p(var), B, D, H * W, _stream())
    return var


def feature_gate(x: torch.Tensor, gate_logits: torch.Tensor, channels_last: bool = False, channels: Optional[int] = None,
                 split: bool = False):
    """FeatureAtt gate (IGEVStereo/submodule.py:236-241): x * sigmoid(gate_logits)[:, :, None].
    x [B,C,D,H,W] fp32, or channels-last 16-bit [B,D,H,W,Cpad] with ``channels`` real channels; gate_logits [B,C,H,W]."""
    _need_cuda(x, gate_logits)
    g = _f32c(gate_logits)
    if not channels_last:
        x = _f32c(x)
        B, C, D, H, W = x.shape
        assert tuple(g.shape) == (B, C, H, W)
        out = torch.empty_like(x)
        _lib.call("stb_feature_gate_f32", _p(x), _p(g), _p(out), B, C, D, H, W, _stream())
        return out
    assert x.is_contiguous() and x.dtype in (torch.float16, torch.bfloat16)
    B, D, H, W, cpad = x.shape
    if split:                      # operand-split fp16 storage: two halves per logical channel
        assert x.dtype == torch.float16 and cpad % 32 == 0
        cpad //= 2
    C = channels or cpad
    assert tuple(g.shape) == (B, C, H, W)
    out = torch.empty_like(x)
    _lib.call("stb_feature_gate_cl16", _p(x), _p(g), _p(out), 2 if split else int(x.dtype == torch.float16), B, C, cpad, D, H, W,
              _stream())
    return out


# ------------------------------------------------------------------------------------ ACVNet pieces
def patch_dw(x: torch.Tensor, weight: torch.Tensor, dilation: int, out: Optional[torch.Tensor] = None,
             c_off: int = 0) -> torch.Tensor:
    """Depthwise (1,3,3) dilated conv, nn.Conv3d(C, C, (1,3,3), groups=C, dilation=d, padding=(0,d,d), bias=False)
    of ACVNet/acv.py:109-112.  x [B,Ct,D,H,W]; processes channels [c_off, c_off+C) (C = weight.shape[0]) and
    writes the same channels of ``out`` (default: a new [B,Ct,D,H,W] tensor whose other channels are undefined
    unless C == Ct)."""
    _need_cuda(x, weight)
    x = _f32c(x)
    w = _f32c(weight)
    B, Ct, D, H, W = x.shape
    C = w.shape[0]
    assert tuple(w.shape[1:]) == (1, 1, 3, 3) and c_off + C <= Ct
    if out is None:
        out = torch.empty_like(x)
    assert out.shape == x.shape and out.is_contiguous() and out.dtype == torch.float32
    _lib.call("stb_patch_dw_f32", _p(x), _p(w), _p(out), B, Ct, c_off, C, D, H, W, int(dilation), _stream())
    return out


_ATT_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def block_attention(qkv: torch.Tensor, qkv_bias: torch.Tensor, num_heads: int, block=(4, 4, 4),
                    channels_last: bool = False) -> torch.Tensor:
    """softmax(q k^T / sqrt(hd) + pad mask) v inside (b0,b1,b2) blocks (ACVNet/submodule.py:381-428 between the qkv
    Linear and final1x1).  qkv: [B,3C,D,H,W] (channels_last=False) or [B,D,H,W,3C]; returns the same layout with C
    channels.  Padded tokens take q,k,v = qkv_bias."""
    _need_cuda(qkv, qkv_bias)
    assert qkv.is_contiguous() and qkv.dtype in _ATT_DT
    if channels_last:
        B, D, H, W, C3 = qkv.shape
        C = C3 // 3
        out = torch.empty(B, D, H, W, C, device=qkv.device, dtype=qkv.dtype)
        qs = (D * H * W * C3, 1, H * W * C3, W * C3, C3)
        os_ = (D * H * W * C, 1, H * W * C, W * C, C)
    else:
        B, C3, D, H, W = qkv.shape
        C = C3 // 3
        out = torch.empty(B, C, D, H, W, device=qkv.device, dtype=qkv.dtype)
        qs = (C3 * D * H * W, D * H * W, H * W, W, 1)
        os_ = (C * D * H * W, D * H * W, H * W, W, 1)
    assert C3 == 3 * C and C % num_heads == 0 and qkv_bias.numel() == C3
    arr = lambda v: (ctypes.c_longlong * 5)(*v)
    _lib.call("stb_block_attention", _p(qkv), _p(_f32c(qkv_bias)), _p(out), _ATT_DT[qkv.dtype], B, C, num_heads,
              D, H, W, block[0], block[1], block[2], arr(qs), arr(os_), _stream())
    return out


# ------------------------------------------------------------------------------------ conv family (fp32)
class ConvPlan:
    """Tap-list form of one Conv3d / ConvTranspose3d (+ folded eval BatchNorm3d) -- see
    include/stb200.h:stb_conv3d_taps_f32.  Built once per layer and cached on the module."""

    def __init__(self, weight: torch.Tensor, bn: Optional[Tuple[torch.Tensor, ...]], stride: int, padding: int,
                 transposed: bool, output_padding: int = 0, eps: float = 1e-5,
                 bias: Optional[torch.Tensor] = None):
        w = weight.detach().float()
        dev = w.device
        if transposed:
            cin, cout = w.shape[0], w.shape[1]
        else:
            cout, cin = w.shape[0], w.shape[1]
        ks = tuple(w.shape[2:])
        assert ks[0] == ks[1] == ks[2], "cubic kernels only"
        k = ks[0]
        if bn is not None:
            gamma, beta, mean, var = [t.detach().float() for t in bn]
            scale = gamma / torch.sqrt(var + eps)
            shift = (beta - mean * scale).contiguous()
        else:
            scale = torch.ones(cout, device=dev)
            shift = None
        if bias is not None:        # conv bias (ACVNet attention_block.final1x1 / qkv Linear): BN(y + b) = scale*y + (shift + scale*b)
            sb = (scale * bias.detach().float()).contiguous()
            shift = sb if shift is None else (shift + sb).contiguous()
        # [kd,kh,kw,Cin,Cout] * scale[co]
        wt = (w.permute(2, 3, 4, 0, 1) if transposed else w.permute(2, 3, 4, 1, 0)) * scale.view(1, 1, 1, 1, -1)
        self.cin, self.cout, self.k, self.stride, self.padding = cin, cout, k, stride, padding
        self.transposed, self.output_padding = transposed, output_padding
        self.shift = shift
        self.classes = []   # (wt[T,Cin,Cout], dd, dh, dw (ctypes int arrays), T, in_stride, out_stride, (od0,oh0,ow0))
        if not transposed:
            offs = [kk - padding for kk in range(k)]
            taps = [(a, b, c) for a in range(k) for b in range(k) for c in range(k)]
            self._add_class(wt, taps, [(offs[a], offs[b], offs[c]) for a, b, c in taps], stride, 1, (0, 0, 0))
        else:
            per_dim = []
            for c in range(stride):
                per_dim.append([(kk, (c + padding - kk) // stride) for kk in range(k) if (c + padding - kk) % stride == 0])
            for cd in range(stride):
                for ch in range(stride):
                    for cw in range(stride):
                        taps, offs = [], []
                        for kd, od in per_dim[cd]:
                            for kh, oh in per_dim[ch]:
                                for kw, ow in per_dim[cw]:
                                    taps.append((kd, kh, kw))
                                    offs.append((od, oh, ow))
                        if taps:
                            self._add_class(wt, taps, offs, 1, stride, (cd, ch, cw))

    def _add_class(self, wt, taps, offs, in_stride, out_stride, origin):
        T = len(taps)
        sel = torch.stack([wt[a, b, c] for a, b, c in taps], 0).contiguous()
        arr = lambda i: (ctypes.c_int * T)(*[o[i] for o in offs])
        self.classes.append((sel, arr(0), arr(1), arr(2), T, in_stride, out_stride, origin))

    def out_size(self, n: int) -> int:
        if self.transposed:
            return (n - 1) * self.stride - 2 * self.padding + self.k + self.output_padding
        return (n + 2 * self.padding - self.k) // self.stride + 1


def conv3d_plan_apply(plan: ConvPlan, x: torch.Tensor, act: str = "none",
                      residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, residual)
    x = _f32c(x)
    B, Cin, Di, Hi, Wi = x.shape
    assert Cin == plan.cin, f"expected {plan.cin} input channels, got {Cin}"
    Do, Ho, Wo = plan.out_size(Di), plan.out_size(Hi), plan.out_size(Wi)
    out = torch.empty(B, plan.cout, Do, Ho, Wo, device=x.device, dtype=torch.float32)
    if residual is not None:
        residual = _f32c(residual)
        assert residual.shape == out.shape
    for sel, dd, dh, dw, T, in_s, out_s, (od0, oh0, ow0) in plan.classes:
        nd = (Do - od0 + out_s - 1) // out_s
        nh = (Ho - oh0 + out_s - 1) // out_s
        nw = (Wo - ow0 + out_s - 1) // out_s
        if nd <= 0 or nh <= 0 or nw <= 0:
            continue
        _lib.call("stb_conv3d_taps_f32", _p(x), _p(sel), _p(plan.shift), _p(residual), _p(out), B, Cin, Di, Hi, Wi,
                  plan.cout, Do, Ho, Wo, T, dd, dh, dw, in_s, out_s, od0, oh0, ow0, nd, nh, nw, ACT[act], _stream())
    if plan.transposed and len(plan.classes) < plan.stride ** 3:
        # k < stride (e.g. ConvTranspose3d(k=1, s=2)): some output-parity classes receive no tap at all; those positions
        # hold act(shift + residual), not whatever torch.empty left there
        have = {c[-1] for c in plan.classes}
        s_ = plan.stride
        for cd in range(s_):
            for ch in range(s_):
                for cw in range(s_):
                    if (cd, ch, cw) in have:
                        continue
                    sl = (slice(None), slice(None), slice(cd, None, s_), slice(ch, None, s_), slice(cw, None, s_))
                    v = torch.zeros_like(out[sl])
                    if plan.shift is not None:
                        v = v + plan.shift.view(1, -1, 1, 1, 1)
                    if residual is not None:
                        v = v + residual[sl]
                    out[sl] = _act_torch(v, act)
    return out


def _act_torch(x, act):
    if act == "relu":
        return torch.relu(x)
    if act == "leaky":
        return torch.nn.functional.leaky_relu(x, 0.01)
    if act == "mish":
        return x * torch.tanh(torch.nn.functional.softplus(x))
    return x


def conv3d_bn_act(x, weight, bn=None, stride=1, padding=1, act="none", residual=None, transposed=False,
                  output_padding=0):
    """Functional one-shot form (plans are normally cached by the model modules)."""
    plan = ConvPlan(weight, bn, stride, padding, transposed, output_padding)
    return conv3d_plan_apply(plan, x, act, residual)


# ------------------------------------------------------------------------------------ 1-D correlation
def corr1d(fmap1: torch.Tensor, fmap2: torch.Tensor, scale: bool = True) -> torch.Tensor:
    """CorrBlock1D.corr (RAFTStereo/corr.py:148-156) -> [B,H,W1,W2] fp32."""
    _need_cuda(fmap1, fmap2)
    fmap1, fmap2 = _f32c(fmap1), _f32c(fmap2)
    B, C, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    assert fmap2.shape[:3] == (B, C, H)
    corr = torch.empty(B, H, W1, W2, device=fmap1.device, dtype=torch.float32)
    s = 1.0 / math.sqrt(C) if scale else 1.0
    _lib.call("stb_corr1d_f32", _p(fmap1), _p(fmap2), _p(corr), B, C, H, W1, W2, ctypes.c_float(s), _stream())
    return corr


def avgpool_last(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(x,[1,2],stride=[1,2]) along the last axis."""
    _need_cuda(x)
    x = _f32c(x)
    W = x.shape[-1]
    out = torch.empty(*x.shape[:-1], W // 2, device=x.device, dtype=torch.float32)
    rows = x.numel() // W
    if W // 2 > 0:
        _lib.call("stb_avgpool_last_f32", _p(x), _p(out), rows, W, _stream())
    return out


def _ptr_array(ts: Sequence[torch.Tensor]):
    for t in ts:
        if t.is_cuda:
            _lib.note_device(t.device.index)
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def corr1d_lookup(pyramid: Sequence[torch.Tensor], coords: torch.Tensor, radius: int, num_levels: int) -> torch.Tensor:
    """CorrBlock1D.__call__ (RAFTStereo/corr.py:127-146). coords [B,2,H,W] (channel 0 used) or [B,1,H,W]."""
    _need_cuda(coords)
    B, _, H, W1 = coords.shape
    coords = _f32c(coords)
    W2 = pyramid[0].shape[-1]
    out = torch.empty(B, num_levels * (2 * radius + 1), H, W1, device=coords.device, dtype=torch.float32)
    _lib.call("stb_corr1d_lookup_f32", _ptr_array(pyramid[:num_levels]), _p(coords), coords.stride(0), _p(out),
              B, H, W1, W2, num_levels, radius, _stream())
    return out


def geo_permute(geo_volume: torch.Tensor) -> torch.Tensor:
    """[B,C,D,H,W] -> [B,H,W,C,D] (IGEVStereo/geometry.py:19)."""
    _need_cuda(geo_volume)
    g = _f32c(geo_volume)
    B, C, D, H, W = g.shape
    out = torch.empty(B, H, W, C, D, device=g.device, dtype=torch.float32)
    _lib.call("stb_geo_permute_f32", _p(g), _p(out), B, C, D, H, W, _stream())
    return out


def geo_lookup(geos, corrs, disp: torch.Tensor, coords: torch.Tensor, radius: int) -> torch.Tensor:
    """Combined_Geo_Encoding_Volume.__call__ (IGEVStereo/geometry.py:35-59)."""
    _need_cuda(disp, coords)
    disp, coords = _f32c(disp), _f32c(coords)
    B, H, W, C, D = geos[0].shape
    W2 = corrs[0].shape[-1]
    L = len(geos)
    out = torch.empty(B, L * (2 * radius + 1) * (C + 1), H, W, device=disp.device, dtype=torch.float32)
    _lib.call("stb_geo_lookup_f32", _ptr_array(geos), _ptr_array(corrs), _p(disp), _p(coords), _p(out),
              B, H, W, C, D, W2, L, radius, _stream())
    return out


# ------------------------------------------------------------------------------------ learned convex upsampling
def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, factor: int) -> torch.Tensor:
    """RAFTStereo.upsample_flow (RAFTStereo/raft_stereo.py:81-93) in one launch: flow [N,D,H,W], mask [N,9*factor^2,H,W]
    (raw logits) -> [N,D,factor*H,factor*W]."""
    _need_cuda(flow, mask)
    flow, mask = _f32c(flow), _f32c(mask)
    N, D, H, W = flow.shape
    assert tuple(mask.shape) == (N, 9 * factor * factor, H, W)
    out = torch.empty(N, D, factor * H, factor * W, device=flow.device, dtype=torch.float32)
    _lib.call("stb_convex_upsample_f32", _p(flow), _p(mask), _p(out), N, D, H, W, factor, _stream())
    return out


def context_upsample(disp_low: torch.Tensor, up_weights: torch.Tensor, scale: float = 1.0, softmax: bool = False) -> torch.Tensor:
    """context_upsample (IGEVStereo/submodule.py:243-255): disp_low [B,1,h,w], up_weights [B,9,f*h,f*w] -> [B,f*h,f*w].
    ``scale`` multiplies the coarse disparity (the caller's ``disp * 4.``), ``softmax=True`` applies the F.softmax(.., 1)
    of igev_stereo.py:164 to the 9 weights inside the kernel."""
    _need_cuda(disp_low, up_weights)
    disp_low, up_weights = _f32c(disp_low), _f32c(up_weights)
    B, c, h, w = disp_low.shape
    assert c == 1 and up_weights.shape[0] == B and up_weights.shape[1] == 9
    f = up_weights.shape[2] // h
    assert tuple(up_weights.shape[2:]) == (f * h, f * w)
    out = torch.empty(B, f * h, f * w, device=disp_low.device, dtype=torch.float32)
    _lib.call("stb_context_upsample_f32", _p(disp_low), _p(up_weights), _p(out), B, h, w, f, ctypes.c_float(scale),
              int(softmax), _stream())
    return out


# ------------------------------------------------------------------------------------ torch.ops registration
# torch.ops.stb200.<op>: same kernels behind the dispatcher, with fake (meta) kernels so that
# tracing / DDP / autocast wrappers around a patched model keep working.
_REGISTERED = False


def register_torch_ops():
    global _REGISTERED
    if _REGISTERED:
        return
    _REGISTERED = True
    lib = torch.library.Library("stb200", "DEF")
    lib.define("gwc_volume(Tensor left, Tensor right, int maxdisp, int num_groups) -> Tensor")
    lib.define("concat_volume(Tensor left, Tensor right, int maxdisp, bool mask_left) -> Tensor")
    lib.define("upsample_softargmin(Tensor cost, int maxdisp, int out_h, int out_w, bool align_corners) -> Tensor")
    lib.define("conv3d_bn_act(Tensor x, Tensor weight, Tensor? gamma, Tensor? beta, Tensor? mean, Tensor? var, "
               "int stride, int padding, str act, Tensor? residual, bool transposed, int output_padding) -> Tensor")
    lib.define("corr1d(Tensor fmap1, Tensor fmap2, bool scale) -> Tensor")
    lib.define("corr1d_lookup(Tensor[] pyramid, Tensor coords, int radius, int num_levels) -> Tensor")

    def _conv(x, weight, gamma, beta, mean, var, stride, padding, act, residual, transposed, output_padding):
        bn = None if gamma is None else (gamma, beta, mean, var)
        return conv3d_bn_act(x, weight, bn, stride, padding, act, residual, transposed, output_padding)

    lib.impl("gwc_volume", lambda l, r, d, g: gwc_volume(l, r, d, g), "CUDA")
    lib.impl("concat_volume", lambda l, r, d, m: concat_volume(l, r, d, m), "CUDA")
    lib.impl("upsample_softargmin", lambda c, d, h, w, a: upsample_softargmin(c, d, h, w, a), "CUDA")
    lib.impl("conv3d_bn_act", _conv, "CUDA")
    lib.impl("corr1d", lambda a, b, s: corr1d(a, b, s), "CUDA")
    lib.impl("corr1d_lookup", lambda p, c, r, n: corr1d_lookup(p, c, r, n), "CUDA")

    def _conv_meta(x, weight, gamma, beta, mean, var, stride, padding, act, residual, transposed, output_padding):
        k = weight.shape[2]
        cout = weight.shape[1] if transposed else weight.shape[0]
        f = (lambda n: (n - 1) * stride - 2 * padding + k + output_padding) if transposed else \
            (lambda n: (n + 2 * padding - k) // stride + 1)
        return x.new_empty(x.shape[0], cout, f(x.shape[2]), f(x.shape[3]), f(x.shape[4]))

    lib.impl("gwc_volume", lambda l, r, d, g: l.new_empty(l.shape[0], g, d, l.shape[2], l.shape[3]), "Meta")
    lib.impl("concat_volume", lambda l, r, d, m: l.new_empty(l.shape[0], 2 * l.shape[1], d, l.shape[2], l.shape[3]), "Meta")
    lib.impl("upsample_softargmin", lambda c, d, h, w, a: c.new_empty(c.shape[0], h, w), "Meta")
    lib.impl("conv3d_bn_act", _conv_meta, "Meta")
    lib.impl("corr1d", lambda a, b, s: a.new_empty(a.shape[0], a.shape[2], a.shape[3], b.shape[3]), "Meta")
    lib.impl("corr1d_lookup", lambda p, c, r, n: c.new_empty(c.shape[0], n * (2 * r + 1), c.shape[2], c.shape[3]), "Meta")
    # ---- the remaining entry points of the path (IGEV geometry lookup / gates, ACVNet pieces, explicit-probability heads)
    lib.define("avgpool_last(Tensor x) -> Tensor")
    lib.define("geo_lookup(Tensor[] geos, Tensor[] corrs, Tensor disp, Tensor coords, int radius) -> Tensor")
    lib.define("feature_gate(Tensor x, Tensor gate_logits) -> Tensor")
    lib.define("softmax_d(Tensor x) -> Tensor")
    lib.define("disparity_regression(Tensor prob, int maxdisp, bool keepdim) -> Tensor")
    lib.define("disparity_variance(Tensor prob, int maxdisp, Tensor disparity) -> Tensor")
    lib.define("patch_dw(Tensor x, Tensor weight, int dilation) -> Tensor")
    lib.define("block_attention(Tensor qkv, Tensor qkv_bias, int num_heads, int[] block) -> Tensor")
    lib.impl("avgpool_last", lambda x: avgpool_last(x), "CUDA")
    lib.impl("geo_lookup", lambda g, c, d, x, r: geo_lookup(list(g), list(c), d, x, r), "CUDA")
    lib.impl("feature_gate", lambda x, g: feature_gate(x, g), "CUDA")
    lib.impl("softmax_d", lambda x: softmax_d(x), "CUDA")
    lib.impl("disparity_regression", lambda p, d, k: disparity_regression(p, d, k), "CUDA")
    lib.impl("disparity_variance", lambda p, d, disp: disparity_variance(p, d, disp), "CUDA")
    lib.impl("patch_dw", lambda x, w, d: patch_dw(x, w, d), "CUDA")
    lib.impl("block_attention", lambda q, b, h, blk: block_attention(q, b, h, tuple(blk)), "CUDA")
    lib.impl("avgpool_last", lambda x: x.new_empty(tuple(x.shape[:-1]) + (x.shape[-1] // 2,)), "Meta")
    lib.impl("geo_lookup", lambda g, c, d, x, r: d.new_empty(g[0].shape[0], len(g) * (2 * r + 1) * (g[0].shape[3] + 1),
                                                             g[0].shape[1], g[0].shape[2]), "Meta")
    lib.impl("feature_gate", lambda x, g: torch.empty_like(x), "Meta")
    lib.impl("softmax_d", lambda x: torch.empty_like(x), "Meta")
    lib.impl("disparity_regression", lambda p, d, k: p.new_empty((p.shape[0], 1) + tuple(p.shape[2:]) if k
                                                                 else (p.shape[0],) + tuple(p.shape[2:])), "Meta")
    lib.impl("disparity_variance", lambda p, d, disp: p.new_empty((p.shape[0], 1) + tuple(p.shape[2:])), "Meta")
    lib.impl("patch_dw", lambda x, w, d: torch.empty_like(x), "Meta")
    lib.impl("block_attention", lambda q, b, h, blk: q.new_empty((q.shape[0], q.shape[1] // 3) + tuple(q.shape[2:])), "Meta")
    # autograd formulas (torch.library.register_autograd): the dispatcher ops are differentiable like the reference's
    # functions; the adjoints are the kernels of csrc/train.cu (see autograd.py)
    from . import autograd as A

    def _gwc_setup(ctx, inputs, output):
        l, r, d, g = inputs
        ctx.save_for_backward(l, r)
        ctx.cfg = (d, g)

    def _gwc_bwd(ctx, gv):
        l, r = ctx.saved_tensors
        gl, gr = A.gwc_volume_backward(gv, l, r, *ctx.cfg)
        return gl, gr, None, None

    def _cat_setup(ctx, inputs, output):
        l, r, d, m = inputs
        ctx.cfg = (tuple(l.shape), d, m)

    def _cat_bwd(ctx, gv):
        shape, d, m = ctx.cfg
        gl, gr = A.concat_volume_backward(gv, shape, d, m)
        return gl, gr, None, None

    def _head_setup(ctx, inputs, output):
        c, d, h, w, a = inputs
        ctx.save_for_backward(c)
        ctx.cfg = (d, h, w, a)

    def _head_bwd(ctx, gd):
        (c,) = ctx.saved_tensors
        c4 = c[:, 0] if c.dim() == 5 else c
        g = A.upsample_softargmin_backward(gd, c4, *ctx.cfg)
        return (g.unsqueeze(1) if c.dim() == 5 else g), None, None, None, None

    torch.library.register_autograd("stb200::gwc_volume", _gwc_bwd, setup_context=_gwc_setup, lib=lib)
    torch.library.register_autograd("stb200::concat_volume", _cat_bwd, setup_context=_cat_setup, lib=lib)
    torch.library.register_autograd("stb200::upsample_softargmin", _head_bwd, setup_context=_head_setup, lib=lib)
    register_torch_ops._lib = lib   # keep alive
