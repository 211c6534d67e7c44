"""Mixed-precision training backend (BASELINE config 3: "PSMNet training step bf16").

Same interface as ``aggregation.TrainBackend`` (every call differentiable, batch-statistic BatchNorm3d), but the
activations of the cost-volume path are channels-last 16-bit tensors ``[B,D,H,W,C]`` and two thirds of the convolution
FLOPs run on the tcgen05 kernel:

  forward        raw convolution (no BN fold: the batch statistics are not known yet) ........ ``stb_conv3d_umma``
  data gradient  the ADJOINT convolution of the same family, through the same kernel:
                   Conv3d(k, s=1, p)          -> Conv3d(k, 1, k-1-p) with flipped taps and swapped channel roles
                   Conv3d(k3, s=2, p1)        -> ConvTranspose3d(k3, 2, p1, output_padding from the shapes), same weight
                   ConvTranspose3d(k3,2,p1,1) -> Conv3d(k3, 2, p1), same weight ................. ``stb_conv3d_umma``
  weight gradient tensor cores (mma.sync, fp32 accumulation) on the two channels-last 16-bit operands
                   (the single-channel classifier: fp32 CUDA-core kernel) ..................... ``stb_conv3d_wgrad_cl16``

BatchNorm3d (train mode: batch statistics, running-stat update), residual adds and activations are torch elementwise /
reduction ops on the 16-bit channels-last tensors, so their autograd is torch's.  The volume builders and the fused
soft-argmin head stay fp32 (``autograd.py``: forward and adjoint kernels), with a torch layout cast at the boundary.
Gradients flow in the storage dtype, so ``bf16`` (fp32's exponent range) is the default; ``fp16`` needs loss scaling.

STATUS: written after the round-1 GPU budget was spent.  The autograd wiring (adjoint construction, channel padding,
layouts, classifier special case) is pinned on CPU with the two kernel-backed primitives ``_raw_conv`` / ``_wgrad``
replaced by torch stand-ins (tests/test_train16_cpu.py); on a GPU both primitives are calls that the inference path and
the fp32 training path already exercise, but this composition has not run on hardware yet.  Opt-in:
``model.train_precision = "bf16"``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from .aggregation import _NoProf, _split

WGRAD_TC = os.environ.get("STB_WGRAD_TC", "1") == "1"      # weight gradient on the tensor cores (csrc/wgrad_cl16.cu); 0: fp32 CUDA cores


def _pad_channels(c: int) -> int:
    for k in (16, 32, 64):
        if c <= k:
            return k
    return (c + 63) // 64 * 64


class _RawConvFn(torch.autograd.Function):
    """y = conv(x) on 16-bit channels-last tensors; backward = adjoint conv (data) + fp32 wgrad (weight)."""

    @staticmethod
    def forward(ctx, x16, weight, be, conv):
        ctx.be, ctx.conv = be, conv
        ctx.save_for_backward(x16)
        return be._raw_conv(conv, x16)

    @staticmethod
    def backward(ctx, gy):
        be, conv = ctx.be, ctx.conv
        (x16,) = ctx.saved_tensors
        gy16 = be._as_operand(gy)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = be._raw_conv(be._adjoint(conv, x16.shape, gy.shape), gy16)
            if gx.dtype != x16.dtype:                     # (an adjoint with < 8 output channels comes back fp32)
                gx = gx.to(x16.dtype)
            if gx.shape[-1] != x16.shape[-1]:             # the input carried zero padding channels
                gx = F.pad(gx, (0, x16.shape[-1] - gx.shape[-1]))
            assert gx.shape == x16.shape, (gx.shape, x16.shape)
        if ctx.needs_input_grad[1]:
            gw = be._wgrad(conv, x16, gy)
        return gx, gw, None, None


class Umma16TrainBackend:
    name = "bf16-train"

    def __init__(self, precision: str = "bf16"):
        self.precision = precision
        self.dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[precision]
        self.prof = _NoProf()
        self._inner = None                                # UmmaBackend, created on first use (needs the CUDA library)
        self._adj: Dict[int, tuple] = {}                  # id(conv) -> (weight version, adjoint module kept alive)

    # ------------------------------------------------------------------ kernel-backed primitives
    def _umma(self):
        if self._inner is None:
            from .aggregation_umma import UmmaBackend
            self._inner = UmmaBackend(self.precision)
        return self._inner

    def _raw_conv(self, conv: nn.Module, x16: torch.Tensor) -> torch.Tensor:
        """conv(x) without BatchNorm / activation on the tcgen05 kernel: [B,D,H,W,Cin(_pad)] 16-bit ->
        [B,Do,Ho,Wo,Cout] 16-bit (fp32 when Cout < 8: the classifier)."""
        return self._umma().conv(conv, x16.contiguous(), "none", None)

    def _wgrad(self, conv: nn.Module, x16: torch.Tensor, gy: torch.Tensor) -> torch.Tensor:
        """dL/dW in the layout of ``conv.weight`` (fp32), from the 16-bit input and the output gradient."""
        from . import _lib
        from .aggregation_umma import from_channels_last
        from .ops import _p, _stream
        tr = isinstance(conv, nn.ConvTranspose3d)
        cin, cout = (conv.weight.shape[0], conv.weight.shape[1]) if tr else (conv.weight.shape[1], conv.weight.shape[0])
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if WGRAD_TC and gy.dtype == torch.float32 and gy.shape[-1] == 1 and not tr:
            # single-channel classifier (fp32 logit gradient [B,D,H,W,1]): rounded to the training dtype like every other
            # gradient of this path and zero-padded to one 32-channel block, so that it runs on the tensor-core kernel too
            # (the fp32 CUDA-core kernel computes a 32x32 channel block for the one live column: 5.8 ms per classifier at
            # 576x960 against 0.4 ms here)
            g32 = torch.zeros(gy.shape[:-1] + (32,), device=gy.device, dtype=self.dtype)
            g32[..., 0] = gy[..., 0].to(self.dtype)
            gy = g32
        if (WGRAD_TC and gy.dtype == self.dtype and x16.shape[-1] % 32 == 0 and gy.shape[-1] % 32 == 0 and k <= 3
                and s in (1, 2)):
            # tensor-core weight gradient straight from the channels-last 16-bit operands (csrc/wgrad_cl16.cu): no fp32 NCDHW
            # copies, fp32 accumulation.  Zero-padded channels come back as zero rows / columns and are sliced off.
            x16, g16 = x16.contiguous(), gy.contiguous()
            P, Q = (g16, x16) if tr else (x16, g16)
            dw = torch.zeros(k, k, k, P.shape[-1], Q.shape[-1], device=x16.device, dtype=torch.float32)
            _lib.call("stb_conv3d_wgrad_cl16", _p(P), _p(Q), _p(dw), int(self.dtype == torch.float16), P.shape[0], P.shape[1],
                      P.shape[2], P.shape[3], P.shape[4], Q.shape[1], Q.shape[2], Q.shape[3], Q.shape[4], k, p, s, 0, _stream())
            cp, cq = (cout, cin) if tr else (cin, cout)
            return dw[..., :cp, :cq].permute(4, 3, 0, 1, 2).contiguous()
        xf = from_channels_last(x16.contiguous(), cin)                       # fp32 [B,Cin,D,H,W], padding dropped
        if gy.dtype == torch.float32:                                        # classifier: [B,D,H,W,1] fp32
            gf = gy.permute(0, 4, 1, 2, 3).contiguous()
        else:
            gf = from_channels_last(gy.contiguous(), cout)
        P, Q = (gf, xf) if tr else (xf, gf)
        dw = torch.zeros(k, k, k, P.shape[1], Q.shape[1], device=xf.device, dtype=torch.float32)
        _lib.call("stb_conv3d_wgrad_f32", _p(P), _p(Q), _p(dw), P.shape[0], P.shape[1], P.shape[2], P.shape[3],
                  P.shape[4], Q.shape[1], Q.shape[2], Q.shape[3], Q.shape[4], k, p, s, _stream())
        return dw.permute(4, 3, 0, 1, 2).contiguous()                        # both layouts: weight[cq, cp, kd, kh, kw]

    # ------------------------------------------------------------------ host logic (pinned on CPU)
    def _as_operand(self, g: torch.Tensor) -> torch.Tensor:
        """An output gradient as a conv INPUT: storage dtype, contiguous, channel count legal for the kernel."""
        g = g.to(self.dtype) if g.dtype != self.dtype else g
        cp = _pad_channels(g.shape[-1])
        if cp != g.shape[-1]:
            g = F.pad(g, (0, cp - g.shape[-1]))
        return g.contiguous()

    def _adjoint(self, conv: nn.Module, x_shape, gy_shape) -> nn.Module:
        """The convolution whose forward is the data gradient of ``conv`` (kept per layer, refreshed in place when the
        optimizer has changed the weight, so that the kernel-plan cache of the inner backend sees stable module ids)."""
        w = conv.weight
        tr = isinstance(conv, nn.ConvTranspose3d)
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        op = 0
        if not tr and s > 1:
            ops_ = {x_shape[i] - ((gy_shape[i] - 1) * s - 2 * p + k) for i in (1, 2, 3)}      # NDHWC: dims 1..3
            if len(ops_) != 1:
                raise ValueError(f"output_padding of the adjoint must be equal along D, H and W (got {sorted(ops_)})")
            op = ops_.pop()
        ver = (w.data_ptr(), w._version, op)
        hit = self._adj.get(id(conv))
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            if tr:                                       # ConvTranspose3d [Cin,Cout,k,k,k] -> Conv3d(Cout -> Cin), same tensor
                cin, cout = w.shape[0], w.shape[1]
                adj = hit[1] if hit is not None and isinstance(hit[1], nn.Conv3d) else \
                    nn.Conv3d(cout, cin, k, s, p, bias=False).to(w.device)
                adj.weight.requires_grad_(False)
                adj.weight.copy_(w.detach())
            elif s == 1:                                 # flipped taps, channel roles swapped
                cout, cin = w.shape[0], w.shape[1]
                adj = hit[1] if hit is not None and isinstance(hit[1], nn.Conv3d) else \
                    nn.Conv3d(cout, cin, k, 1, k - 1 - p, bias=False).to(w.device)
                adj.weight.requires_grad_(False)
                adj.weight.copy_(w.detach().flip(2, 3, 4).transpose(0, 1))
            else:                                        # strided Conv3d [Cout,Cin,...] -> ConvTranspose3d(Cout -> Cin), same tensor
                cout, cin = w.shape[0], w.shape[1]
                adj = None
                if hit is not None and isinstance(hit[1], nn.ConvTranspose3d) and hit[1].output_padding[0] == op:
                    adj = hit[1]
                if adj is None:
                    adj = nn.ConvTranspose3d(cout, cin, k, s, p, output_padding=op, bias=False).to(w.device)
                adj.weight.requires_grad_(False)
                adj.weight.copy_(w.detach())
        self._adj[id(conv)] = (ver, adj)
        return adj

    # ------------------------------------------------------------------ backend interface (aggregation.TrainBackend)
    def _to_cl(self, vol: torch.Tensor) -> torch.Tensor:
        """fp32 NCDHW -> channels-last storage dtype, channels zero-padded to a legal K width (torch glue: differentiable)."""
        v = vol.permute(0, 2, 3, 4, 1).to(self.dtype)
        cp = _pad_channels(v.shape[-1])
        if cp != v.shape[-1]:
            v = F.pad(v, (0, cp - v.shape[-1]))
        return v.contiguous()

    def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
        from . import autograd as A
        vol = A.gwc_volume(gwc_l, gwc_r, maxdisp4, groups)
        if cat_l is not None:
            vol = torch.cat((vol, A.concat_volume(cat_l, cat_r, maxdisp4, True)), 1)
        return self._to_cl(vol)

    def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
        from . import autograd as A
        vol = A.concat_volume(l, r, maxdisp4, mask_left)
        return self._to_cl(vol if att_prob is None else vol * att_prob)

    def conv(self, layer, x, act="none", residual=None):
        conv, bn = _split(layer)
        if conv.bias is not None:
            raise NotImplementedError("biased 3-D convs are not on the 16-bit training path")
        y = _RawConvFn.apply(x, conv.weight, self, conv)
        if bn is not None:                               # batch statistics; fp32 parameters on a 16-bit tensor
            y = bn(y.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
        if residual is not None:
            y = y + residual
        if act == "relu":
            y = torch.relu(y)
        elif act == "leaky":
            y = F.leaky_relu(y, 0.01)
        elif act == "mish":
            y = y * torch.tanh(F.softplus(y))
        elif act != "none":
            raise ValueError(act)
        return y.contiguous()

    def from_ncdhw(self, x):
        return self._to_cl(x)

    def to_ncdhw(self, x, channels=None):
        x = x if channels is None else x[..., :channels]
        return x.permute(0, 4, 1, 2, 3).float().contiguous()

    def cost_ncdhw(self, cost):
        return cost.permute(0, 4, 1, 2, 3)

    def cost_native(self, cost):
        return cost.permute(0, 2, 3, 4, 1)

    def cat(self, xs):
        xs = list(xs)
        c = sum(t.shape[-1] for t in xs)
        cp = _pad_channels(c)
        if cp != c:
            xs.append(torch.zeros(xs[0].shape[:-1] + (cp - c,), device=xs[0].device, dtype=xs[0].dtype))
        return torch.cat(xs, dim=-1)

    def gate(self, x, gate_logits):
        g = torch.sigmoid(gate_logits).permute(0, 2, 3, 1).unsqueeze(1).to(x.dtype)      # [B,1,H,W,C]
        if g.shape[-1] != x.shape[-1]:
            g = F.pad(g, (0, x.shape[-1] - g.shape[-1]))
        return x * g

    def head(self, cost, maxdisp, H, W, align_corners=False):
        from . import autograd as A
        assert cost.dtype == torch.float32 and cost.shape[-1] == 1
        B, D, h, w, _ = cost.shape
        return A.upsample_softargmin(cost.reshape(B, D, h, w), maxdisp, H, W, align_corners)
