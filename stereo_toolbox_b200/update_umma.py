"""RAFT-Stereo's update block on the tensor-core 2-D convolution path.

Reference: models/RAFTStereo/update.py:29-44 (ConvGRU), :61-79 (BasicMotionEncoder), :7-14 (FlowHead), :115-138
(BasicMultiUpdateBlock.forward), executed ``iters`` times by raft_stereo.py:153-182.

Every convolution of the block runs on csrc/conv3d_umma.cu in the exact tensor-core format ('fp16x2': channels-last
operand-split fp16, three MMAs per K-step, fp32 accumulation in TMEM -- fp32-level accuracy, which the 32-iteration
recurrence needs: it amplifies perturbations), the recurrent state stays in that layout between iterations:

* a convolution over a channel concatenation (hx = [h, x...]) is the sum of the convolutions of its pieces, chained
  through the kernel's residual port -- no concatenated tensor is ever built; the context term (cz | cr, cq) enters the
  chain as its first addend, the bias and the gate activation (sigmoid / tanh, kernel epilogue) leave with the last;
* convz and convr read the same input and are fused into one convolution with 2C output channels;
* r * h and (1 - z) * h + z * q, the 3x3 / stride-2 average pooling and the bilinear (align_corners) resampling between
  the GRU levels are single-pass kernels on the same storage (csrc/gru2d.cu);
* the mask head only matters for the LAST iterate at test time (raft_stereo.py:176-182 upsamples every iterate only in
  training), so it runs once, in torch, on the final hidden state.

Parameters are read from the reference-named torch modules (nothing is duplicated in the state dict); fused / padded
weight views are rebuilt when a parameter changes.  Inference only (no autograd through this path).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .aggregation_umma import from_channels_last, to_channels_last
from .features_umma import UmmaGwcFeatures
from .ops import _p, _stream


class _VConv:
    """The attributes features_umma.Conv2dPlan reads from an nn.Conv2d, around fused / padded weight tensors."""

    def __init__(self, weight: torch.Tensor, bias, padding: int):
        self.weight, self.bias = weight.contiguous(), (None if bias is None else bias.contiguous())
        self.stride, self.padding, self.dilation, self.groups = (1, 1), (padding, padding), (1, 1), 1
        self.out_channels, self.in_channels = weight.shape[0], weight.shape[1]


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class UmmaRaftUpdate:
    # reference module names: ConvGRUs finest -> coarsest, the motion encoder's two flow convs, the regression head
    GRUS = ("gru08", "gru16", "gru32")
    ENC_F = ("convf1", "convf2")
    HEAD = "flow_head"

    def __init__(self, update_block: nn.Module, args):
        self.ub, self.args = update_block, args
        self.fx = UmmaGwcFeatures("fp16x2")
        self._v: Dict[str, _VConv] = {}
        self._ver = None

    # ------------------------------------------------------------------ fused / padded weights
    def _prepare(self):
        ver = tuple((p.data_ptr(), p._version) for p in self.ub.parameters())
        if ver == self._ver:
            return
        ub, v = self.ub, {}
        with torch.no_grad():
            for name in self.GRUS:
                g = getattr(ub, name)
                v[name + ".zr"] = _VConv(torch.cat((g.convz.weight, g.convr.weight), 0).detach().float(),
                                         torch.cat((g.convz.bias, g.convr.bias), 0).detach().float(), g.convz.padding[0])
            e = ub.encoder
            co = e.conv.out_channels                                   # 126: padded to a whole 16-channel block with zero rows
            cp = _pad16(co)
            w = torch.zeros(cp, e.conv.in_channels, 3, 3, device=e.conv.weight.device)
            b = torch.zeros(cp, device=w.device)
            w[:co], b[:co] = e.conv.weight.detach().float(), e.conv.bias.detach().float()
            v["enc.conv"] = _VConv(w, b, 1)
            fh = getattr(ub, self.HEAD)
            half = fh.conv1.out_channels // 2
            v["fh.conv1a"] = _VConv(fh.conv1.weight[:half].detach().float(), fh.conv1.bias[:half].detach().float(), 1)
            v["fh.conv1b"] = _VConv(fh.conv1.weight[half:].detach().float(), fh.conv1.bias[half:].detach().float(), 1)
        self._v, self._ver = v, ver
        self.fx._plans.clear()

    # ------------------------------------------------------------------ layout
    def to_cl(self, x: torch.Tensor, cpad: int = None) -> torch.Tensor:
        """[N,C,H,W] fp32 -> [1,N,H,W,2*cpad] operand-split fp16 (N images along the kernel's depth axis)."""
        N, C, H, W = x.shape
        cpad = cpad or _pad16(C)
        return to_channels_last(x, cpad, torch.float16, split=True).view(1, N, H, W, 2 * cpad)

    def from_cl(self, x: torch.Tensor, c: int = None) -> torch.Tensor:
        _, N, H, W, c2 = x.shape
        return from_channels_last(x.view(N, H, W, c2), c, split=True)

    # ------------------------------------------------------------------ elementwise / resampling kernels (csrc/gru2d.cu)
    def _rh(self, zr, h):
        out = torch.empty_like(h)
        _lib.call("stb_gru_rh_split", _p(zr), _p(h), _p(out), h.numel() // h.shape[-1], h.shape[-1] // 2, _stream())
        return out

    def _blend(self, zr, h, q):
        out = torch.empty_like(h)
        _lib.call("stb_gru_blend_split", _p(zr), _p(h), _p(q), _p(out), h.numel() // h.shape[-1], h.shape[-1] // 2, _stream())
        return out

    def pool2x(self, x):
        _, N, H, W, c2 = x.shape
        out = torch.empty(1, N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, c2, device=x.device, dtype=x.dtype)
        _lib.call("stb_pool2x_split", _p(x), _p(out), N, H, W, c2 // 2, _stream())
        return out

    def interp(self, x, dest):
        _, N, H, W, c2 = x.shape
        Ho, Wo = dest.shape[2], dest.shape[3]
        out = torch.empty(1, N, Ho, Wo, c2, device=x.device, dtype=x.dtype)
        _lib.call("stb_interp_split", _p(x), _p(out), N, H, W, Ho, Wo, c2 // 2, _stream())
        return out

    # ------------------------------------------------------------------ convolution over a concatenation
    def _conv_cat(self, conv, pieces, act, first=None):
        """conv(cat(pieces)) + first, then bias and activation: one launch per piece, chained through the residual port.
        pieces: [(tensor [1,N,H,W,2*Cpad], (c0, c1) input-channel range of ``conv`` it carries)]."""
        y = first
        for i, (t, rng) in enumerate(pieces):
            last = i == len(pieces) - 1
            y = self.fx.conv(conv, None, t, act if last else "none", residual=y, cin_range=rng, with_shift=last)
        return y

    def gru(self, name, h, czr, cq, xs):
        """ConvGRU.forward (update.py:36-44); xs: [(tensor, channels)] in the order of x_list."""
        g = getattr(self.ub, name)
        C = h.shape[-1] // 2
        rng, c0 = [], C
        for t, c in xs:
            rng.append((t, (c0, c0 + c)))
            c0 += c
        assert c0 == g.convz.in_channels, (name, c0, g.convz.in_channels)
        zr = self._conv_cat(self._v[name + ".zr"], [(h, (0, C))] + rng, "sigmoid", first=czr)
        rh = self._rh(zr, h)
        q = self._conv_cat(g.convq, [(rh, (0, C))] + rng, "tanh", first=cq)
        return self._blend(zr, h, q)

    def motion(self, flow, corr):
        """BasicMotionEncoder.forward (update.py:71-79) -> (conv output, 126 channels in a 128-channel tensor; flow in a
        16-channel tensor): the concatenation [out, flow] is consumed piecewise by the finest GRU."""
        e = self.ub.encoder
        cc = corr.shape[1]                        # (RAFT: 36 lookup channels, IGEV: 162 -> whole 64-element K-chunks)
        cor = self.fx.conv(e.convc1, None, self.to_cl(corr, 64 if cc <= 64 else (cc + 31) // 32 * 32), "relu")
        cor = self.fx.conv(e.convc2, None, cor, "relu")
        f1, f2 = getattr(e, self.ENC_F[0]), getattr(e, self.ENC_F[1])
        fl = self.to_cl(flow, 16)
        flo = self.fx.conv(f1, None, fl, "relu")
        flo = self.fx.conv(f2, None, flo, "relu")
        c1 = e.convc2.out_channels
        out = self._conv_cat(self._v["enc.conv"], [(cor, (0, c1)), (flo, (c1, c1 + f2.out_channels))], "relu")
        return out, fl

    def flow_head(self, h):
        fh = getattr(self.ub, self.HEAD)
        half = fh.conv1.out_channels // 2
        ya = self.fx.conv(self._v["fh.conv1a"], None, h, "relu")
        yb = self.fx.conv(self._v["fh.conv1b"], None, h, "relu")
        d = self._conv_cat(fh.conv2, [(ya, (0, half)), (yb, (half, 2 * half))], "none")
        return self.from_cl(d, fh.conv2.out_channels)                  # [N,2,h,w] fp32

    # ------------------------------------------------------------------ one iteration of BasicMultiUpdateBlock.forward
    def context(self, inp_list) -> List[tuple]:
        """inp_list[i] = [cz, cr, cq] (NCHW fp32, constant over the iterations) -> [(cz | cr, cq)] in kernel layout."""
        return [(self.to_cl(torch.cat((cz, cr), 1)), self.to_cl(cq)) for cz, cr, cq in inp_list]

    def step(self, net, ctx, corr, flow):
        """net: [h08, h16, h32] kernel-layout hidden states (updated list returned), ctx from ``context``; corr [N,36,h,w],
        flow [N,2,h,w] fp32.  Returns (net, delta_flow [N,2,h,w] fp32)."""
        self._prepare()
        n = self.args.n_gru_layers
        net = list(net)
        hid = lambda t: t.shape[-1] // 2
        if n == 3:
            net[2] = self.gru(self.GRUS[2], net[2], ctx[2][0], ctx[2][1], [(self.pool2x(net[1]), hid(net[1]))])
        if n >= 2:
            xs = [(self.pool2x(net[0]), hid(net[0]))]
            if n > 2:
                xs.append((self.interp(net[2], net[1]), hid(net[2])))
            net[1] = self.gru(self.GRUS[1], net[1], ctx[1][0], ctx[1][1], xs)
        mo, fl = self.motion(flow, corr)
        e = self.ub.encoder
        xs = [(mo, e.conv.out_channels), (fl, flow.shape[1])]
        if n > 1:
            xs.append((self.interp(net[1], net[0]), hid(net[1])))
        net[0] = self.gru(self.GRUS[0], net[0], ctx[0][0], ctx[0][1], xs)
        return net, self.flow_head(net[0])

    def mask(self, h08):
        """0.25 * mask head on the final hidden state (update.py:112-113,136): once per forward, torch."""
        return 0.25 * self.ub.mask(self.from_cl(h08))


class UmmaIgevUpdate(UmmaRaftUpdate):
    """IGEV-Stereo's update block (models/IGEVStereo/update.py:72-92 BasicMotionEncoder on the 162-channel geometry lookup
    and the 1-channel disparity, :115-153 BasicMultiUpdateBlock: gru16 -> gru08 -> gru04, DispHead, mask_feat_4): the same
    structure under other names; ``step`` returns (net, delta_disp [N,1,h,w])."""
    GRUS = ("gru04", "gru08", "gru16")
    ENC_F = ("convd1", "convd2")
    HEAD = "disp_head"

    def mask(self, h04):
        """mask_feat_4 of the final hidden state (update.py:152): once per forward, torch."""
        return self.ub.mask_feat_4(self.from_cl(h04))
