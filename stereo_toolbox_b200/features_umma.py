"""2-D feature extractor of GwcNet on the tcgen05 convolution kernel ("next").

The extractor's Conv2d+BatchNorm2d(+ReLU, +residual) layers run through the SAME kernel as the 3-D aggregation
(csrc/conv3d_umma.cu): a 2-D convolution is a tap list with dz = 0 in which the image index of the (left ‖ right)
batch plays the role of depth, so one CTA keeps its weight tiles in shared memory while it marches over all images.
Activations stay channels-last fp16 from the RGB input (zero-padded to 16 channels) to the 320-channel
gwc_feature / 12-channel concat_feature; parameters are read from the reference-named torch modules
(features2d.GwcFeatures), nothing is duplicated.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .aggregation_umma import (BO_MODE, ES_VARIANT, KWMERGE, SPLIT_FLAG, TORCH_DT, _iarr, from_channels_last, pad_channels,
                               split_pack, split_weight_exponent, to_channels_last)
from .ops import ACT, _p, _stream

KWMERGE_2D = os.environ.get("STB_UMMA_KWMERGE2D", "0")
KDEPTH = os.environ.get("STB_UMMA_KDEPTH", "1") == "1"     # K-chunks accumulated in TMEM (one launch) instead of K-split passes


class Conv2dPlan:
    """Tap tables + weight tiles of one Conv2d(+BN) for stb_conv3d_umma (dz = 0 everywhere)."""

    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], cin_tensor: int, dtype, split: bool = False,
                 cin_range=None, with_shift: bool = True):
        """``cin_tensor``: logical channels of the input tensor.  split: operand-split fp16 storage (aggregation_umma.split_pack).
        ``cin_range`` = (c0, c1): the plan covers input channels [c0, c1) of ``conv`` only (one addend of a convolution over a
        channel concatenation, see UmmaGwcFeatures._convbn_multi); ``with_shift=False`` drops the BatchNorm shift (the scale
        stays folded into the weights) for all but the last addend."""
        w = conv.weight.detach().float()
        if cin_range is not None:
            w = w[:, cin_range[0]:cin_range[1]]
        cout, cin, kh_, kw_ = w.shape
        assert kh_ == kw_ and conv.groups == 1
        bias = None if conv.bias is None else conv.bias.detach().float()
        k, stride, pad, dil = kh_, conv.stride[0], conv.padding[0], conv.dilation[0]
        assert stride in (1, 2) and (stride == 1 or dil == 1)
        if cin_tensor != cin:
            wp = torch.zeros(cout, cin_tensor, k, k, device=w.device)
            wp[:, :cin] = w
            w, cin = wp, cin_tensor
        if bn is not None:
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            self.shift = (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous()
            if bias is not None:
                self.shift = (self.shift + bias * scale).contiguous()
        else:
            scale, self.shift = torch.ones(cout, device=w.device), (None if bias is None else bias.contiguous())
        if not with_shift:
            self.shift = None
        self.conv_cin = cin                                                  # real input channels (FLOP accounting)
        self.cin, self.cout, self.k, self.stride, self.pad, self.dil = cin, cout, k, stride, pad, dil
        self.in_stride = stride
        cin_st = 2 * cin if split else cin                                   # storage elements per pixel
        kc = min(cin_st, 32 if stride == 2 else 64)
        assert cin_st % kc == 0 and kc in (16, 32, 64) and not (split and kc == 16)
        self.kc, self.nk = kc, cin_st // kc
        cpad = (cout + 15) // 16 * 16
        wt = w.permute(2, 3, 0, 1) * scale.view(1, 1, -1, 1)                # [kh,kw,co,ci]
        tiles = torch.zeros(k * k, cpad, cin, device=w.device)
        tiles[:, :cout] = wt.reshape(k * k, cout, cin)
        wexp = 0
        if split:
            wexp = split_weight_exponent(tiles)
            tiles = split_pack(tiles * float(2.0 ** wexp))
        self.nwtiles = k * k
        e = [kk * dil - pad for kk in range(k)]
        par = [x % stride for x in e]
        off = [(x - p) // stride for x, p in zip(e, par)]
        mn = min(off)
        self.in_off = mn
        # kw-merge trades MMA work (3x fewer A-operand reads) for epilogue work (3 TMEM loads + 64 shuffles per 32 channels).
        # A 2-D conv has 3x fewer taps per output than a 3-D one, so with single-fp16 storage its epilogue already bounds it:
        # merging is off there by default (extractor 9.0 -> 8.0 ms at the benchmark shape); STB_UMMA_KWMERGE2D = 1 | 32 | 64
        # re-enables it for all / only Cout = 32 / only Cout = 64 layers.  Split storage issues 3x the MMAs per tap against the
        # same epilogue, so there the MMA side bounds the layer and merging (N = 3*Cn: 86-100 % instead of 40-67 % of the
        # tensor rate) is on.
        m2d = KWMERGE_2D == "1" or KWMERGE_2D == str(cout)
        if split:
            self.merge = bool(KWMERGE and KWMERGE_2D != "off" and stride == 1 and k == 3)
        else:
            self.merge = bool(KWMERGE and m2d and stride == 1 and k == 3 and cout <= 64)   # N >= 128 is math-bound without merging
        # Cin beyond one K-chunk: instead of K-split passes chained through an fp32 workspace, lay the K-chunks along a
        # pseudo-depth axis (plane = image*nk + chunk, taps of chunk c carry dz = c) so they accumulate in TMEM inside
        # one launch (conv3d_umma flags bit5).  Needs all nk*k*k weight tiles of an output-channel slice resident.
        taps_per_chunk = k if self.merge else k * k
        if split:
            # (weights of a 32-channel output slice must fit next to a short plane ring: 128->128 yes, the 320->128 lastconv no)
            self.kdepth = bool(KDEPTH and self.nk > 1 and stride == 1 and self.nk * taps_per_chunk <= 64 and self.nk <= 10
                               and self.nk * k * k * 32 * kc * 2 <= 150 * 1024)
        else:
            self.kdepth = bool(KDEPTH and self.nk > 1 and stride == 1 and not self.merge and self.nk * k * k <= 64 and self.nk <= 2)
        tiles = tiles.view(k * k, cpad, self.nk, kc)
        if self.kdepth:
            # tiles ordered [chunk][tap]: the three kw tiles of a merged tap stay contiguous
            self.wt = tiles.permute(2, 0, 1, 3).contiguous().to(dtype)
            self.nwtiles = k * k * self.nk
        else:
            self.wt = tiles.permute(0, 2, 1, 3).contiguous().to(dtype)
        dz, dh, dw, sub, widx = [], [], [], [], []
        for c in range(self.nk if self.kdepth else 1):
            for a in range(k):
                if self.merge:
                    dz.append(c); dh.append(off[a] - mn); dw.append(0); sub.append(0); widx.append(c * k * k + a * k)
                    continue
                for b in range(k):
                    dz.append(c); dh.append(off[a] - mn); dw.append(off[b] - mn)
                    sub.append(par[a] * 2 + par[b] if stride == 2 else 0)
                    widx.append(c * k * k + a * k + b)
        self.ntaps = len(dz)
        self.c = [_iarr(v) for v in (dz, dh, dw, sub, widx)]
        self.c_tb, self.c_te, self.c_z = _iarr([0]), _iarr([self.ntaps]), _iarr([0])
        self.flags = BO_MODE | (ES_VARIANT << 1) | (4 if self.merge else 0) | 16 | ((dil & 7) << 8) | (32 if self.kdepth else 0) \
            | ((SPLIT_FLAG | (wexp << 16)) if split else 0)

    def out_size(self, n):
        return (n + 2 * self.pad - self.dil * (self.k - 1) - 1) // self.stride + 1


class UmmaGwcFeatures:
    """Runs features2d.GwcFeatures (GwcNet/gwcnet.py:12-65) on the tensor-core conv kernel."""

    def __init__(self, precision: str = "fp16"):
        self.precision = precision
        self.dtype = TORCH_DT[precision]
        self.split = precision == "fp16x2"
        self.cmul = 2 if self.split else 1
        self.f16 = int(precision in ("fp16", "fp16x2"))
        self._plans: Dict[int, tuple] = {}
        self._ws = None
        self.prof = None              # bench.py's KernelProfiler: one bracket per layer ("conv2d_umma" family)

    def _plan(self, conv, bn, cin_tensor, cin_range=None, with_shift=True):
        ver = (conv.weight.data_ptr(), conv.weight._version, cin_tensor) + \
              (() if getattr(conv, "bias", None) is None else (conv.bias.data_ptr(), conv.bias._version)) + \
              (() if bn is None else (bn.weight._version, bn.bias._version, bn.running_mean._version,
                                      bn.running_var._version, bn.running_mean.data_ptr()))
        key = (id(conv), cin_range, with_shift)
        hit = self._plans.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, Conv2dPlan(conv, bn, cin_tensor, self.dtype, self.split, cin_range, with_shift))
            self._plans[key] = hit
        return hit[1]

    def conv(self, conv, bn, x, act="none", residual=None, cin_range=None, with_shift=True):
        """x [1, N, H, W, C] channels-last 16-bit (N images as depth) -> [1, N, Ho, Wo, Cout]."""
        _, N, H, W, Cst = x.shape
        C = Cst // self.cmul
        p = self._plan(conv, bn, C, cin_range, with_shift)
        Ho, Wo = p.out_size(H), p.out_size(W)
        cout_t = (p.cout + 15) // 16 * 16 if self.split else p.cout          # split rows are whole 16-channel (hi, lo) blocks
        alloc = torch.zeros if cout_t != p.cout else torch.empty
        out = alloc(1, N, Ho, Wo, cout_t * self.cmul, device=x.device, dtype=self.dtype)
        ws = None
        if p.nk > 1 and not p.kdepth:
            need = N * Ho * Wo * cout_t
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, device=x.device, dtype=torch.float32)
            ws = self._ws
        if residual is not None:
            assert residual.shape == out.shape and residual.dtype == self.dtype and residual.is_contiguous()
        call = lambda: _lib.call("stb_conv3d_umma", _p(x), _p(p.wt), _p(p.shift), _p(residual), _p(out), _p(ws), self.f16,
                                 1, Cst, p.kc, N, H, W, cout_t, p.cout, N, Ho, Wo, p.ntaps, p.c[0], p.c[1], p.c[2], p.c[3], p.c[4],
                                 None, None, p.nwtiles, 1, p.c_tb, p.c_te, p.c_z, p.c_z, p.c_z, p.in_stride, 1, N, Ho, Wo,
                                 p.in_off, p.in_off, ACT[act], 0, p.flags, 0, _stream())
        if self.prof is not None and self.prof.enabled:
            flops = 2.0 * p.k * p.k * p.conv_cin * p.cout * N * Ho * Wo
            nbytes = 2.0 * self.cmul * (x.numel() / self.cmul + out.numel() / self.cmul * (2 if residual is not None else 1))
            with self.prof.bracket("conv2d_umma", flops, nbytes,
                                   detail=f"2d {p.conv_cin}->{p.cout} k{p.k} s{p.stride} d{p.dil} @{H}x{W}"):
                call()
        else:
            call()
        return out

    def _convbn(self, seq, x, act="none", residual=None):
        return self.conv(seq[0], seq[1], x, act, residual)

    def _convbn_multi(self, seq, xs, act="none"):
        """convbn on the channel concatenation of ``xs`` (the 320-channel lastconv input).
        Split storage: a convolution over a concatenation is the sum of the convolutions of its pieces, so it runs as one
        launch per piece chained through the residual port (y1 = conv_a(l2); y2 = conv_b(l3) + y1; y = act(conv_c(l4) + shift
        + y2)): every piece has few enough K-chunks to accumulate in TMEM (the 10-chunk whole does not fit and took ten
        K-split passes through an fp32 workspace: 2.7 ms -> see profiles/), and the 320-channel concatenation is never built.
        Single-fp16 storage: the pieces are concatenated (a 16-bit copy) and convolved in one call as before."""
        if not self.split:
            return self.conv(seq[0], seq[1], torch.cat(xs, dim=-1), act)
        y, c0 = None, 0
        for i, x in enumerate(xs):
            c = x.shape[-1] // self.cmul
            last = i == len(xs) - 1
            y = self.conv(seq[0], seq[1], x, act if last else "none", residual=y, cin_range=(c0, c0 + c), with_shift=last)
            c0 += c
        assert c0 == seq[0].in_channels
        return y

    def _block(self, blk, x):
        """BasicBlock (GwcNet/submodule.py:66-91): conv1+BN+ReLU, conv2+BN, + shortcut (no ReLU after the add)."""
        y = self._convbn(blk.conv1[0], x, "relu")
        short = x if blk.downsample is None else self.conv(blk.downsample[0], blk.downsample[1], x)
        return self._convbn(blk.conv2, y, "none", residual=short)

    def _trunk(self, fe, left, right):
        """firstconv + layer1..4 shared by both extractors (features2d._Backbone) -> (l2, l3, l4), [1,2B,h,w,C] each."""
        B = left.shape[0]
        x = torch.cat((left, right), 0)                                   # [2B,3,H,W]
        H, W = x.shape[2:]
        x = to_channels_last(x.view(2 * B, 3, H, W), 16, self.dtype, split=self.split).view(1, 2 * B, H, W, 16 * self.cmul)
        fc = fe.firstconv
        x = self._convbn(fc[0], x, "relu")
        x = self._convbn(fc[2], x, "relu")
        x = self._convbn(fc[4], x, "relu")
        for blk in fe.layer1:
            x = self._block(blk, x)
        l2 = x
        for blk in fe.layer2:
            l2 = self._block(blk, l2)
        l3 = l2
        for blk in fe.layer3:
            l3 = self._block(blk, l3)
        l4 = l3
        for blk in fe.layer4:
            l4 = self._block(blk, l4)
        return l2, l3, l4

    @torch.no_grad()
    def psm(self, fe, left, right):
        """features2d.PsmFeatures (PSMNet/submodule.py:57-132) on the tensor-core conv kernel: trunk and lastconv
        (320 -> 128 k3 over the concatenation [l2, l4, branch4..1] as three chained addends, 128 -> 32 k1) run on tcgen05; the
        four SPP branches -- 64/32/16/8-pixel average pools of l4, a 1x1 conv on a handful of pixels, bilinear upsampling --
        are fp32 torch ops on tiny tensors.  Returns (feat_left, feat_right) [B,32,h,w] fp32 like the reference."""
        B = left.shape[0]
        l2, _, l4 = self._trunk(fe, left, right)
        _, N, h, w, _ = l4.shape
        l4f = from_channels_last(l4.view(N, h, w, l4.shape[-1]), split=self.split)                   # [2B,128,h,w] fp32
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            br = [F.interpolate(getattr(fe, f"branch{i}")(l4f), (h, w), mode="bilinear", align_corners=False) for i in (4, 3, 2, 1)]
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        brc = torch.cat(br, 1)                                                                       # [2B,128,h,w]
        brc = to_channels_last(brc, brc.shape[1], self.dtype, split=self.split).view(1, N, h, w, brc.shape[1] * self.cmul)
        if self.split:
            y = self._convbn_multi(fe.lastconv[0], (l2, l4, brc), "relu")
        else:
            y = self.conv(fe.lastconv[0][0], fe.lastconv[0][1], torch.cat((l2, l4, brc), dim=-1), "relu")
        y = self.conv(fe.lastconv[2], None, y)
        f = from_channels_last(y.view(N, h, w, y.shape[-1]), fe.lastconv[2].out_channels, split=self.split)
        return f[:B], f[B:]

    @torch.no_grad()
    def __call__(self, fe, left, right, concat_head=None, channels_last_out=False):
        """fe: features2d.GwcFeatures; left/right [B,3,H,W] fp32.  Returns (feat_left, feat_right) dicts of NCHW fp32
        tensors like the reference feature_extraction (the layout the volume builder reads).  ``concat_head``:
        (convbn Sequential, 1x1 Conv2d) applied to the gwc feature -- GwcNet's own lastconv by default, ACVNet's
        model-level concatconv (ACVNet/acv.py:104-107) when given."""
        B = left.shape[0]
        l2, l3, l4 = self._trunk(fe, left, right)
        if channels_last_out:
            # hand the layer2/3/4 outputs (and the concat head's output) to the volume builder as they are
            # (aggregation_umma.UmmaBackend.volume_from_cl): no torch.cat, no NCHW fp32 copy
            cat, cc = None, 0
            if concat_head is None and fe.concat_feature:
                concat_head = (fe.lastconv[0], fe.lastconv[2])
            if concat_head is not None:
                y = self._convbn_multi(concat_head[0], (l2, l3, l4), "relu")
                cat = self.conv(concat_head[1], None, y)
                cc = concat_head[1].out_channels
                if cc % 8 and not self.split:                                  # the builder reads whole 16-bit elements
                    pad = torch.zeros(cat.shape[:-1] + (8 - cc % 8,), device=cat.device, dtype=cat.dtype)
                    cat = torch.cat((cat, pad), dim=-1)
            cl = {"feats": [l2, l3, l4], "cat": cat, "B": B, "cc": cc}
            return {"_cl": cl}, {"_cl": cl}
        gwc = torch.cat((l2, l3, l4), dim=-1)                             # [1,2B,h,w,320]
        _, N, h, w, _ = gwc.shape
        gwc_f = from_channels_last(gwc.view(N, h, w, 320 * self.cmul), split=self.split)                # [2B,320,h,w] fp32
        outs = ({"gwc_feature": gwc_f[:B]}, {"gwc_feature": gwc_f[B:]})
        if concat_head is None and fe.concat_feature:
            concat_head = (fe.lastconv[0], fe.lastconv[2])
        if concat_head is not None:
            y = self._convbn_multi(concat_head[0], (l2, l3, l4), "relu")
            y = self.conv(concat_head[1], None, y)
            cat_f = from_channels_last(y.view(N, h, w, y.shape[-1]), concat_head[1].out_channels, split=self.split)
            outs[0]["concat_feature"], outs[1]["concat_feature"] = cat_f[:B], cat_f[B:]
        return outs
