"""PCWNet drop-in (reference: models/PCWNet/pcwnet.py:12-518, models/PCWNet/submodule.py).

Same constructors (``PCWNet_G(d)`` / ``PCWNet_GC(d)``), same ``forward(left, right)`` contract in eval mode ([B,H,W]) and
the same state-dict names/shapes.  The multi-scale cost-volume path -- gwc (+ concat) volumes at 1/4, 1/8, 1/16, 1/32,
dres0/1, the three-level ``hourglassup`` that fuses them, three Mish hourglasses, ``classif3`` and the
``align_corners=True`` trilinear soft-argmin head -- runs in libstb200.so through a backend; the 2-D feature net and
the full-resolution 2-D refinement (warp, +-24 px correlation, ``refinenet_version3``) are host-side torch glue.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .aggregation import TrainBackend, convbn_3d, make_backend
from .cascade import hourglass, hourglassup
from .cfnet import BasicBlock, Mish, _classif, _dres, _dres1, _head2d, convbn


def _make_layer(inplanes, planes, blocks, stride, pad, dilation):
    downsample = None
    if stride != 1 or inplanes != planes:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(planes))
    layers = [BasicBlock(inplanes, planes, stride, downsample, pad, dilation)]
    layers += [BasicBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


class feature_extraction(nn.Module):
    """PCWNet/pcwnet.py:12-131: GwcNet-style trunk with Mish + three more scales; gw1..gw4, concat_feature1..4,
    finetune_feature."""

    def __init__(self, concat_feature=False, concat_feature_channel=12):
        super().__init__()
        self.concat_feature = concat_feature
        self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), Mish(), convbn(32, 32, 3, 1, 1, 1), Mish(),
                                       convbn(32, 32, 3, 1, 1, 1), Mish())
        self.layer1 = _make_layer(32, 32, 3, 1, 1, 1)
        self.layer2 = _make_layer(32, 64, 16, 2, 1, 1)
        self.layer3 = _make_layer(64, 128, 3, 1, 1, 1)
        self.layer4 = _make_layer(128, 128, 3, 1, 1, 2)
        self.layer5 = _make_layer(128, 192, 3, 2, 1, 1)
        self.layer7 = _make_layer(192, 256, 3, 2, 1, 1)
        self.layer9 = _make_layer(256, 512, 3, 2, 1, 1)
        self.gw2, self.gw3, self.gw4 = _head2d(192, 320, 320), _head2d(256, 320, 320), _head2d(512, 320, 320)
        self.layer11 = _head2d(320, 320, 320)
        self.layer_refine = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), Mish(), convbn(128, 32, 1, 1, 0, 1), Mish())
        if concat_feature:
            c = concat_feature_channel
            self.lastconv = _head2d(320, 128, c)
            self.concat2, self.concat3, self.concat4 = _head2d(192, 128, c), _head2d(256, 128, c), _head2d(512, 128, c)

    def forward(self, x):
        x = self.layer1(self.firstconv(x))
        l2 = self.layer2(x)
        l3 = self.layer3(l2)
        l4 = self.layer4(l3)        # 1/4
        l5 = self.layer5(l4)        # 1/8
        l6 = self.layer7(l5)        # 1/16
        l7 = self.layer9(l6)        # 1/32
        fc = torch.cat((l2, l3, l4), dim=1)
        out = {"gw1": self.layer11(fc), "gw2": self.gw2(l5), "gw3": self.gw3(l6), "gw4": self.gw4(l7)}
        if not self.concat_feature:
            return out
        out.update(concat_feature1=self.lastconv(fc), finetune_feature=self.layer_refine(fc), concat_feature2=self.concat2(l5),
                   concat_feature3=self.concat3(l6), concat_feature4=self.concat4(l7))
        return out


class refinenet_version3(nn.Module):
    """PCWNet/pcwnet.py:254-300 (2-D, full resolution)."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(in_channels, 128, 3, 1, 1, 1), Mish())
        self.conv2 = nn.Sequential(convbn(128, 128, 3, 1, 1, 1), Mish())
        self.conv3 = nn.Sequential(convbn(128, 128, 3, 1, 2, 2), Mish())
        self.conv4 = nn.Sequential(convbn(128, 128, 3, 1, 4, 4), Mish())
        self.conv5 = _make_layer(128, 96, 1, 1, 1, 8)
        self.conv6 = _make_layer(96, 64, 1, 1, 1, 16)
        self.conv7 = _make_layer(64, 32, 1, 1, 1, 1)
        self.conv8 = nn.Conv2d(32, 1, kernel_size=3, padding=1, stride=1, bias=False)

    def forward(self, x, disp):
        x = self.conv4(self.conv3(self.conv2(self.conv1(x))))
        return disp + self.conv8(self.conv7(self.conv6(self.conv5(x))))


def _kernel_path(*ts):
    return all(t.is_cuda for t in ts) and not (torch.is_grad_enabled() and any(t.requires_grad for t in ts))


def warp(x, disp):
    """PCWNet/submodule.py:122-152: bilinear resampling of the right features at x - disp (grid normalised with W-1 but
    sampled with grid_sample's default align_corners=False, exactly like the reference) and a validity mask.
    CUDA inference: one kernel (csrc/refine2d.cu); the torch form below is the differentiable / CPU-test form."""
    B, C, H, W = x.shape
    if _kernel_path(x, disp):
        from . import _lib
        from .ops import _f32c, _p, _stream
        x, disp = _f32c(x), _f32c(disp)
        out = torch.empty_like(x)
        _lib.call("stb_warp_disp_f32", _p(x), _p(disp), _p(out), B, C, H, W, _stream())
        return out
    xx = torch.arange(0, W, device=x.device).view(1, 1, 1, W).expand(B, 1, H, W).float()
    yy = torch.arange(0, H, device=x.device).view(1, 1, H, 1).expand(B, 1, H, W).float()
    gx = 2.0 * (xx - disp) / max(W - 1, 1) - 1.0
    gy = 2.0 * yy / max(H - 1, 1) - 1.0
    grid = torch.cat((gx, gy), 1).permute(0, 2, 3, 1)
    out = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    mask = F.grid_sample(torch.ones_like(x), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    mask = (mask >= 0.999).to(x.dtype)
    return out * mask


def build_correlation_volume(left, right, maxdisp):
    """build_corrleation_volume(..., num_groups=1) (PCWNet/submodule.py:104-120) -> [B, 2*maxdisp+1, H, W]: for i >= 0
    plane i+maxdisp holds mean_c L[x] * R[x - i] (x >= i).  For i < 0 the reference's slices are ``[..., :-i]`` and
    ``[..., i:]``, i.e. with k = -i the FIRST k columns of the left times the LAST k columns of the right, written to
    columns [0, k) -- reproduced as is."""
    B, C, H, W = left.shape
    if _kernel_path(left, right) and maxdisp in (4, 8, 24) and maxdisp < W:
        # all 2*maxdisp+1 planes in one pass over the staged row tiles instead of 49 slice-multiply-mean-assign sequences
        from . import _lib
        from .ops import _f32c, _p, _stream
        left, right = _f32c(left), _f32c(right)
        vol = torch.empty(B, 2 * maxdisp + 1, H, W, device=left.device, dtype=torch.float32)
        _lib.call("stb_corr_volume_1d_f32", _p(left), _p(right), _p(vol), B, C, H, W, maxdisp, _stream())
        return vol
    vol = left.new_zeros(B, 2 * maxdisp + 1, H, W)
    for i in range(-maxdisp, maxdisp + 1):
        if i > 0:
            vol[:, i + maxdisp, :, i:] = (left[:, :, :, i:] * right[:, :, :, :-i]).mean(1)
        elif i < 0:
            vol[:, i + maxdisp, :, :-i] = (left[:, :, :, :-i] * right[:, :, :, i:]).mean(1)
        else:
            vol[:, maxdisp] = (left * right).mean(1)
    return vol


class PCWNet(nn.Module):
    def __init__(self, maxdisp, use_concat_volume=False, precision="fp32"):
        super().__init__()
        self.maxdisp = maxdisp
        self.use_concat_volume = use_concat_volume
        self.num_groups = 40
        self.concat_channels = 12 if use_concat_volume else 0
        self.feature_extraction = feature_extraction(concat_feature=use_concat_volume, concat_feature_channel=12)
        self.dres0 = _dres(self.num_groups + self.concat_channels * 2, 32)
        self.dres1 = _dres1(32)
        self.combine1 = hourglassup(32, levels=3)
        self.dres2, self.dres3, self.dres4 = hourglass(32), hourglass(32), hourglass(32)
        self.classif0, self.classif1, self.classif2 = _classif(32), _classif(32), _classif(32)
        self.classif3, self.classif4 = _classif(32), _classif(32)
        self.refinenet3 = refinenet_version3(146)
        self.dispupsample = nn.Sequential(convbn(1, 32, 1, 1, 0, 1), Mish())
        self.precision = precision
        self._be = make_backend(precision)

    def set_precision(self, precision):
        self.precision = precision
        self._be = make_backend(precision)
        return self

    def forward(self, left, right):
        if getattr(self, "channels_last", False):     # opt-in NHWC torch glue (raft_stereo.glue_channels_last)
            from .raft_stereo import glue_channels_last
            left, right = glue_channels_last(self, left, right)
        if not self.use_concat_volume:
            raise NotImplementedError("PCWNet_G: the reference's refinement reads features_left['finetune_feature'], which its "
                                      "feature_extraction only returns with concat_feature=True (pcwnet.py:127-131, 493)")
        # train(): exact fp32 -- TrainBackend (forward + backward of the 3-D path in libstb200.so, batch-statistic BatchNorm)
        be = TrainBackend() if self.training else self._be
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and self.precision != "fp32" and not self.training    # exact 2-D nets on the exact path
        try:
            fl, fr = self.feature_extraction(left), self.feature_extraction(right)
            H, W = left.shape[2:]
            vols = [be.volume_gwc_concat(fl[f"gw{i}"], fr[f"gw{i}"], fl.get(f"concat_feature{i}"), fr.get(f"concat_feature{i}"),
                                         self.maxdisp // d, self.num_groups) for i, d in ((1, 4), (2, 8), (3, 16), (4, 32))]
            c = be.conv(self.dres0[2], be.conv(self.dres0[0], vols[0], "mish"), "mish")
            cost0 = be.conv(self.dres1[2], be.conv(self.dres1[0], c, "mish"), "none", residual=c)
            combine = self.combine1.run(be, cost0, vols[1], vols[2], vols[3])
            out1 = self.dres2.run(be, combine)
            out2 = self.dres3.run(be, out1)
            out3 = self.dres4.run(be, out2)
            cost3 = be.conv(self.classif3[2], be.conv(self.classif3[0], out3, "mish"))
            self._last_cost = cost3
            pred3 = be.head(cost3, self.maxdisp, H, W, align_corners=True).unsqueeze(1)          # pcwnet.py:486-489
            if self.training:                                                                    # pcwnet.py:430-458
                head = lambda cls, t: be.head(be.conv(cls[2], be.conv(cls[0], t, "mish")), self.maxdisp, H, W, align_corners=True)
                pred0, pred1, pred2 = head(self.classif0, cost0), head(self.classif1, out1), head(self.classif2, out2)
                pred_combine = head(self.classif4, combine)
            # ---- 2-D refinement at full resolution (pcwnet.py:491-506)
            up = lambda t: F.interpolate(t, [H, W], mode="bilinear", align_corners=True)
            rl, rr = up(fl["finetune_feature"]), up(fr["finetune_feature"])
            rr_warp = warp(rr, pred3)
            corr = build_correlation_volume(rl, rr_warp, 24)
            x = torch.cat((rl - rr_warp, rl, self.dispupsample(pred3), pred3, corr), dim=1)
            self._last = dict(pred3=pred3.squeeze(1))
            disp_finetune = self.refinenet3(x, pred3).squeeze(1)
            if self.training:                                                                    # pcwnet.py:480
                return [pred0, pred_combine, pred1, pred2, pred3.squeeze(1), disp_finetune]
            return disp_finetune
        finally:
            torch.backends.cudnn.allow_tf32 = prev


def PCWNet_G(d=192, **kw):
    return PCWNet(d, use_concat_volume=False, **kw)


def PCWNet_GC(d=192, **kw):
    return PCWNet(d, use_concat_volume=True, **kw)
