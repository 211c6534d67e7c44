"""CFNet drop-in (reference: models/CFNet/cfnet.py:12-669, models/CFNet/submodule.py).

Same constructor (``CFNet(d=192)`` -> ``cfnet(maxdisp, use_concat_volume=True)``), same ``forward(left, right)`` contract
in eval mode ([B,H,W]) and the same state-dict names/shapes.  The fused-volume stage (1/8, 1/16, 1/32 gwc+concat
volumes, dres0/1*, the two-level ``hourglassup``, ``dres3``, ``classif2``, soft-argmin + variance) and the two cascade
stages' 3-D aggregation (confidence0/1/2/3_s3/_s2, classifiers, softmax over the samples) run in libstb200.so through a
backend; the 2-D feature pyramid, the uniform disparity sampler and the gather-based warping that builds the sampled
volumes are host-side torch glue around them.
"""
from __future__ import annotations

import numpy as np
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .aggregation import TrainBackend, convbn_3d, make_backend
from .cascade import hourglass, hourglassup


class Mish(nn.Module):
    """x * tanh(softplus(x)) (CFNet/submodule.py:99-106)."""

    def forward(self, x):
        return x * torch.tanh(F.softplus(x))


def convbn(cin, cout, k, stride, pad, dilation):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=dilation if dilation > 1 else pad,
                                   dilation=dilation, bias=False), nn.BatchNorm2d(cout))


class BasicBlock(nn.Module):
    """CFNet/submodule.py:250-278: conv-bn-Mish, conv-bn, + shortcut (no activation after the add)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), Mish())
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        return out + x


class conv2DBatchNormRelu(nn.Module):
    """CFNet/submodule.py:76-97 (the 'Relu' is a Mish)."""

    def __init__(self, cin, cout, k, stride, padding, bias=True, with_bn=True):
        super().__init__()
        conv = nn.Conv2d(int(cin), int(cout), kernel_size=k, padding=padding, stride=stride, bias=bias, dilation=1)
        self.cbr_unit = nn.Sequential(conv, nn.BatchNorm2d(int(cout)), Mish()) if with_bn else nn.Sequential(conv, Mish())

    def forward(self, x):
        return self.cbr_unit(x)


class pyramidPooling(nn.Module):
    """CFNet/submodule.py:11-74, the configuration CFNet uses: pool_sizes=None, fusion_mode='sum' (icnet)."""

    def __init__(self, in_channels):
        super().__init__()
        self.path_module_list = nn.ModuleList([conv2DBatchNormRelu(in_channels, in_channels, 1, 1, 0, bias=False)
                                               for _ in range(4)])

    def forward(self, x):
        h, w = x.shape[2:]
        sizes = [(int(h / p), int(w / p)) for p in np.linspace(2, min(h, w), 4, dtype=int)][::-1]
        pp_sum = x
        for module, k in zip(self.path_module_list, sizes):
            out = module(F.avg_pool2d(x, k, stride=k, padding=0))
            pp_sum = pp_sum + 0.25 * F.interpolate(out, size=(h, w), mode="bilinear", align_corners=False)
        z = pp_sum / 2.0
        return z * torch.tanh(F.softplus(z))


def _head2d(cin, mid, cout):
    return nn.Sequential(convbn(cin, mid, 3, 1, 1, 1), Mish(), nn.Conv2d(mid, cout, kernel_size=1, padding=0, stride=1, bias=False))


class feature_extraction(nn.Module):
    """CFNet/cfnet.py:12-175: 6-scale encoder + pyramid pooling + top-down decoder; gw2..gw6 / concat_feature2..6."""

    def __init__(self, concat_feature=False, concat_feature_channel=12):
        super().__init__()
        self.concat_feature = concat_feature
        self.inplanes = 32
        self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), Mish(), convbn(32, 32, 3, 1, 1, 1), Mish(),
                                       convbn(32, 32, 3, 1, 1, 1), Mish())
        self.layer2 = self._make_layer(64, 1, 1)
        self.layer3 = self._make_layer(128, 1, 2)
        self.layer4 = self._make_layer(192, 1, 2)
        self.layer5 = self._make_layer(256, 1, 2)
        self.layer6 = self._make_layer(512, 1, 2)
        self.pyramid_pooling = pyramidPooling(512)
        up = lambda cin, cout: nn.Sequential(nn.Upsample(scale_factor=2), convbn(cin, cout, 3, 1, 1, 1), Mish())
        ic = lambda cin, cout: nn.Sequential(convbn(cin, cout, 3, 1, 1, 1), Mish())
        self.upconv6, self.iconv5 = up(512, 256), ic(512, 256)
        self.upconv5, self.iconv4 = up(256, 192), ic(384, 192)
        self.upconv4, self.iconv3 = up(192, 128), ic(256, 128)
        self.upconv3, self.iconv2 = up(128, 64), ic(128, 64)
        self.gw2, self.gw3, self.gw4 = _head2d(64, 80, 80), _head2d(128, 160, 160), _head2d(192, 160, 160)
        self.gw5, self.gw6 = _head2d(256, 320, 320), _head2d(512, 320, 320)
        if concat_feature:
            c = concat_feature_channel
            self.concat2, self.concat3 = _head2d(64, 32, c // 2), _head2d(128, 128, c)
            self.concat4, self.concat5, self.concat6 = _head2d(192, 128, c), _head2d(256, 128, c), _head2d(512, 128, c)

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample, 1, 1)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes, 1, None, 1, 1) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.firstconv(x)
        l2 = self.layer2(x)
        l3 = self.layer3(l2)
        l4 = self.layer4(l3)
        l5 = self.layer5(l4)
        l6 = self.pyramid_pooling(self.layer6(l5))
        d5 = self.iconv5(torch.cat((l5, self.upconv6(l6)), dim=1))
        d4 = self.iconv4(torch.cat((l4, self.upconv5(d5)), dim=1))
        d3 = self.iconv3(torch.cat((l3, self.upconv4(d4)), dim=1))
        d2 = self.iconv2(torch.cat((l2, self.upconv3(d3)), dim=1))
        out = {"gw2": self.gw2(d2), "gw3": self.gw3(d3), "gw4": self.gw4(d4)}
        if not self.concat_feature:
            return out
        out.update(gw5=self.gw5(d5), gw6=self.gw6(l6), concat_feature2=self.concat2(d2), concat_feature3=self.concat3(d3),
                   concat_feature4=self.concat4(d4), concat_feature5=self.concat5(d5), concat_feature6=self.concat6(l6))
        return out


def _dres(cin, c):
    ident = nn.Identity            # placeholders keep the reference's Sequential indices (1, 3 = Mish)
    return nn.Sequential(convbn_3d(cin, c, 3, 1, 1), ident(), convbn_3d(c, c, 3, 1, 1), ident())


def _dres1(c):
    return nn.Sequential(convbn_3d(c, c, 3, 1, 1), nn.Identity(), convbn_3d(c, c, 3, 1, 1))


def _classif(c):
    return nn.Sequential(convbn_3d(c, c, 3, 1, 1), nn.Identity(), nn.Conv3d(c, 1, kernel_size=3, padding=1, stride=1, bias=False))


# STB_CFNET_SAMPLED=1: build the cascade stages' sampled volumes with stb_sampled_volume_f32 instead of the torch
# expand / gather / multiply / mean / cat chain.  Opt-in until the kernel has been confirmed on hardware.
_SAMPLED_KERNEL = os.environ.get("STB_CFNET_SAMPLED", "1") == "1"


class cfnet(nn.Module):
    def __init__(self, maxdisp, use_concat_volume=False, precision="fp32"):
        super().__init__()
        self.maxdisp = maxdisp
        self.use_concat_volume = use_concat_volume
        self.sample_count_s1, self.sample_count_s2, self.sample_count_s3 = 6, 10, 14
        self.num_groups = 40
        self.concat_channels = 12 if use_concat_volume else 0
        self.feature_extraction = feature_extraction(concat_feature=use_concat_volume, concat_feature_channel=12)
        cv = self.num_groups + self.concat_channels * 2
        self.dres0, self.dres1 = _dres(cv, 32), _dres1(32)
        self.dres0_5, self.dres1_5 = _dres(cv, 64), _dres1(64)
        self.dres0_6, self.dres1_6 = _dres(cv, 64), _dres1(64)
        self.combine1 = hourglassup(32, levels=2)
        self.dres3 = hourglass(32)
        self.confidence0_s3, self.confidence1_s3 = _dres(cv + 1, 32), _dres1(32)
        self.confidence2_s3, self.confidence3_s3 = hourglass(32), hourglass(32)
        self.confidence0_s2 = _dres(self.num_groups // 2 + self.concat_channels + 1, 16)
        self.confidence1_s2 = _dres1(16)
        self.confidence2_s2, self.confidence3_s2 = hourglass(16), hourglass(16)
        self.confidence_classif0_s3, self.confidence_classif1_s3, self.confidence_classifmid_s3 = _classif(32), _classif(32), _classif(32)
        self.confidence_classif0_s2, self.confidence_classif1_s2, self.confidence_classifmid_s2 = _classif(16), _classif(16), _classif(16)
        self.classif0, self.classif1, self.classif2 = _classif(32), _classif(32), _classif(32)
        self.gamma_s3 = nn.Parameter(torch.zeros(1))
        self.beta_s3 = nn.Parameter(torch.zeros(1))
        self.gamma_s2 = nn.Parameter(torch.zeros(1))
        self.beta_s2 = nn.Parameter(torch.zeros(1))
        self.precision = precision
        self._be = make_backend(precision)

    def set_precision(self, precision):
        self.precision = precision
        self._be = make_backend(precision)
        return self

    # ---- host-side glue of the cascade (cfnet.py:437-496, submodule.py:280-349)
    def generate_search_range(self, sample_count, mn, mx, scale):
        hi = self.maxdisp // (2 ** scale) - 1
        slack = torch.clamp(sample_count - mx + mn, min=0) / 2.0
        return torch.clamp(mn - slack, min=0, max=hi), torch.clamp(mx + slack, min=0, max=hi)

    @staticmethod
    def generate_disparity_samples(mn, mx, sample_count):
        mult = (mx - mn) / (sample_count + 1)
        rng = torch.arange(1.0, sample_count + 1, 1, device=mn.device).view(sample_count, 1, 1)
        samples = mn + mult * rng                                              # UniformSampler
        return torch.cat((torch.floor(mn), samples, torch.ceil(mx)), dim=1).long()

    @staticmethod
    def _warp(left, right, samples):
        """SpatialTransformer (submodule.py:302-349): right features gathered at x - sample, zero where out of range."""
        B, C, H, W = left.shape
        S = samples.shape[1]
        xs = torch.arange(0.0, W, device=left.device).view(1, 1, 1, W).expand(B, S, H, W)
        coord = xs - samples.float()
        idx = torch.clamp(coord, min=0, max=W - 1).long()
        rf = right.unsqueeze(2).expand(B, C, S, H, W)
        warped = torch.gather(rf, dim=4, index=idx.unsqueeze(1).expand(B, C, S, H, W))
        valid = (1 - ((coord < 0) + (coord > W - 1)).float()).unsqueeze(1)
        return warped * valid, left.unsqueeze(2).expand(B, C, S, H, W)

    def _sampled_volume(self, fl, fr, key_gw, key_cat, samples, groups):
        """cost_volume_generator x2 + cat (cfnet.py:545-550): [gwc(groups) | left | warped right | samples]."""
        if _SAMPLED_KERNEL and samples.is_cuda and not torch.is_grad_enabled():    # opt-in: one fused launch (csrc/sampled.cu)
            return ops.sampled_volume(fl[key_gw], fr[key_gw], fl[key_cat], fr[key_cat], samples, groups)
        wr, lf = self._warp(fl[key_cat], fr[key_cat], samples)
        concat = torch.cat((lf, wr), dim=1)
        wr, lf = self._warp(fl[key_gw], fr[key_gw], samples)
        B, C, S, H, W = lf.shape
        gwc = (lf * wr).view(B, groups, C // groups, S, H, W).mean(dim=2)
        return torch.cat((gwc, concat, samples.unsqueeze(1).float()), dim=1)

    def _stage(self, be, vol, c0, c1, hg2, hg3, classif1, samples, softmax):
        """confidence0/1 + two hourglasses + classifier + softmax over the samples -> expectation over the samples.
        Also returns the two intermediate volumes the training-only classifiers read (cfnet.py:615-628)."""
        x = be.from_ncdhw(vol)
        c = be.conv(c0[2], be.conv(c0[0], x, "mish"), "mish")
        c = be.conv(c1[2], be.conv(c1[0], c, "mish"), "none", residual=c)
        out1 = hg2.run(be, c)
        out2 = hg3.run(be, out1)
        cost = be.cost_ncdhw(be.conv(classif1[2], be.conv(classif1[0], out2, "mish")))[:, 0]       # [B,S,H,W]
        prob = softmax(cost)
        return prob, torch.sum(prob * samples, dim=1, keepdim=True), c, out1

    def forward(self, left, right):
        if getattr(self, "channels_last", False):     # opt-in NHWC torch glue (raft_stereo.glue_channels_last)
            from .raft_stereo import glue_channels_last
            left, right = glue_channels_last(self, left, right)
        tr = self.training
        # train(): exact fp32 -- TrainBackend (forward + backward of the 3-D path in libstb200.so, batch-statistic BatchNorm);
        # the explicit-probability ops on [B,S,H,W] tensors are torch there (their autograd)
        be = TrainBackend() if tr else self._be
        if tr:
            softmax = lambda c: torch.softmax(c, dim=1)
            regression = lambda p, D: torch.sum(p * torch.arange(D, device=p.device, dtype=p.dtype).view(1, D, 1, 1), 1)
            variance = lambda p, D, d: torch.sum(p * (torch.arange(D, device=p.device, dtype=p.dtype).view(1, D, 1, 1) - d) ** 2,
                                                 1, keepdim=True)
        else:
            softmax, regression, variance = ops.softmax_d, ops.disparity_regression, ops.disparity_variance
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and self.precision != "fp32" and not tr       # exact 2-D features on the exact path
        try:
            fl, fr = self.feature_extraction(left), self.feature_extraction(right)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        H, W = left.shape[2:]
        vols = []
        for lvl, d in ((4, 8), (5, 16), (6, 32)):
            vols.append(be.volume_gwc_concat(fl[f"gw{lvl}"], fr[f"gw{lvl}"], fl.get(f"concat_feature{lvl}"),
                                             fr.get(f"concat_feature{lvl}"), self.maxdisp // d, self.num_groups))
        costs = []
        for vol, d0, d1 in zip(vols, (self.dres0, self.dres0_5, self.dres0_6), (self.dres1, self.dres1_5, self.dres1_6)):
            c = be.conv(d0[2], be.conv(d0[0], vol, "mish"), "mish")
            costs.append(be.conv(d1[2], be.conv(d1[0], c, "mish"), "none", residual=c))
        out1_4 = self.combine1.run(be, costs[0], costs[1], costs[2])
        out2_4 = self.dres3.run(be, out1_4)
        cost2 = be.cost_ncdhw(be.conv(self.classif2[2], be.conv(self.classif2[0], out2_4, "mish")))[:, 0]
        D8 = self.maxdisp // 8
        prob = softmax(cost2)
        pred2_s4 = regression(prob, D8).unsqueeze(1)
        cur = pred2_s4.detach()                                                                 # cfnet.py:541
        var = variance(prob, D8, cur).sqrt()
        mn = cur - (self.gamma_s3 + 1) * var - self.beta_s3
        mx = cur + (self.gamma_s3 + 1) * var + self.beta_s3
        up = lambda t, s: F.interpolate(t * 2, [H // s, W // s], mode="bilinear", align_corners=True)
        mn, mx = self.generate_search_range(self.sample_count_s3 + 1, up(mn, 4), up(mx, 4), scale=2)
        samples_s3 = self.generate_disparity_samples(mn, mx, self.sample_count_s3).float()
        vol3 = self._sampled_volume(fl, fr, "gw3", "concat_feature3", samples_s3, self.num_groups)
        prob3, pred1_s3, cost0_s3, out1_s3 = self._stage(be, vol3, self.confidence0_s3, self.confidence1_s3, self.confidence2_s3,
                                                         self.confidence3_s3, self.confidence_classif1_s3, samples_s3, softmax)
        cur = pred1_s3.detach()                                                                 # cfnet.py:568
        var3 = torch.sum(prob3 * (cur - samples_s3) ** 2, 1, keepdim=True).sqrt()               # disparity_variance_confidence
        mn = cur - (self.gamma_s2 + 1) * var3 - self.beta_s2
        mx = cur + (self.gamma_s2 + 1) * var3 + self.beta_s2
        mn, mx = self.generate_search_range(self.sample_count_s2 + 1, up(mn, 2), up(mx, 2), scale=1)
        samples_s2 = self.generate_disparity_samples(mn, mx, self.sample_count_s2).float()
        vol2 = self._sampled_volume(fl, fr, "gw2", "concat_feature2", samples_s2, self.num_groups // 2)
        _, pred1_s2, cost0_s2, out1_s2 = self._stage(be, vol2, self.confidence0_s2, self.confidence1_s2, self.confidence2_s2,
                                                     self.confidence3_s2, self.confidence_classif1_s2, samples_s2, softmax)
        self._last = dict(pred2_s4=pred2_s4, pred1_s3=pred1_s3)
        full = lambda t, k: F.interpolate(t * k, [H, W], mode="bilinear", align_corners=True).squeeze(1)
        if not tr:
            return full(pred1_s2, 2)
        # ---- training: the nine predictions of cfnet.py:651
        head = lambda cls, t: be.head(be.conv(cls[2], be.conv(cls[0], t, "mish")), self.maxdisp, H, W, align_corners=True)
        sampled = lambda cls, t, smp: torch.sum(softmax(be.cost_ncdhw(be.conv(cls[2], be.conv(cls[0], t, "mish")))[:, 0]) * smp,
                                                dim=1, keepdim=True)
        return [head(self.classif0, costs[0]), head(self.classif1, out1_4), full(pred2_s4, 8),
                full(sampled(self.confidence_classif0_s3, cost0_s3, samples_s3), 4),
                full(sampled(self.confidence_classifmid_s3, out1_s3, samples_s3), 4), full(pred1_s3, 4),
                full(sampled(self.confidence_classif0_s2, cost0_s2, samples_s2), 2),
                full(sampled(self.confidence_classifmid_s2, out1_s2, samples_s2), 2), full(pred1_s2, 2)]


def CFNet(d=192, **kw):
    return cfnet(d, use_concat_volume=True, **kw)
