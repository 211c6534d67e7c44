"""IGEV-Stereo cost-volume stage on the CUDA hot path (reference: models/IGEVStereo/igev_stereo.py:23-90 hourglass,
:148-151 corr_stem / corr_feature_att / cost_agg / classifier, :205-213 the forward between the 2-D features and the
GRU loop; BasicConv / FeatureAtt models/IGEVStereo/submodule.py:9-37,228-241; Combined_Geo_Encoding_Volume
models/IGEVStereo/geometry.py:7-70).

``IGEVStereo()`` itself cannot be constructed without timm's MobileNetV2 (extractor.py:331, weights downloaded at
construction), so this file mirrors the part of the model that is the hot path -- ``IGEVCostVolume`` holds the four
sub-modules under the reference's own attribute names, so ``load_state_dict`` of an IGEVStereo checkpoint filtered to
``corr_stem. / corr_feature_att. / cost_agg. / classifier.`` loads unchanged.  The 2-D pieces of FeatureAtt (1x1 convs on
the feature maps) stay in torch; everything that touches the B x C x D x H x W volume runs in libstb200.so: group-wise
correlation volume (8 groups of 96 channels), 3x3x3 / strided / k4-s2 transposed convolutions with folded BatchNorm and
LeakyReLU(0.01), the sigmoid feature gate, the 8->1 classifier, softmax-over-D regression at 1/4 resolution (keepdim)
and the geometry-encoding pyramid + 9-tap lookup used by every GRU iteration.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops
from .aggregation import TrainBackend, make_backend
from .functional import Combined_Geo_Encoding_Volume


class BasicConv(nn.Module):
    """Parameter container with the reference's names (``conv``, ``bn``): IGEVStereo/submodule.py:9-37.
    3-D instances run on a backend (``run``); 2-D instances are ordinary torch modules (``forward``)."""

    def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, bn=True, relu=True, **kwargs):
        super().__init__()
        self.relu, self.use_bn, self.is_3d = relu, bn, is_3d
        if is_3d:
            self.conv = (nn.ConvTranspose3d if deconv else nn.Conv3d)(in_channels, out_channels, bias=False, **kwargs)
            if bn:
                self.bn = nn.BatchNorm3d(out_channels)
        else:
            self.conv = (nn.ConvTranspose2d if deconv else nn.Conv2d)(in_channels, out_channels, bias=False, **kwargs)
            if bn:
                self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):                     # 2-D use only (FeatureAtt); LeakyReLU default slope 0.01 (:36)
        x = self.conv(x)
        if self.use_bn:
            x = self.bn(x)
        return nn.functional.leaky_relu(x, 0.01) if self.relu else x

    def _layer(self):
        hit = self.__dict__.get("_pair")
        if hit is None:
            hit = nn.Sequential(self.conv, self.bn) if self.use_bn else self.conv
            self.__dict__["_pair"] = hit      # not registered: the state dict keeps conv.* / bn.* only
        return hit

    def run(self, be, x, residual=None):
        return be.conv(self._layer(), x, "leaky" if self.relu else "none", residual)


class FeatureAtt(nn.Module):
    """IGEVStereo/submodule.py:228-241: cv * sigmoid(conv1x1(BasicConv1x1(feat)))[:, :, None]."""

    def __init__(self, cv_chan, feat_chan):
        super().__init__()
        self.feat_att = nn.Sequential(BasicConv(feat_chan, feat_chan // 2, kernel_size=1, stride=1, padding=0),
                                      nn.Conv2d(feat_chan // 2, cv_chan, 1))

    def run(self, be, cv, feat):
        return be.gate(cv, self.feat_att(feat))


def _seq_run(be, seq, x):
    for m in seq:
        x = m.run(be, x)
    return x


class hourglass(nn.Module):
    """IGEVStereo/igev_stereo.py:23-90 (channels c, 2c, 4c, 6c; k4 s2 p1 transposed convs; feature gates)."""

    def __init__(self, in_channels):
        super().__init__()
        c = in_channels
        k3 = dict(is_3d=True, bn=True, relu=True, kernel_size=3, padding=1, dilation=1)
        self.conv1 = nn.Sequential(BasicConv(c, c * 2, stride=2, **k3), BasicConv(c * 2, c * 2, stride=1, **k3))
        self.conv2 = nn.Sequential(BasicConv(c * 2, c * 4, stride=2, **k3), BasicConv(c * 4, c * 4, stride=1, **k3))
        self.conv3 = nn.Sequential(BasicConv(c * 4, c * 6, stride=2, **k3), BasicConv(c * 6, c * 6, stride=1, **k3))
        up = dict(deconv=True, is_3d=True, kernel_size=(4, 4, 4), padding=(1, 1, 1), stride=(2, 2, 2))
        self.conv3_up = BasicConv(c * 6, c * 4, bn=True, relu=True, **up)
        self.conv2_up = BasicConv(c * 4, c * 2, bn=True, relu=True, **up)
        self.conv1_up = BasicConv(c * 2, 8, bn=False, relu=False, **up)
        self.agg_0 = nn.Sequential(BasicConv(c * 8, c * 4, is_3d=True, kernel_size=1, padding=0, stride=1),
                                   BasicConv(c * 4, c * 4, is_3d=True, kernel_size=3, padding=1, stride=1),
                                   BasicConv(c * 4, c * 4, is_3d=True, kernel_size=3, padding=1, stride=1))
        self.agg_1 = nn.Sequential(BasicConv(c * 4, c * 2, is_3d=True, kernel_size=1, padding=0, stride=1),
                                   BasicConv(c * 2, c * 2, is_3d=True, kernel_size=3, padding=1, stride=1),
                                   BasicConv(c * 2, c * 2, is_3d=True, kernel_size=3, padding=1, stride=1))
        self.feature_att_8 = FeatureAtt(c * 2, 64)
        self.feature_att_16 = FeatureAtt(c * 4, 192)
        self.feature_att_32 = FeatureAtt(c * 6, 160)
        self.feature_att_up_16 = FeatureAtt(c * 4, 192)
        self.feature_att_up_8 = FeatureAtt(c * 2, 64)

    def run(self, be, x, features: Sequence[torch.Tensor]):
        conv1 = self.feature_att_8.run(be, _seq_run(be, self.conv1, x), features[1])
        conv2 = self.feature_att_16.run(be, _seq_run(be, self.conv2, conv1), features[2])
        conv3 = self.feature_att_32.run(be, _seq_run(be, self.conv3, conv2), features[3])
        conv3_up = self.conv3_up.run(be, conv3)
        conv2 = _seq_run(be, self.agg_0, be.cat((conv3_up, conv2)))
        conv2 = self.feature_att_up_16.run(be, conv2, features[2])
        conv2_up = self.conv2_up.run(be, conv2)
        conv1 = _seq_run(be, self.agg_1, be.cat((conv2_up, conv1)))
        conv1 = self.feature_att_up_8.run(be, conv1, features[1])
        return self.conv1_up.run(be, conv1)


class IGEVCostVolume(nn.Module):
    """The cost-volume stage of IGEVStereo.forward (igev_stereo.py:205-213, 229-233): from the matching features and the
    left multi-scale features to (init_disp [B,1,H/4,W/4], geometry-encoding lookup ``geo_fn(disp, coords)``)."""

    def __init__(self, max_disp=192, corr_levels=2, corr_radius=4, precision="fp32"):
        super().__init__()
        self.max_disp, self.corr_levels, self.corr_radius = max_disp, corr_levels, corr_radius
        self.corr_stem = BasicConv(8, 8, is_3d=True, kernel_size=3, stride=1, padding=1)
        self.corr_feature_att = FeatureAtt(8, 96)
        self.cost_agg = hourglass(8)
        self.classifier = nn.Conv3d(8, 1, 3, 1, 1, bias=False)
        self.precision = precision
        self._be = make_backend(precision)

    def forward(self, match_left, match_right, features_left: List[torch.Tensor]):
        return self.stage(match_left, match_right, features_left)

    def stage(self, match_left, match_right, features_left: List[torch.Tensor]):
        """The whole hot path of one IGEV forward up to the GRU loop; ``IGEVStereo`` (igev_stereo.py) inherits it."""
        training = self.training
        be = TrainBackend() if training else self._be          # train(): exact fp32, forward + backward in libstb200.so
        D4 = self.max_disp // 4
        if training:
            vol = be.volume_gwc_concat(match_left, match_right, None, None, D4, 8)
        else:
            vol = ops.gwc_volume(match_left, match_right, D4, 8)                   # build_gwc_volume(..., 8)  :206
        x = self.corr_stem.run(be, be.from_ncdhw(vol))                             # :207
        x = self.corr_feature_att.run(be, x, features_left[0])                     # :208
        geo = self.cost_agg.run(be, x, features_left)                              # :209
        cost = be.cost_ncdhw(be.conv(self.classifier, geo))                        # :212  [B,1,D/4,H/4,W/4]
        if training:      # a [B,D/4,H/4,W/4] tensor: torch softmax + expectation (their autograd), :212-213
            prob = torch.softmax(cost[:, 0], dim=1)
            init_disp = (prob * torch.arange(D4, device=cost.device, dtype=prob.dtype).view(1, D4, 1, 1)).sum(1, keepdim=True)
        else:
            init_disp = ops.upsample_softargmin(cost, D4, cost.shape[3], cost.shape[4]).unsqueeze(1)   # :212-213 (keepdim)
        geo_ncdhw = be.to_ncdhw(geo, 8)
        geo_fn = Combined_Geo_Encoding_Volume(match_left.float(), match_right.float(), geo_ncdhw,
                                              num_levels=self.corr_levels, radius=self.corr_radius)   # :229-230
        return init_disp, geo_fn, geo_ncdhw
