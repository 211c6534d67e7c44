"""2-D feature extractors of PSMNet / GwcNet (parameter containers + plain torch forward).

NOT part of the hot path ("next"): these run through torch/cuDNN
exactly as in the reference.  They exist here because the drop-in models must own parameters
with the reference's state-dict names (GwcNet/gwcnet.py:12-65, PSMNet/submodule.py:57-132).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def convbn(cin, cout, k, stride, pad, dilation):
    return nn.Sequential(
        nn.Conv2d(cin, cout, k, stride, dilation if dilation > 1 else pad, dilation, bias=False),
        nn.BatchNorm2d(cout))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + (x if self.downsample is None else self.downsample(x))


class _Backbone(nn.Module):
    """firstconv + layer1..4 shared by both extractors."""

    def __init__(self):
        super().__init__()
        self.inplanes = 32
        self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), nn.ReLU(inplace=True),
                                       convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                       convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True))
        self.layer1 = self._make_layer(32, 3, 1, 1, 1)
        self.layer2 = self._make_layer(64, 16, 2, 1, 1)
        self.layer3 = self._make_layer(128, 3, 1, 1, 1)
        self.layer4 = self._make_layer(128, 3, 1, 1, 2)

    def _make_layer(self, planes, blocks, stride, pad, dilation):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, down, pad, dilation)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def trunk(self, x):
        x = self.layer1(self.firstconv(x))
        l2 = self.layer2(x)
        l3 = self.layer3(l2)
        l4 = self.layer4(l3)
        return l2, l3, l4


class GwcFeatures(_Backbone):
    """feature_extraction of GwcNet (GwcNet/gwcnet.py:12-65): 320-ch gwc feature (+ 12-ch concat feature)."""

    def __init__(self, concat_feature=False, concat_feature_channel=12):
        super().__init__()
        self.concat_feature = concat_feature
        if concat_feature:
            self.lastconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                          nn.Conv2d(128, concat_feature_channel, 1, 1, 0, bias=False))

    def forward(self, x):
        gwc = torch.cat(self.trunk(x), dim=1)
        if not self.concat_feature:
            return {"gwc_feature": gwc}
        return {"gwc_feature": gwc, "concat_feature": self.lastconv(gwc)}


class PsmFeatures(_Backbone):
    """feature_extraction of PSMNet with the SPP branches (PSMNet/submodule.py:57-132)."""

    def __init__(self):
        super().__init__()
        for i, k in ((1, 64), (2, 32), (3, 16), (4, 8)):
            setattr(self, f"branch{i}", nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)),
                                                      convbn(128, 32, 1, 1, 0, 1), nn.ReLU(inplace=True)))
        self.lastconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, 32, 1, 1, 0, bias=False))

    def forward(self, x):
        l2, _, l4 = self.trunk(x)
        size = l4.shape[2:]
        br = [F.interpolate(getattr(self, f"branch{i}")(l4), size, mode="bilinear", align_corners=False)
              for i in (4, 3, 2, 1)]
        return self.lastconv(torch.cat([l2, l4] + br, dim=1))
