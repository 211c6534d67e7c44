"""MobileNetV2-1.0 trunk with timm's module names, for the IGEV-Stereo feature network.

The reference builds ``Feature`` (models/IGEVStereo/extractor.py:327-346) around
``timm_0_5_4.create_model('mobilenetv2_100', pretrained=True, features_only=True)`` and keeps ``conv_stem``, ``bn1``,
``act1`` and ``blocks[0:6]`` of it.  timm is not a dependency of this package (and is absent from the build image), so
the trunk is defined here with the same attribute names as timm 0.5.4's EfficientNet builder -- ``conv_dw / bn1 / act1
/ conv_pw / bn2`` for the depthwise-separable first stage, ``conv_pw / bn1 / act1 / conv_dw / bn2 / act2 / conv_pwl /
bn3`` for inverted residuals -- so an IGEV-Stereo checkpoint trained with the reference loads key for key.  The stage
table is MobileNetV2's published one (t, c, n, s) = (1,16,1,1) (6,24,2,2) (6,32,3,2) (6,64,4,2) (6,96,3,1) (6,160,3,2)
[(6,320,1,1) is not used by IGEV and is not built]; activations are ReLU6, BatchNorm eps 1e-5, symmetric padding.

2-D feature extraction is outside the hot path; this is plain torch.
"""
from __future__ import annotations

import torch.nn as nn

# (expansion, out channels, repeats, first stride) -- Sandler et al. 2018, table 2
STAGES = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2))


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, 3, stride, 1, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.bn2(self.conv_pw(self.act1(self.bn1(self.conv_dw(x)))))
        return x + y if self.skip else y


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, stride, expansion):
        super().__init__()
        mid = cin * expansion
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_dw = nn.Conv2d(mid, mid, 3, stride, 1, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.act2 = nn.ReLU6(inplace=True)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return x + y if self.skip else y


class MobileNetV2Trunk(nn.Module):
    """``conv_stem`` (3->32, k3 s2) + ``bn1`` + ``act1`` + ``blocks`` (a Sequential of per-stage Sequentials)."""

    def __init__(self):
        super().__init__()
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.act1 = nn.ReLU6(inplace=True)
        stages, cin = [], 32
        for t, c, n, s in STAGES:
            blocks = []
            for i in range(n):
                stride = s if i == 0 else 1
                blocks.append(DepthwiseSeparableConv(cin, c, stride) if t == 1 else InvertedResidual(cin, c, stride, t))
                cin = c
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)

    def forward(self, x):
        return self.blocks(self.act1(self.bn1(self.conv_stem(x))))
