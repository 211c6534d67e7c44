"""Cost-aggregation blocks of the cascade models CFNet and PCWNet on the CUDA hot path: the Mish hourglass
(models/CFNet/cfnet.py:231-271, models/PCWNet/pcwnet.py:211-251 -- identical), CFNet's two-level ``hourglassup``
(cfnet.py:178-229), PCWNet's three-level ``hourglassup`` (pcwnet.py:133-209), their heads
(trilinear upsample with ``align_corners=True`` + softmax + regression: cfnet.py:605-613, pcwnet.py:486-489;
``disparity_variance`` CFNet/submodule.py:127-133).

Parameter containers use the reference's attribute names, so sub-dicts of a CFNet / PCWNet checkpoint load unchanged.
The whole-model wrappers (uniform disparity samplers, gather-based warping, the 2-D refinement nets) are host-side glue
around these blocks and are listed under SURVEY.md section 8f.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .aggregation import convbn_3d, deconvbn_3d


def _plain_s2(cin, cout):
    return nn.Conv3d(cin, cout, kernel_size=3, stride=2, padding=1, bias=False)


class hourglass(nn.Module):
    """CFNet/cfnet.py:231-271 == PCWNet/pcwnet.py:211-251: GwcNet's hourglass with Mish instead of ReLU."""

    def __init__(self, in_channels):
        super().__init__()
        c = in_channels
        mish = lambda: nn.Identity()          # placeholder modules keep the Sequential indices (0 = convbn, 1 = Mish)
        self.conv1 = nn.Sequential(convbn_3d(c, c * 2, 3, 2, 1), mish())
        self.conv2 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 1, 1), mish())
        self.conv3 = nn.Sequential(convbn_3d(c * 2, c * 4, 3, 2, 1), mish())
        self.conv4 = nn.Sequential(convbn_3d(c * 4, c * 4, 3, 1, 1), mish())
        self.conv5 = deconvbn_3d(c * 4, c * 2)
        self.conv6 = deconvbn_3d(c * 2, c)
        self.redir1 = convbn_3d(c, c, 1, 1, 0)
        self.redir2 = convbn_3d(c * 2, c * 2, 1, 1, 0)

    def run(self, be, x):
        c1 = be.conv(self.conv1[0], x, "mish")
        c2 = be.conv(self.conv2[0], c1, "mish")
        c3 = be.conv(self.conv3[0], c2, "mish")
        c4 = be.conv(self.conv4[0], c3, "mish")
        c5 = be.conv(self.conv5, c4, "mish", residual=be.conv(self.redir2, c2))     # FMish(conv5 + redir2)
        return be.conv(self.conv6, c5, "mish", residual=be.conv(self.redir1, x))


class hourglassup(nn.Module):
    """``levels=2``: CFNet/cfnet.py:178-229; ``levels=3``: PCWNet/pcwnet.py:133-209.  The strided convs have no BN and no
    activation; their outputs are concatenated with the next-scale volumes and fused by ``combine*``."""

    def __init__(self, in_channels, levels=2):
        super().__init__()
        assert levels in (2, 3)
        c = in_channels
        self.levels = levels
        mish = lambda: nn.Identity()
        self.conv1 = _plain_s2(c, c * 2)
        self.conv2 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 1, 1), mish())
        self.conv3 = _plain_s2(c * 2, c * 4)
        self.conv4 = nn.Sequential(convbn_3d(c * 4, c * 4, 3, 1, 1), mish())
        if levels == 3:
            self.conv5 = _plain_s2(c * 4, c * 4)
            self.conv6 = nn.Sequential(convbn_3d(c * 4, c * 4, 3, 1, 1), mish())
            self.conv7 = deconvbn_3d(c * 4, c * 4)
        self.conv8 = deconvbn_3d(c * 4, c * 2)
        self.conv9 = deconvbn_3d(c * 2, c)
        self.combine1 = nn.Sequential(convbn_3d(c * 4, c * 2, 3, 1, 1), mish())
        self.combine2 = nn.Sequential(convbn_3d(c * 6, c * 4, 3, 1, 1), mish())
        self.combine3 = nn.Sequential(convbn_3d(c * 6, c * 4, 3, 1, 1), mish())
        self.redir1 = convbn_3d(c, c, 1, 1, 0)
        self.redir2 = convbn_3d(c * 2, c * 2, 1, 1, 0)
        self.redir3 = convbn_3d(c * 4, c * 4, 1, 1, 0)

    def run(self, be, x, feature4, feature5, feature6=None):
        c1 = be.conv(self.combine1[0], be.cat((be.conv(self.conv1, x), feature4)), "mish")
        c2 = be.conv(self.conv2[0], c1, "mish")
        c3 = be.conv(self.combine2[0], be.cat((be.conv(self.conv3, c2), feature5)), "mish")
        c4 = be.conv(self.conv4[0], c3, "mish")
        if self.levels == 3:
            c5 = be.conv(self.combine3[0], be.cat((be.conv(self.conv5, c4), feature6)), "mish")
            c6 = be.conv(self.conv6[0], c5, "mish")
            c4u = be.conv(self.conv7, c6, "mish", residual=be.conv(self.redir3, c4))
        else:
            c4u = c4
        c8 = be.conv(self.conv8, c4u, "mish", residual=be.conv(self.redir2, c2))
        return be.conv(self.conv9, c8, "mish", residual=be.conv(self.redir1, x))


def head_align_corners(be, cost, maxdisp, height, width):
    """F.upsample(cost, [maxdisp, H, W], mode='trilinear', align_corners=True) + softmax + disparity_regression
    (cfnet.py:605-613, pcwnet.py:486-489) -> [B, H, W]."""
    return be.head(cost, maxdisp, height, width, align_corners=True)


def disparity_variance(prob: torch.Tensor, maxdisp: int, disparity: torch.Tensor) -> torch.Tensor:
    """CFNet/submodule.py:127-133 -> [B,1,H,W]."""
    return ops.disparity_variance(prob, maxdisp, disparity)
