"""Reference-named entry points of the hot path (drop-in for the module-level functions the
toolbox models import from their ``submodule.py`` / ``corr.py`` / ``geometry.py``).

Same names, argument meaning, return shapes/dtypes and assertion behaviour as the reference;
the arithmetic runs in libstb200.so.  See INTEGRATION.md for how a toolbox checkout is patched.
"""
from __future__ import annotations

import torch

from . import ops


def groupwise_correlation(fea1, fea2, num_groups):
    """GwcNet/submodule.py:44-50 -> [B,G,H,W]."""
    B, C, H, W = fea1.shape
    assert C % num_groups == 0
    cost = ops.gwc_volume(fea1, fea2, 1, num_groups)[:, :, 0]
    assert cost.shape == (B, num_groups, H, W)
    return cost


def build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
    """GwcNet/submodule.py:53-63 (also ACVNet/CFNet/PCWNet/IGEVStereo copies) -> [B,G,D,H,W]."""
    return ops.gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups)


def build_concat_volume(refimg_fea, targetimg_fea, maxdisp):
    """Variant A -- GwcNet/submodule.py:30-41, CFNet, PCWNet, PSMNet inline loop: left half masked."""
    return ops.concat_volume(refimg_fea, targetimg_fea, maxdisp, mask_left=True)


def build_concat_volume_unmasked(refimg_fea, targetimg_fea, maxdisp):
    """Variant B -- ACVNet/submodule.py:180-191, IGEVStereo/submodule.py:208-219: left half not masked."""
    return ops.concat_volume(refimg_fea, targetimg_fea, maxdisp, mask_left=False)


def disparity_regression(x, maxdisp, keepdim=False):
    """GwcNet/submodule.py:23-27 (keepdim=False); IGEVStereo/submodule.py:221-225 passes keepdim=True."""
    return ops.disparity_regression(x, maxdisp, keepdim)


class disparityregression(torch.nn.Module):
    """PSMNet/submodule.py:46-54 -> [B,1,H,W]."""

    def __init__(self, maxdisp=192):
        super().__init__()
        self.maxdisp = maxdisp

    def forward(self, x):
        return ops.disparity_regression(x, self.maxdisp, keepdim=True)


def upsample_softargmin(cost, maxdisp, height, width, align_corners=False, keepdim=False):
    """Fused replacement of the 4-line head  F.upsample(trilinear) -> squeeze -> softmax -> regression
    (GwcNet/gwcnet.py:220-223, PSMNet/stackhourglass.py:150-156)."""
    disp = ops.upsample_softargmin(cost, maxdisp, height, width, align_corners)
    return disp.unsqueeze(1) if keepdim else disp


class CorrBlock1D:
    """RAFTStereo/corr.py:110-156 -- same constructor, ``corr_pyramid`` layout and call result."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        # training (gradients reach the feature maps): the differentiable forms of autograd.py -- same forward kernels
        self._diff = torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad)
        if self._diff:
            from . import autograd as A
            corr1d_, pool_ = A.corr1d, A.avgpool_last
        else:
            corr1d_, pool_ = ops.corr1d, ops.avgpool_last
        corr = corr1d_(fmap1, fmap2, True)                    # [B,H,W1,W2]
        self._levels = [corr]
        for _ in range(self.num_levels):                      # reference stores num_levels+1 entries (:122-125)
            corr = pool_(corr)
            self._levels.append(corr)
        b, h, w1, _ = self._levels[0].shape
        self.corr_pyramid = [c.view(b * h * w1, 1, 1, c.shape[-1]) for c in self._levels]

    def __call__(self, coords):
        if self._diff:
            from . import autograd as A
            return A.corr1d_lookup(self._levels, coords, self.radius, self.num_levels)
        return ops.corr1d_lookup(self._levels, coords, self.radius, self.num_levels)

    @staticmethod
    def corr(fmap1, fmap2):
        c = ops.corr1d(fmap1, fmap2, scale=True)
        B, H, W1, W2 = c.shape
        return c.view(B, H, W1, 1, W2)


class Combined_Geo_Encoding_Volume:
    """IGEVStereo/geometry.py:7-70 (byte-identical copies in MonSter / SelectiveIGEV)."""

    def __init__(self, init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        # training: differentiable forms (autograd.py) -- same forward kernels, the layout permute as a torch op
        self._diff = torch.is_grad_enabled() and any(t.requires_grad for t in (init_fmap1, init_fmap2, geo_volume))
        if self._diff:
            from . import autograd as A
            corr = A.corr1d(init_fmap1, init_fmap2, False)
            geo = geo_volume.float().permute(0, 3, 4, 1, 2).contiguous()
            pool_ = A.avgpool_last
        else:
            corr = ops.corr1d(init_fmap1, init_fmap2, scale=False)
            geo = ops.geo_permute(geo_volume)                 # [B,H,W,C,D]
            pool_ = ops.avgpool_last
        self._geos, self._corrs = [geo], [corr]
        for _ in range(self.num_levels - 1):
            geo = pool_(geo)
            self._geos.append(geo)
        for _ in range(self.num_levels - 1):
            corr = pool_(corr)
            self._corrs.append(corr)
        b, h, w, c, _ = self._geos[0].shape
        self.geo_volume_pyramid = [g.view(b * h * w, c, 1, g.shape[-1]) for g in self._geos]
        self.init_corr_pyramid = [x.view(b * h * w, 1, 1, x.shape[-1]) for x in self._corrs]

    def __call__(self, disp, coords):
        if self._diff:
            from . import autograd as A
            return A.geo_lookup(self._geos, self._corrs, disp, coords, self.radius)
        return ops.geo_lookup(self._geos, self._corrs, disp, coords, self.radius)

    @staticmethod
    def corr(fmap1, fmap2):
        c = ops.corr1d(fmap1, fmap2, scale=False)
        B, H, W1, W2 = c.shape
        return c.view(B, H, W1, 1, W2)
