"""Oracle restatement of the hot-path *operators* (CPU, fp32/fp64 torch + explicit index math).

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

Every function cites the reference lines (relative to
/root/reference/stereo_toolbox/models/) whose arithmetic it restates.  Where the
reference delegates to a torch library call with non-obvious semantics
(trilinear ``F.upsample``, ``F.grid_sample``) the restatement here spells the
index math out explicitly, so that it also pins the formula the CUDA kernels
implement; ``tests/test_oracle_golden.py`` checks these against fixtures made by
running the reference itself.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- volumes
def groupwise_correlation(fea1: torch.Tensor, fea2: torch.Tensor, num_groups: int) -> torch.Tensor:
    """GwcNet/submodule.py:44-50 -- per-group MEAN of the channel products."""
    B, C, H, W = fea1.shape
    if C % num_groups:
        raise AssertionError("C must be divisible by num_groups")
    cpg = C // num_groups
    return (fea1 * fea2).reshape(B, num_groups, cpg, H, W).sum(2) / cpg


def build_gwc_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, num_groups: int) -> torch.Tensor:
    """GwcNet/submodule.py:53-63 (same arithmetic in ACVNet:228-238, CFNet:171-181,
    PCWNet:109-119, IGEVStereo:161-171).

    vol[b,g,d,h,w] = mean_c L[b,g*k+c,h,w] * R[b,g*k+c,h,w-d]  (w >= d), 0 otherwise.
    """
    B, C, H, W = left.shape
    vol = left.new_zeros(B, num_groups, maxdisp, H, W)
    for d in range(min(maxdisp, W)):
        vol[:, :, d, :, d:] = groupwise_correlation(left[..., d:], right[..., : W - d], num_groups)
    return vol


def build_concat_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, mask_left: bool = True) -> torch.Tensor:
    """Variant A (mask_left=True): GwcNet/submodule.py:30-41, CFNet:141-152, PCWNet:86-97 and
    the inline loop PSMNet/stackhourglass.py:111-120 -- both halves zero where w < d.
    Variant B (mask_left=False): ACVNet/submodule.py:180-191, IGEVStereo/submodule.py:208-219 --
    the left half is copied for every w; the right half is still zero where w < d.
    """
    B, C, H, W = left.shape
    vol = left.new_zeros(B, 2 * C, maxdisp, H, W)
    for d in range(maxdisp):
        if mask_left:
            if d < W:
                vol[:, :C, d, :, d:] = left[..., d:]
        else:
            vol[:, :C, d] = left
        if d < W:
            vol[:, C:, d, :, d:] = right[..., : W - d]
    return vol


def attention_weighted_volume(att: torch.Tensor, concat_volume: torch.Tensor) -> torch.Tensor:
    """ACVNet/acv.py:196 -- softmax over the disparity axis of [B,1,D,H,W] times the concat volume."""
    return torch.softmax(att, dim=2) * concat_volume


# --------------------------------------------------------------------------- head
def _linear_taps(out_size: int, in_size: int, align_corners: bool, device=None):
    """Source index pair and weight of 1-D linear resampling as ATen's upsample_*linear does it
    (the op behind F.upsample(mode='trilinear'), GwcNet/gwcnet.py:220, PSMNet/stackhourglass.py:150;
    align_corners=True at CFNet/cfnet.py:605-613, PCWNet/pcwnet.py:486)."""
    dst = torch.arange(out_size, dtype=torch.float64, device=device)
    if align_corners:
        scale = (in_size - 1) / (out_size - 1) if out_size > 1 else 0.0
        src = dst * scale
    else:
        scale = in_size / out_size
        src = ((dst + 0.5) * scale - 0.5).clamp_min(0.0)
    i0 = src.floor().to(torch.int64).clamp_max(in_size - 1)
    i1 = (i0 + 1).clamp_max(in_size - 1)
    w1 = (src - i0.to(torch.float64)).to(torch.float32)
    return i0, i1, w1


def trilinear_upsample(cost: torch.Tensor, out_d: int, out_h: int, out_w: int, align_corners: bool = False) -> torch.Tensor:
    """[B,D,H,W] -> [B,out_d,out_h,out_w], separable linear interpolation in W, then H, then D."""
    B, D, H, W = cost.shape
    i0, i1, t = _linear_taps(out_w, W, align_corners, cost.device)
    x = cost[..., i0] * (1 - t) + cost[..., i1] * t
    i0, i1, t = _linear_taps(out_h, H, align_corners, cost.device)
    t = t.view(-1, 1)
    x = x[:, :, i0] * (1 - t) + x[:, :, i1] * t
    i0, i1, t = _linear_taps(out_d, D, align_corners, cost.device)
    t = t.view(-1, 1, 1)
    x = x[:, i0] * (1 - t) + x[:, i1] * t
    return x


def disparity_regression(prob: torch.Tensor, maxdisp: int, keepdim: bool = False) -> torch.Tensor:
    """GwcNet/submodule.py:23-27 (keepdim=False); PSMNet/submodule.py:46-54 and
    IGEVStereo/submodule.py:221-225 (keepdim=True)."""
    assert prob.dim() == 4
    values = torch.arange(maxdisp, dtype=prob.dtype, device=prob.device).view(1, maxdisp, 1, 1)
    return (prob * values).sum(1, keepdim=keepdim)


def upsample_softargmin(cost: torch.Tensor, maxdisp: int, out_h: int, out_w: int,
                        align_corners: bool = False, keepdim: bool = False) -> torch.Tensor:
    """The fused head: GwcNet/gwcnet.py:220-223 / PSMNet/stackhourglass.py:150-156.
    ``cost`` is [B,1,D,H,W] (or [B,D,H,W]); returns [B,out_h,out_w] (or [B,1,..] with keepdim)."""
    if cost.dim() == 5:
        cost = cost[:, 0]
    up = trilinear_upsample(cost, maxdisp, out_h, out_w, align_corners)
    prob = torch.softmax(up, dim=1)
    return disparity_regression(prob, maxdisp, keepdim)


def disparity_variance(prob: torch.Tensor, maxdisp: int, disparity: torch.Tensor) -> torch.Tensor:
    """CFNet/submodule.py:127-133: sum_d prob[:, d] * (d - disparity)^2, keepdim.  prob [B,D,H,W], disparity [B,1,H,W]."""
    assert len(prob.shape) == 4
    d = torch.arange(0, maxdisp, dtype=prob.dtype).view(1, maxdisp, 1, 1)
    return torch.sum(prob * (d - disparity) ** 2, 1, keepdim=True)


def softargmin(cost: torch.Tensor, keepdim: bool = True) -> torch.Tensor:
    """IGEVStereo/igev_stereo.py:212-213 -- softmax over D and regression at the volume's own
    resolution (no upsampling). ``cost`` [B,D,H,W]."""
    prob = torch.softmax(cost, dim=1)
    return disparity_regression(prob, cost.shape[1], keepdim)


# --------------------------------------------------------------------------- 3-D conv family
def activation(x: torch.Tensor, act: str) -> torch.Tensor:
    if act == "none":
        return x
    if act == "relu":
        return torch.relu(x)
    if act == "leaky":  # IGEVStereo/submodule.py:36 -- LeakyReLU() default slope 0.01
        return F.leaky_relu(x, 0.01)
    if act == "mish":   # CFNet/submodule.py:99-106 -- x * tanh(softplus(x))
        return x * torch.tanh(F.softplus(x))
    raise ValueError(act)


def fold_bn(bn: Optional[dict], c_out: int, eps: float = 1e-5, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm3d as y = x*scale + shift (PSMNet/submodule.py:16-19)."""
    if bn is None:
        return torch.ones(c_out, device=device), torch.zeros(c_out, device=device)
    scale = bn["weight"] / torch.sqrt(bn["running_var"] + eps)
    shift = bn["bias"] - bn["running_mean"] * scale
    return scale, shift


def conv3d_bn_act(x, weight, bn=None, stride=1, padding=1, act="none", residual=None, transposed=False,
                  output_padding=0):
    """convbn_3d (+ReLU/Mish/LeakyReLU, + residual added BEFORE the activation), eval-mode BN:
    PSMNet/submodule.py:16-19, GwcNet/gwcnet.py:72-105; transposed: ConvTranspose3d(k3,s2,p1,op1)
    PSMNet/stackhourglass.py:25-29 and k4 s2 p1 IGEVStereo/igev_stereo.py:43-50."""
    if transposed:
        y = F.conv_transpose3d(x, weight, stride=stride, padding=padding, output_padding=output_padding)
        c_out = weight.shape[1]
    else:
        y = F.conv3d(x, weight, stride=stride, padding=padding)
        c_out = weight.shape[0]
    scale, shift = fold_bn(bn, c_out, device=y.device)
    y = y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if residual is not None:
        y = y + residual
    return activation(y, act)


# --------------------------------------------------------------------------- ACVNet pieces
def depthwise_patch(vol: torch.Tensor, weight: torch.Tensor, dilation: int) -> torch.Tensor:
    """Depthwise (1,3,3) 'patch' convolution of ACVNet (ACVNet/acv.py:109-112: Conv3d(C, C, (1,3,3), groups=C,
    dilation=d, padding=(0,d,d), bias=False)); written as 9 shifted multiply-adds.  vol [B,C,D,H,W], weight [C,1,1,3,3]."""
    B, C, D, H, W = vol.shape
    d = dilation
    xp = F.pad(vol, (d, d, d, d))
    out = torch.zeros_like(vol)
    for a in range(3):
        for b in range(3):
            out = out + weight[:, 0, 0, a, b].view(1, C, 1, 1, 1) * xp[:, :, :, a * d:a * d + H, b * d:b * d + W]
    return out


def block_attention_core(qkv: torch.Tensor, qkv_b: torch.Tensor, num_heads: int, block=(4, 4, 4)) -> torch.Tensor:
    """The part of attention_block.forward between the qkv Linear and final1x1 (ACVNet/submodule.py:392-428).
    qkv [B,3C,D,H0,W0] (channel order (3, heads, head_dim), i.e. the Linear's output order) -> [B,C,D,H0,W0].

    The reference zero-pads H and W on the bottom/right up to a multiple of the block BEFORE the qkv Linear, so padded
    tokens carry qkv = bias; here the qkv volume is padded with the bias instead.  The mask marks padded rows/columns;
    the reference writes ``mask[:, -pad_b:, :]`` / ``mask[:, :, -pad_r:]`` (:405-406), and a slice ``-0:`` selects
    EVERYTHING, so when exactly one of pad_b / pad_r is zero the whole mask is 1 and nothing is masked -- kept here on
    purpose."""
    B, C3, D, H0, W0 = qkv.shape
    C = C3 // 3
    b0, b1, b2 = block
    pad_r = (b2 - W0 % b2) % b2
    pad_b = (b1 - H0 % b1) % b1
    H, W = H0 + pad_b, W0 + pad_r
    full = qkv_b.view(1, C3, 1, 1, 1).expand(B, C3, D, H, W).clone()
    full[:, :, :, :H0, :W0] = qkv
    d, h, w = D // b0, H // b1, W // b2
    hd = C // num_heads
    nt = b0 * b1 * b2
    tok = full.view(B, C3, d, b0, h, b1, w, b2).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, d * h * w, nt, C3)
    tok = tok.view(B, d * h * w, nt, 3, num_heads, hd)
    q, k, v = (tok[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))      # [B,nb,heads,nt,hd]
    attn = torch.einsum("bnhie,bnhje->bnhij", q, k) * (hd ** -0.5)
    if pad_r > 0 or pad_b > 0:
        m = torch.zeros(H, W)
        m[(H - pad_b) if pad_b > 0 else 0:, :] = 1                           # '-0:' == whole axis
        m[:, (W - pad_r) if pad_r > 0 else 0:] = 1
        m = m.view(h, b1, w, b2).permute(0, 2, 1, 3).reshape(h * w, b1 * b2)  # per (h,w) block, per in-plane token
        diff = (m[:, :, None] != m[:, None, :]).float() * -1000.0             # [h*w, b1b2, b1b2]
        diff = diff.repeat(d, b0, b0)                                        # token order (b0,b1,b2); block order (d,h,w)
        attn = attn + diff.view(1, d * h * w, 1, nt, nt)
    attn = torch.softmax(attn, dim=-1)
    o = torch.einsum("bnhij,bnhje->bnhie", attn, v)                          # [B,nb,heads,nt,hd]
    o = o.view(B, d, h, w, num_heads, b0, b1, b2, hd).permute(0, 4, 8, 1, 5, 2, 6, 3, 7).reshape(B, C, D, H, W)
    return o[:, :, :, :H0, :W0]


def block_attention(x: torch.Tensor, qkv_w: torch.Tensor, qkv_b: torch.Tensor, fin_w: torch.Tensor,
                    fin_b: torch.Tensor, num_heads: int, block=(4, 4, 4)) -> torch.Tensor:
    """attention_block.forward (ACVNet/submodule.py:381-430): multi-head self-attention inside non-overlapping
    (b0,b1,b2) blocks of the volume, then a 1x1x1 conv with bias.  x [B,C,D,H0,W0].  The qkv Linear acts on the channel
    axis of every voxel (:392-400)."""
    C = x.shape[1]
    qkv = torch.einsum("oc,bcdhw->bodhw", qkv_w, x) + qkv_b.view(1, -1, 1, 1, 1)
    o = block_attention_core(qkv, qkv_b, num_heads, block)
    return torch.einsum("oc,bcdhw->bodhw", fin_w.view(C, C), o) + fin_b.view(1, C, 1, 1, 1)


# --------------------------------------------------------------------------- 1-D all-pairs correlation
def corr1d(fmap1: torch.Tensor, fmap2: torch.Tensor, scale: bool = True) -> torch.Tensor:
    """RAFTStereo/corr.py:148-156 (scale=True: divide by sqrt(C)); IGEVStereo/geometry.py:62-70
    (scale=False).  Returns [B,H,W1,W2]."""
    B, C, H, W1 = fmap1.shape
    corr = torch.einsum("bchw,bchv->bhwv", fmap1.float(), fmap2.float())
    if scale:
        corr = corr / math.sqrt(C)
    return corr


def avg_pool_last(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(x,[1,2],stride=[1,2]) on the last axis (RAFTStereo/corr.py:123-125):
    floor(W/2) outputs, an odd trailing element is dropped."""
    n = x.shape[-1] // 2
    return 0.5 * (x[..., 0:2 * n:2] + x[..., 1:2 * n:2])


def corr_pyramid(corr: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    """RAFTStereo/corr.py:119-125: level 0 plus ``num_levels`` pooled copies are stored
    (num_levels+1 entries); __call__ reads levels 0..num_levels-1 (corr.py:133)."""
    pyr = [corr]
    for _ in range(num_levels):
        corr = avg_pool_last(corr)
        pyr.append(corr)
    return pyr


def sample_1d_zero(vol: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """1-D linear sampling along the last axis with zero padding, pixel coordinates --
    what bilinear_sampler (RAFTStereo/utils/utils.py:59-74) does for an H==1 image through
    F.grid_sample(align_corners=True, padding_mode='zeros').

    vol [..., W]; x [..., K] (same leading dims) -> [..., K]
    """
    W = vol.shape[-1]
    x0 = torch.floor(x)
    f = x - x0
    i0 = x0.to(torch.int64)
    i1 = i0 + 1
    def tap(i):
        ok = (i >= 0) & (i < W)
        v = torch.gather(vol, -1, i.clamp(0, W - 1))
        return torch.where(ok, v, torch.zeros_like(v))
    return tap(i0) * (1 - f) + tap(i1) * f


def corr_lookup(pyramid: Sequence[torch.Tensor], coords_x: torch.Tensor, radius: int, num_levels: int) -> torch.Tensor:
    """CorrBlock1D.__call__ (RAFTStereo/corr.py:127-146).
    pyramid[i] [B,H,W1,W2/2^i]; coords_x [B,H,W1] -> [B, num_levels*(2r+1), H, W1] fp32."""
    dx = torch.arange(-radius, radius + 1, dtype=coords_x.dtype)
    outs = []
    for i in range(num_levels):
        x = coords_x[..., None] / (2 ** i) + dx
        outs.append(sample_1d_zero(pyramid[i], x))
    out = torch.cat(outs, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


def geo_pyramids(fmap1, fmap2, geo_volume, num_levels: int):
    """Combined_Geo_Encoding_Volume.__init__ (IGEVStereo/geometry.py:8-30).
    geo_volume [B,C,D,H,W] -> list of [B,H,W,C,D/2^i]; corr -> list of [B,H,W,W2/2^i]."""
    corr = corr1d(fmap1, fmap2, scale=False)
    geo = geo_volume.permute(0, 3, 4, 1, 2).contiguous()
    geos, corrs = [geo], [corr]
    for _ in range(num_levels - 1):
        geo = avg_pool_last(geo)
        geos.append(geo)
    for _ in range(num_levels - 1):
        corr = avg_pool_last(corr)
        corrs.append(corr)
    return geos, corrs


def geo_lookup(geos, corrs, disp: torch.Tensor, coords_x: torch.Tensor, radius: int) -> torch.Tensor:
    """Combined_Geo_Encoding_Volume.__call__ (IGEVStereo/geometry.py:35-59).
    disp, coords_x [B,H,W] -> [B, L*(2r+1)*(C+1), H, W]; per level: C*(2r+1) geo taps
    (channel-major, tap-minor) then (2r+1) correlation taps."""
    dx = torch.arange(-radius, radius + 1, dtype=disp.dtype)
    outs = []
    for i, (geo, corr) in enumerate(zip(geos, corrs)):
        B, H, W, C, D = geo.shape
        x = (disp[..., None] / (2 ** i) + dx)                   # [B,H,W,K]
        g = sample_1d_zero(geo, x[..., None, :].expand(B, H, W, C, x.shape[-1]))
        outs.append(g.reshape(B, H, W, -1))
        xc = coords_x[..., None] / (2 ** i) - disp[..., None] / (2 ** i) + dx
        outs.append(sample_1d_zero(corr, xc))
    out = torch.cat(outs, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


class CorrBlock1D:
    """CPU stand-in with the interface of RAFTStereo/corr.py:110-156, built from the restatements above; tests swap
    it into the drop-in RAFTStereo to obtain an oracle for the whole iterative model."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels, self.radius = num_levels, radius
        self.pyramid = corr_pyramid(corr1d(fmap1, fmap2, True), num_levels)

    def __call__(self, coords):
        return corr_lookup(self.pyramid, coords[:, 0], self.radius, self.num_levels)


# --------------------------------------------------------------------------- "next" rows of SURVEY 8f: learned upsampling
def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, factor: int) -> torch.Tensor:
    """RAFTStereo.upsample_flow (RAFTStereo/raft_stereo.py:81-93): every fine pixel (fy, fx) of coarse pixel (h, w) is a convex
    combination -- softmax over the 9 mask logits of that fine pixel -- of the 3x3 coarse neighbourhood of ``factor * flow``
    (zero outside the image).  flow [N,D,H,W], mask [N, 9*factor^2, H, W] (channel = tap*factor^2 + fy*factor + fx)
    -> [N, D, factor*H, factor*W].  Written as explicit loops over the 9 taps."""
    N, D, H, W = flow.shape
    f = factor
    w = torch.softmax(mask.view(N, 9, f, f, H, W), dim=1)
    padded = F.pad(f * flow, (1, 1, 1, 1))
    out = torch.zeros(N, D, f, f, H, W, dtype=flow.dtype)
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        out = out + w[:, tap][:, None] * padded[:, :, dy:dy + H, dx:dx + W][:, :, None, None]
    return out.permute(0, 1, 4, 2, 5, 3).reshape(N, D, f * H, f * W)


def context_upsample(disp_low: torch.Tensor, up_weights: torch.Tensor) -> torch.Tensor:
    """IGEVStereo/submodule.py:243-255: fine pixel (Y, X) takes the 3x3 neighbourhood of its coarse parent (Y//4, X//4)
    (zero outside), weighted by its own 9 weights (already soft-maxed by the caller).  disp_low [B,1,h,w],
    up_weights [B,9,4h,4w] -> [B,4h,4w]."""
    B, _, h, w = disp_low.shape
    padded = F.pad(disp_low[:, 0], (1, 1, 1, 1))
    out = torch.zeros(B, 4 * h, 4 * w, dtype=disp_low.dtype)
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        nb = padded[:, dy:dy + h, dx:dx + w]
        out = out + up_weights[:, tap] * nb.repeat_interleave(4, dim=1).repeat_interleave(4, dim=2)
    return out


# --------------------------------------------------------------------------- CFNet sampled volume
def sampled_volume(gw_left, gw_right, cat_left, cat_right, samples, num_groups: int) -> torch.Tensor:
    """CFNet cascade-stage volume (CFNet/cfnet.py:545-550): cat of cost_volume_generator(..., 'gwc') (:472-496 with
    groupwise_correlation_4D, submodule.py:162-168), cost_volume_generator(..., 'concat') and the sample channel.
    SpatialTransformer (submodule.py:302-349): the right feature at x = w - sample, x clamped to [0, W-1] for the gather,
    the gathered value zeroed where w - sample is outside [0, W-1]; the left feature broadcast over the samples.
    gw_* [B,Cg,H,W], cat_* [B,Cc,H,W], samples [B,S,H,W] (integer valued) -> [B, G + 2*Cc + 1, S, H, W]."""
    B, Cg, H, W = gw_left.shape
    S = samples.shape[1]
    coord = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) - samples.float()            # [B,S,H,W]
    valid = ~((coord < 0) | (coord > W - 1))
    x = coord.clamp(0, W - 1).long()

    def warp(right):                                   # [B,C,H,W] -> [B,C,S,H,W]
        C = right.shape[1]
        out = torch.zeros(B, C, S, H, W, dtype=right.dtype)
        for s in range(S):
            idx = x[:, s][:, None].expand(B, C, H, W)
            out[:, :, s] = torch.gather(right, 3, idx) * valid[:, s][:, None]
        return out

    k = Cg // num_groups
    prod = gw_left[:, :, None] * warp(gw_right)                                               # [B,Cg,S,H,W]
    gwc = prod.view(B, num_groups, k, S, H, W).mean(dim=2)
    left = cat_left[:, :, None].expand(B, cat_left.shape[1], S, H, W)
    return torch.cat((gwc, left, warp(cat_right), samples.float()[:, None]), dim=1)
