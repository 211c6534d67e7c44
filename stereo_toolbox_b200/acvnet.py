"""ACVNet drop-in (reference: models/ACVNet/acv.py:95-250, hourglass-with-attention :54-93,
attention_block models/ACVNet/submodule.py:366-430).

Same constructor (``ACVNet(maxdisp, attn_weights_only, freeze_attn_weights)``), same ``forward(left, right)``
contract in eval mode and the same state-dict names/shapes.  Everything after the 2-D extractor runs in
libstb200.so: gwc volume, the depthwise patch convolutions, the attention-weight branch, softmax over D fused
into the concat-volume build, the two attention hourglasses (block self-attention kernel + 1x1x1 convs on the
conv family), classif2 and the trilinear soft-argmin head.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .aggregation import TrainBackend, convbn_3d, deconvbn_3d, make_backend
from .features2d import GwcFeatures, convbn
from .gwcnet import GwcNet, _classif


def block_attention_train(qkv, qkv_bias, num_heads, block):
    """Training-path form of the windowed attention core (ACVNet/submodule.py:392-428) as torch ops, so that autograd
    provides the adjoint (the inference path runs csrc/acv.cu; an adjoint kernel is not built).  qkv [B,3C,D,H,W] with
    channel order (3, heads, head_dim) -> [B,C,D,H,W].  Padding semantics of the reference: H and W are padded up to a
    multiple of the window with tokens whose q,k,v equal the bias; padded and real tokens are masked from each other only
    when BOTH pads are non-zero -- with exactly one pad the reference's ``mask[:, -0:, :] = 1`` marks every token."""
    B, C3, D, H0, W0 = qkv.shape
    C, (b0, b1, b2) = C3 // 3, block
    ph, pw = (-H0) % b1, (-W0) % b2
    if ph or pw:
        canvas = qkv_bias.to(qkv.dtype).view(1, C3, 1, 1, 1).expand(B, C3, D, H0 + ph, W0 + pw).clone()
        canvas[:, :, :, :H0, :W0] = qkv
        qkv = canvas
    H, W = H0 + ph, W0 + pw
    nd, nh, nw, hd, T = D // b0, H // b1, W // b2, C // num_heads, b0 * b1 * b2
    win = qkv.view(B, 3, num_heads, hd, nd, b0, nh, b1, nw, b2).permute(1, 0, 4, 6, 8, 2, 5, 7, 9, 3)
    q, k, v = win.reshape(3, B, nd * nh * nw, num_heads, T, hd).unbind(0)
    logits = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5                       # [B,windows,heads,T,T]
    if ph and pw:
        m = torch.zeros(H, W, device=qkv.device)
        m[H - ph:, :] = 1
        m[:, W - pw:] = 1
        m = m.view(nh, b1, nw, b2).permute(0, 2, 1, 3).reshape(nh * nw, 1, b1 * b2).expand(-1, b0, -1).reshape(nh * nw, T)
        sep = (m[:, :, None] != m[:, None, :]).to(logits.dtype) * -1000.0            # [nh*nw, T, T]
        logits = logits + sep.repeat(nd, 1, 1)[None, :, None]
    out = torch.matmul(torch.softmax(logits, dim=-1), v)                             # [B,windows,heads,T,hd]
    out = out.view(B, nd, nh, nw, num_heads, b0, b1, b2, hd).permute(0, 4, 8, 1, 5, 2, 6, 3, 7).reshape(B, C, D, H, W)
    return out[:, :, :, :H0, :W0]


class attention_block(nn.Module):
    """Parameter container of ACVNet/submodule.py:366-379; ``run`` is forward :381-430 on a backend."""

    def __init__(self, channels_3d, num_heads=8, block=(4, 4, 4)):
        super().__init__()
        self.block = tuple(block)
        self.dim_3d = channels_3d
        self.num_heads = num_heads
        self.qkv_3d = nn.Linear(channels_3d, channels_3d * 3, bias=True)
        self.final1x1 = nn.Conv3d(channels_3d, channels_3d, 1)

    def _qkv_conv(self) -> nn.Conv3d:
        """The qkv Linear acts on the channel axis of every voxel: a 1x1x1 convolution with bias.  A parameter-sharing
        Conv3d view of it (not registered: the state dict keeps the reference's ``qkv_3d.*`` names only)."""
        w, b = self.qkv_3d.weight, self.qkv_3d.bias
        tag = (w.data_ptr(), w._version, b.data_ptr(), b._version)
        hit = self.__dict__.get("_qkv_shadow")
        if hit is None or hit[0] != tag:
            c = nn.Conv3d(self.dim_3d, 3 * self.dim_3d, 1, bias=True)
            c.weight = nn.Parameter(w.detach().view(3 * self.dim_3d, self.dim_3d, 1, 1, 1), requires_grad=False)
            c.bias = nn.Parameter(b.detach(), requires_grad=False)
            hit = (tag, c)
            self.__dict__["_qkv_shadow"] = hit
        return hit[1]

    def run(self, be, x):
        if getattr(be, "training_path", False):           # training: qkv Linear as a differentiable 1x1x1 conv, torch attention core
            w = self.qkv_3d.weight
            qkv = torch.nn.functional.conv3d(x, w.view(w.shape[0], w.shape[1], 1, 1, 1), self.qkv_3d.bias)
            return be.conv(self.final1x1, block_attention_train(qkv, self.qkv_3d.bias, self.num_heads, self.block))
        qkv = be.conv(self._qkv_conv(), x)
        o = be.block_attention(qkv, self.qkv_3d.bias, self.num_heads, self.block)
        return be.conv(self.final1x1, o)


class hourglass(nn.Module):
    """ACVNet/acv.py:54-93."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(convbn_3d(c, c * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv3 = nn.Sequential(convbn_3d(c * 2, c * 4, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(c * 4, c * 4, 3, 1, 1), nn.ReLU(inplace=True))
        self.attention_block = attention_block(channels_3d=c * 4, num_heads=16, block=(4, 4, 4))
        self.conv5 = deconvbn_3d(c * 4, c * 2)
        self.conv6 = deconvbn_3d(c * 2, c)
        self.redir1 = convbn_3d(c, c, 1, 1, 0)
        self.redir2 = convbn_3d(c * 2, c * 2, 1, 1, 0)

    def run(self, be, x):
        c1 = be.conv(self.conv1[0], x, "relu")
        c2 = be.conv(self.conv2[0], c1, "relu")
        c3 = be.conv(self.conv3[0], c2, "relu")
        c4 = be.conv(self.conv4[0], c3, "relu")
        c4 = self.attention_block.run(be, c4)
        r2 = be.conv(self.redir2, c2)
        c5 = be.conv(self.conv5, c4, "relu", residual=r2)
        r1 = be.conv(self.redir1, x)
        return be.conv(self.conv6, c5, "relu", residual=r1)


def _patch(c, dil):
    return nn.Conv3d(c, c, kernel_size=(1, 3, 3), stride=1, dilation=dil, groups=c, padding=(0, dil, dil), bias=False)


class ACVNet(nn.Module):
    def __init__(self, maxdisp=192, attn_weights_only=False, freeze_attn_weights=False, precision="auto"):
        super().__init__()
        self.maxdisp = maxdisp
        self.attn_weights_only = attn_weights_only
        self.freeze_attn_weights = freeze_attn_weights
        self.num_groups = 40
        self.concat_channels = 32
        self.feature_extraction = GwcFeatures(False, 0)
        self.concatconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                        nn.Conv2d(128, self.concat_channels, kernel_size=1, padding=0, stride=1, bias=False))
        self.patch = _patch(40, 1)
        self.patch_l1 = _patch(8, 1)
        self.patch_l2 = _patch(16, 2)
        self.patch_l3 = _patch(16, 3)
        self.dres1_att_ = nn.Sequential(convbn_3d(40, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
        self.dres2_att_ = hourglass(32)
        self.classif_att_ = _classif()
        self.dres0 = nn.Sequential(convbn_3d(self.concat_channels * 2, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                   convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.classif0, self.classif1, self.classif2 = _classif(), _classif(), _classif()
        self.feature_tf32 = None
        self.feature_mode = None
        self.set_precision(precision)

    set_precision = GwcNet.set_precision
    _resolve_auto_precision = GwcNet._resolve_auto_precision
    _features = GwcNet._features

    def _concat_features(self, fl, fr):
        """concatconv (ACVNet/acv.py:192-193) on the gwc features; on the 16-bit path the tensor-core extractor already
        produced it (features_umma.py, ``concat_head``)."""
        if "concat_feature" in fl:
            return fl["concat_feature"], fr["concat_feature"]
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and self._feature_mode_resolved() != "fp32"
        try:
            return self.concatconv(fl["gwc_feature"]), self.concatconv(fr["gwc_feature"])
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def _feature_mode_resolved(self):
        if self.feature_tf32 is False:
            return "fp32"
        return self.feature_mode or ("fp32" if self.precision == "fp32" else "umma")

    def attention_weights(self, fl, fr):
        """acv.py:180-190: gwc volume -> patch convs -> dres1_att_ -> hourglass -> classif_att_.
        Returns [B,1,D/4,H/4,W/4] fp32."""
        be = self._be
        D4 = self.maxdisp // 4
        vol = ops.gwc_volume(fl["gwc_feature"], fr["gwc_feature"], D4, self.num_groups)
        v1 = ops.patch_dw(vol, self.patch.weight, 1)
        pv = torch.empty_like(v1)                       # the three slices cover all 40 channels (replaces torch.cat)
        ops.patch_dw(v1, self.patch_l1.weight, 1, out=pv, c_off=0)
        ops.patch_dw(v1, self.patch_l2.weight, 2, out=pv, c_off=8)
        ops.patch_dw(v1, self.patch_l3.weight, 3, out=pv, c_off=24)
        x = be.from_ncdhw(pv)
        c = be.conv(self.dres1_att_[0], x, "relu")
        c = be.conv(self.dres1_att_[2], c)
        c = self.dres2_att_.run(be, c)
        c = be.conv(self.classif_att_[0], c, "relu")
        return be.cost_ncdhw(be.conv(self.classif_att_[2], c))

    def aggregate(self, fl, fr, height, width):
        be = self._be
        att = self.attention_weights(fl, fr)
        self._last_att = att
        if self.attn_weights_only:
            return be.head(be.cost_native(att), self.maxdisp, height, width, align_corners=False)
        cl, cr = self._concat_features(fl, fr)
        prob = ops.softmax_d(att)                                            # F.softmax(att_weights, dim=2)
        ac = ops.concat_volume(cl, cr, self.maxdisp // 4, mask_left=False, att_prob=prob)
        x = be.from_ncdhw(ac)
        c = be.conv(self.dres0[0], x, "relu")
        cost0 = be.conv(self.dres0[2], c, "relu")
        c = be.conv(self.dres1[0], cost0, "relu")
        cost0 = be.conv(self.dres1[2], c, "none", residual=cost0)
        out1 = self.dres2.run(be, cost0)
        out2 = self.dres3.run(be, out1)
        c = be.conv(self.classif2[0], out2, "relu")
        cost2 = be.conv(self.classif2[2], c)
        self._last_cost = cost2
        return be.head(cost2, self.maxdisp, height, width, align_corners=False)

    def forward(self, left, right):
        if self.training:
            return self._forward_train(left, right)
        self._resolve_auto_precision(left)
        fl, fr = self._features(left, right)
        return self.aggregate(fl, fr, left.shape[2], left.shape[3])

    def _forward_train(self, left, right):
        """Training step forward (exact fp32; acv.py:162-235).  3-D convolutions, volumes and heads on TrainBackend (forward
        and backward kernels of libstb200.so, batch-statistic BatchNorm3d); the pieces without an adjoint kernel -- the
        depthwise patch convs, the windowed attention core, softmax over D times the concat volume -- as torch ops.
        Returns [pred_attention, pred0, pred1, pred2] / [pred0, pred1, pred2] (freeze_attn_weights) / [pred_attention]
        (attn_weights_only)."""
        import contextlib
        be = TrainBackend()
        H, W = left.shape[2:]
        D4 = self.maxdisp // 4
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            with (torch.no_grad() if self.freeze_attn_weights else contextlib.nullcontext()):        # acv.py:164-178
                fl, fr = self.feature_extraction(left), self.feature_extraction(right)
                vol = be.volume_gwc_concat(fl["gwc_feature"], fr["gwc_feature"], None, None, D4, self.num_groups)
                v1 = self.patch(vol)
                pv = torch.cat((self.patch_l1(v1[:, :8]), self.patch_l2(v1[:, 8:24]), self.patch_l3(v1[:, 24:40])), dim=1)
                c = be.conv(self.dres1_att_[2], be.conv(self.dres1_att_[0], pv, "relu"))
                c = self.dres2_att_.run(be, c)
                att = be.conv(self.classif_att_[2], be.conv(self.classif_att_[0], c, "relu"))    # [B,1,D/4,H/4,W/4]
            preds = []
            if not self.freeze_attn_weights:
                preds.append(be.head(att, self.maxdisp, H, W, align_corners=False))
            if self.attn_weights_only:
                return preds
            cl, cr = self.concatconv(fl["gwc_feature"]), self.concatconv(fr["gwc_feature"])
            ac = torch.softmax(att, dim=2) * be.volume_concat(cl, cr, D4, mask_left=False)       # acv.py:196
            cost0 = be.conv(self.dres0[2], be.conv(self.dres0[0], ac, "relu"), "relu")
            cost0 = be.conv(self.dres1[2], be.conv(self.dres1[0], cost0, "relu"), "none", residual=cost0)
            out1 = self.dres2.run(be, cost0)
            out2 = self.dres3.run(be, out1)
            for cls, t in ((self.classif0, cost0), (self.classif1, out1), (self.classif2, out2)):
                preds.append(be.head(be.conv(cls[2], be.conv(cls[0], t, "relu")), self.maxdisp, H, W, align_corners=False))
            return preds
        finally:
            torch.backends.cudnn.allow_tf32 = prev
