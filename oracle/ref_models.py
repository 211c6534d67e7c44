"""Oracle restatement of the reference MODELS as pure functions of a state dict (CPU torch).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  These follow the reference's eval-mode
``forward`` line by line but hold no modules: every layer is an ``F.*`` call on tensors taken
from the state dict by the reference's own parameter names, so a passing comparison also
proves the CUDA-backed models keep the reference's state-dict layout.

  gwcnet_forward : GwcNet/gwcnet.py:171-224 (+ feature_extraction :12-65, hourglass :68-105)
  psmnet_forward : PSMNet/stackhourglass.py:103-161 (+ feature_extraction PSMNet/submodule.py:57-132)
  acvnet_forward : ACVNet/acv.py:159-250 eval branch (+ hourglass-with-attention :54-93, attention_block submodule.py:366-430)
  igev_cost_volume : IGEVStereo/igev_stereo.py:205-213 (+ hourglass :23-90, BasicConv / FeatureAtt submodule.py:9-37,228-241)
  mish_hourglass / mish_hourglassup : CFNet/cfnet.py:178-271, PCWNet/pcwnet.py:133-251
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ref_ops as R

SD = Dict[str, torch.Tensor]


def _bn(sd: SD, p: str) -> dict:
    return {k: sd[f"{p}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}


# ------------------------------------------------------------------ 2-D pieces (not hot path)
def convbn2d(sd: SD, p: str, x, stride=1, pad=1, dilation=1):
    """convbn (GwcNet/submodule.py:11-14): padding = dilation if dilation > 1 else pad."""
    y = F.conv2d(x, sd[f"{p}.0.weight"], stride=stride, padding=dilation if dilation > 1 else pad, dilation=dilation)
    b = _bn(sd, f"{p}.1")
    return F.batch_norm(y, b["running_mean"], b["running_var"], b["weight"], b["bias"], False, 0.0, 1e-5)


def basic_block(sd: SD, p: str, x, stride, pad, dilation):
    """BasicBlock (GwcNet/submodule.py:66-91): no ReLU after the residual add."""
    out = torch.relu(convbn2d(sd, f"{p}.conv1.0", x, stride, pad, dilation))
    out = convbn2d(sd, f"{p}.conv2", out, 1, pad, dilation)
    if f"{p}.downsample.0.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.downsample.0.weight"], stride=stride)
        b = _bn(sd, f"{p}.downsample.1")
        x = F.batch_norm(x, b["running_mean"], b["running_var"], b["weight"], b["bias"], False, 0.0, 1e-5)
    return out + x


def res_layer(sd: SD, p: str, x, blocks, stride, pad, dilation):
    for i in range(blocks):
        x = basic_block(sd, f"{p}.{i}", x, stride if i == 0 else 1, pad, dilation)
    return x


def backbone(sd: SD, p: str, x):
    x = torch.relu(convbn2d(sd, f"{p}.firstconv.0", x, 2))
    x = torch.relu(convbn2d(sd, f"{p}.firstconv.2", x, 1))
    x = torch.relu(convbn2d(sd, f"{p}.firstconv.4", x, 1))
    x = res_layer(sd, f"{p}.layer1", x, 3, 1, 1, 1)
    l2 = res_layer(sd, f"{p}.layer2", x, 16, 2, 1, 1)
    l3 = res_layer(sd, f"{p}.layer3", l2, 3, 1, 1, 1)
    l4 = res_layer(sd, f"{p}.layer4", l3, 3, 1, 1, 2)
    return l2, l3, l4


def gwc_features(sd: SD, x, concat: bool):
    """feature_extraction.forward (GwcNet/gwcnet.py:52-65)."""
    p = "feature_extraction"
    l2, l3, l4 = backbone(sd, p, x)
    gwc = torch.cat((l2, l3, l4), 1)
    if not concat:
        return gwc, None
    y = torch.relu(convbn2d(sd, f"{p}.lastconv.0", gwc, 1))
    return gwc, F.conv2d(y, sd[f"{p}.lastconv.2.weight"])


def psm_features(sd: SD, x):
    """feature_extraction.forward (PSMNet/submodule.py:107-132), SPP branches 64/32/16/8."""
    p = "feature_extraction"
    l2, _, l4 = backbone(sd, p, x)
    h, w = l4.shape[2:]
    br = {}
    for i, k in ((1, 64), (2, 32), (3, 16), (4, 8)):
        y = F.avg_pool2d(l4, (k, k), stride=(k, k))
        y = torch.relu(convbn2d(sd, f"{p}.branch{i}.1", y, 1, 0))
        br[i] = F.interpolate(y, (h, w), mode="bilinear", align_corners=False)
    feat = torch.cat((l2, l4, br[4], br[3], br[2], br[1]), 1)
    y = torch.relu(convbn2d(sd, f"{p}.lastconv.0", feat, 1))
    return F.conv2d(y, sd[f"{p}.lastconv.2.weight"])


# ------------------------------------------------------------------ 3-D aggregation (hot path)
def cbn3(sd: SD, p: str, x, stride=1, pad=1, act="none", residual=None):
    """convbn_3d (+act): Sequential index 0 = Conv3d, 1 = BatchNorm3d."""
    return R.conv3d_bn_act(x, sd[f"{p}.0.weight"], _bn(sd, f"{p}.1"), stride, pad, act, residual)


def dbn3(sd: SD, p: str, x, act="none", residual=None):
    """ConvTranspose3d(k3,s2,p1,op1)+BN3d (+residual, act)."""
    return R.conv3d_bn_act(x, sd[f"{p}.0.weight"], _bn(sd, f"{p}.1"), 2, 1, act, residual, transposed=True,
                           output_padding=1)


def gwc_hourglass(sd: SD, p: str, x):
    """hourglass.forward (GwcNet/gwcnet.py:95-105)."""
    c1 = cbn3(sd, f"{p}.conv1.0", x, 2, 1, "relu")
    c2 = cbn3(sd, f"{p}.conv2.0", c1, 1, 1, "relu")
    c3 = cbn3(sd, f"{p}.conv3.0", c2, 2, 1, "relu")
    c4 = cbn3(sd, f"{p}.conv4.0", c3, 1, 1, "relu")
    r2 = cbn3(sd, f"{p}.redir2", c2, 1, 0)
    c5 = dbn3(sd, f"{p}.conv5", c4, "relu", r2)
    r1 = cbn3(sd, f"{p}.redir1", x, 1, 0)
    return dbn3(sd, f"{p}.conv6", c5, "relu", r1)


def psm_hourglass(sd: SD, p: str, x, presqu, postsqu):
    """hourglass.forward (PSMNet/stackhourglass.py:31-50)."""
    out = cbn3(sd, f"{p}.conv1.0", x, 2, 1, "relu")
    pre = cbn3(sd, f"{p}.conv2", out, 1, 1, "relu", postsqu)
    out = cbn3(sd, f"{p}.conv3.0", pre, 2, 1, "relu")
    out = cbn3(sd, f"{p}.conv4.0", out, 1, 1, "relu")
    post = dbn3(sd, f"{p}.conv5", out, "relu", presqu if presqu is not None else pre)
    out = dbn3(sd, f"{p}.conv6", post)
    return out, pre, post


def classif(sd: SD, p: str, x):
    y = cbn3(sd, f"{p}.0", x, 1, 1, "relu")
    return F.conv3d(y, sd[f"{p}.2.weight"], padding=1)


@torch.no_grad()
def gwcnet_forward(sd: SD, left, right, maxdisp: int, use_concat: bool, return_aux: bool = False):
    gl, cl = gwc_features(sd, left, use_concat)
    gr, cr = gwc_features(sd, right, use_concat)
    vol = R.build_gwc_volume(gl, gr, maxdisp // 4, 40)
    if use_concat:
        vol = torch.cat((vol, R.build_concat_volume(cl, cr, maxdisp // 4, mask_left=True)), 1)
    c = cbn3(sd, "dres0.0", vol, 1, 1, "relu")
    cost0 = cbn3(sd, "dres0.2", c, 1, 1, "relu")
    c = cbn3(sd, "dres1.0", cost0, 1, 1, "relu")
    cost0 = cbn3(sd, "dres1.2", c, 1, 1, "none", cost0)
    out1 = gwc_hourglass(sd, "dres2", cost0)
    out2 = gwc_hourglass(sd, "dres3", out1)
    out3 = gwc_hourglass(sd, "dres4", out2)
    cost3 = classif(sd, "classif3", out3)
    disp = R.upsample_softargmin(cost3, maxdisp, left.shape[2], left.shape[3], False, False)
    if return_aux:
        return disp, dict(cost3=cost3, volume=vol, features=(gl, gr, cl, cr))
    return disp


@torch.no_grad()
def psmnet_forward(sd: SD, left, right, maxdisp: int, return_aux: bool = False):
    fl, fr = psm_features(sd, left), psm_features(sd, right)
    vol = R.build_concat_volume(fl, fr, maxdisp // 4, mask_left=True)   # inline loop :111-120
    c = cbn3(sd, "dres0.0", vol, 1, 1, "relu")
    cost0 = cbn3(sd, "dres0.2", c, 1, 1, "relu")
    c = cbn3(sd, "dres1.0", cost0, 1, 1, "relu")
    cost0 = cbn3(sd, "dres1.2", c, 1, 1, "none", cost0)
    out1, pre1, post1 = psm_hourglass(sd, "dres2", cost0, None, None)
    out1 = out1 + cost0
    out2, pre2, post2 = psm_hourglass(sd, "dres3", out1, pre1, post1)
    out2 = out2 + cost0
    out3, pre3, post3 = psm_hourglass(sd, "dres4", out2, pre1, post2)
    out3 = out3 + cost0
    cost1 = classif(sd, "classif1", out1)
    cost2 = classif(sd, "classif2", out2) + cost1
    cost3 = classif(sd, "classif3", out3) + cost2
    disp = R.upsample_softargmin(cost3, maxdisp, left.shape[2], left.shape[3], False, True)
    if return_aux:
        return disp, dict(cost3=cost3, volume=vol, features=(fl, fr))
    return disp


# --------------------------------------------------------------------------- ACVNet
def acv_hourglass(sd: SD, p: str, x):
    """hourglass.forward (ACVNet/acv.py:84-93): GwcNet's hourglass with the block attention after conv4."""
    c1 = cbn3(sd, f"{p}.conv1.0", x, 2, 1, "relu")
    c2 = cbn3(sd, f"{p}.conv2.0", c1, 1, 1, "relu")
    c3 = cbn3(sd, f"{p}.conv3.0", c2, 2, 1, "relu")
    c4 = cbn3(sd, f"{p}.conv4.0", c3, 1, 1, "relu")
    a = f"{p}.attention_block"
    c4 = R.block_attention(c4, sd[f"{a}.qkv_3d.weight"], sd[f"{a}.qkv_3d.bias"], sd[f"{a}.final1x1.weight"],
                           sd[f"{a}.final1x1.bias"], 16, (4, 4, 4))
    r2 = cbn3(sd, f"{p}.redir2", c2, 1, 0)
    c5 = dbn3(sd, f"{p}.conv5", c4, "relu", r2)
    r1 = cbn3(sd, f"{p}.redir1", x, 1, 0)
    return dbn3(sd, f"{p}.conv6", c5, "relu", r1)


@torch.no_grad()
def acvnet_forward(sd: SD, left, right, maxdisp: int, attn_weights_only: bool = False, return_aux: bool = False):
    """ACVNet.forward, eval mode (ACVNet/acv.py:159-250)."""
    p = "feature_extraction"
    gl = torch.cat(backbone(sd, p, left), 1)
    gr = torch.cat(backbone(sd, p, right), 1)
    vol = R.build_gwc_volume(gl, gr, maxdisp // 4, 40)
    vol = R.depthwise_patch(vol, sd["patch.weight"], 1)
    pv = torch.cat((R.depthwise_patch(vol[:, :8], sd["patch_l1.weight"], 1),
                    R.depthwise_patch(vol[:, 8:24], sd["patch_l2.weight"], 2),
                    R.depthwise_patch(vol[:, 24:40], sd["patch_l3.weight"], 3)), 1)
    c = cbn3(sd, "dres1_att_.0", pv, 1, 1, "relu")
    c = cbn3(sd, "dres1_att_.2", c, 1, 1)
    c = acv_hourglass(sd, "dres2_att_", c)
    att = classif(sd, "classif_att_", c)                                      # [B,1,D/4,H/4,W/4]
    H, W = left.shape[2:]
    if attn_weights_only:
        return R.upsample_softargmin(att, maxdisp, H, W, False, False)
    cat = lambda g: F.conv2d(torch.relu(convbn2d(sd, "concatconv.0", g, 1)), sd["concatconv.2.weight"])
    cv = R.build_concat_volume(cat(gl), cat(gr), maxdisp // 4, mask_left=False)   # ACVNet/submodule.py:179-191: left unmasked
    ac = R.attention_weighted_volume(att, cv)
    c = cbn3(sd, "dres0.0", ac, 1, 1, "relu")
    cost0 = cbn3(sd, "dres0.2", c, 1, 1, "relu")
    c = cbn3(sd, "dres1.0", cost0, 1, 1, "relu")
    cost0 = cbn3(sd, "dres1.2", c, 1, 1, "none", cost0)
    out1 = acv_hourglass(sd, "dres2", cost0)
    out2 = acv_hourglass(sd, "dres3", out1)
    cost2 = classif(sd, "classif2", out2)
    disp = R.upsample_softargmin(cost2, maxdisp, H, W, False, False)
    if return_aux:
        return disp, dict(cost2=cost2, att=att, patch_volume=pv, features=(gl, gr))
    return disp


# ------------------------------------------------------------------ IGEV cost-volume stage (hot path)
def igev_basic3d(sd: SD, p: str, x, stride=1, pad=1, act="leaky", transposed=False, bn=True):
    """BasicConv(is_3d=True): IGEVStereo/submodule.py:9-37 (conv -> bn -> LeakyReLU(0.01))."""
    return R.conv3d_bn_act(x, sd[f"{p}.conv.weight"], _bn(sd, f"{p}.bn") if bn else None, stride, pad, act, None,
                           transposed=transposed)


def igev_feature_att(sd: SD, p: str, cv, feat):
    """FeatureAtt.forward: IGEVStereo/submodule.py:236-241."""
    y = F.conv2d(feat, sd[f"{p}.feat_att.0.conv.weight"])
    b = _bn(sd, f"{p}.feat_att.0.bn")
    y = F.batch_norm(y, b["running_mean"], b["running_var"], b["weight"], b["bias"], False, 0.0, 1e-5)
    y = F.leaky_relu(y, 0.01)
    y = F.conv2d(y, sd[f"{p}.feat_att.1.weight"], sd[f"{p}.feat_att.1.bias"])
    return torch.sigmoid(y).unsqueeze(2) * cv


def igev_hourglass(sd: SD, p: str, x, features):
    """hourglass.forward: IGEVStereo/igev_stereo.py:66-90."""
    c1 = igev_basic3d(sd, f"{p}.conv1.1", igev_basic3d(sd, f"{p}.conv1.0", x, 2))
    c1 = igev_feature_att(sd, f"{p}.feature_att_8", c1, features[1])
    c2 = igev_basic3d(sd, f"{p}.conv2.1", igev_basic3d(sd, f"{p}.conv2.0", c1, 2))
    c2 = igev_feature_att(sd, f"{p}.feature_att_16", c2, features[2])
    c3 = igev_basic3d(sd, f"{p}.conv3.1", igev_basic3d(sd, f"{p}.conv3.0", c2, 2))
    c3 = igev_feature_att(sd, f"{p}.feature_att_32", c3, features[3])
    c3u = igev_basic3d(sd, f"{p}.conv3_up", c3, 2, 1, transposed=True)
    c2 = torch.cat((c3u, c2), 1)
    c2 = igev_basic3d(sd, f"{p}.agg_0.0", c2, 1, 0)
    c2 = igev_basic3d(sd, f"{p}.agg_0.2", igev_basic3d(sd, f"{p}.agg_0.1", c2))
    c2 = igev_feature_att(sd, f"{p}.feature_att_up_16", c2, features[2])
    c2u = igev_basic3d(sd, f"{p}.conv2_up", c2, 2, 1, transposed=True)
    c1 = torch.cat((c2u, c1), 1)
    c1 = igev_basic3d(sd, f"{p}.agg_1.0", c1, 1, 0)
    c1 = igev_basic3d(sd, f"{p}.agg_1.2", igev_basic3d(sd, f"{p}.agg_1.1", c1))
    c1 = igev_feature_att(sd, f"{p}.feature_att_up_8", c1, features[1])
    return igev_basic3d(sd, f"{p}.conv1_up", c1, 2, 1, act="none", transposed=True, bn=False)


def igev_cost_volume(sd: SD, match_left, match_right, features_left, max_disp: int):
    """IGEVStereo.forward lines 205-213: gwc volume (8 groups) -> corr_stem -> corr_feature_att -> cost_agg ->
    classifier -> softmax regression at 1/4 resolution (keepdim).  Returns (init_disp, geo_encoding_volume)."""
    vol = R.build_gwc_volume(match_left, match_right, max_disp // 4, 8)
    x = igev_basic3d(sd, "corr_stem", vol)
    x = igev_feature_att(sd, "corr_feature_att", x, features_left[0])
    geo = igev_hourglass(sd, "cost_agg", x, features_left)
    cost = F.conv3d(geo, sd["classifier.weight"], padding=1)
    return R.softargmin(cost[:, 0], keepdim=True), geo


# ------------------------------------------------------------------ CFNet / PCWNet aggregation blocks (hot path)
def mish_hourglass(sd: SD, p: str, x):
    """hourglass.forward: CFNet/cfnet.py:257-271 == PCWNet/pcwnet.py:237-251."""
    c1 = cbn3(sd, f"{p}.conv1.0", x, 2, 1, "mish")
    c2 = cbn3(sd, f"{p}.conv2.0", c1, 1, 1, "mish")
    c3 = cbn3(sd, f"{p}.conv3.0", c2, 2, 1, "mish")
    c4 = cbn3(sd, f"{p}.conv4.0", c3, 1, 1, "mish")
    c5 = dbn3(sd, f"{p}.conv5", c4, "mish", cbn3(sd, f"{p}.redir2", c2, 1, 0))
    return dbn3(sd, f"{p}.conv6", c5, "mish", cbn3(sd, f"{p}.redir1", x, 1, 0))


def mish_hourglassup(sd: SD, p: str, x, feature4, feature5, feature6=None):
    """hourglassup.forward: CFNet/cfnet.py:213-229 (two levels) / PCWNet/pcwnet.py:183-209 (three levels)."""
    c1 = torch.cat((F.conv3d(x, sd[f"{p}.conv1.weight"], stride=2, padding=1), feature4), 1)
    c1 = cbn3(sd, f"{p}.combine1.0", c1, 1, 1, "mish")
    c2 = cbn3(sd, f"{p}.conv2.0", c1, 1, 1, "mish")
    c3 = torch.cat((F.conv3d(c2, sd[f"{p}.conv3.weight"], stride=2, padding=1), feature5), 1)
    c3 = cbn3(sd, f"{p}.combine2.0", c3, 1, 1, "mish")
    c4 = cbn3(sd, f"{p}.conv4.0", c3, 1, 1, "mish")
    if feature6 is not None:
        c5 = torch.cat((F.conv3d(c4, sd[f"{p}.conv5.weight"], stride=2, padding=1), feature6), 1)
        c5 = cbn3(sd, f"{p}.combine3.0", c5, 1, 1, "mish")
        c6 = cbn3(sd, f"{p}.conv6.0", c5, 1, 1, "mish")
        c4 = dbn3(sd, f"{p}.conv7", c6, "mish", cbn3(sd, f"{p}.redir3", c4, 1, 0))
    c8 = dbn3(sd, f"{p}.conv8", c4, "mish", cbn3(sd, f"{p}.redir2", c2, 1, 0))
    return dbn3(sd, f"{p}.conv9", c8, "mish", cbn3(sd, f"{p}.redir1", x, 1, 0))
