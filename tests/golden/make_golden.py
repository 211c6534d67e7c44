#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, read-only).  The reference is
imported with the stub recipe of SURVEY.md section 8c (its ``models/__init__.py`` pulls in timm /
opt_einsum, which are absent).  Inputs are seeded; both inputs and reference outputs are
stored, so the fixtures do not depend on RNG reproducibility.  Model fixtures store only the
input seed, the weight checksum and the outputs (weights come from
``stereo_toolbox_b200.synth.synth_state_dict``, keyed by parameter name).

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import stereo_toolbox  # noqa: E402  (empty __init__)

_m = types.ModuleType("stereo_toolbox.models")
_m.__path__ = ["/root/reference/stereo_toolbox/models"]
sys.modules["stereo_toolbox.models"] = _m
_oe = types.ModuleType("opt_einsum")
_oe.contract = None
sys.modules["opt_einsum"] = _oe
sys.modules.setdefault("timm_0_5_4", types.ModuleType("timm_0_5_4"))

from stereo_toolbox_b200.synth import synth_state_dict, state_checksum, synth_pair  # noqa: E402


def ref(mod):
    return importlib.import_module("stereo_toolbox.models." + mod)


def rnd(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


@torch.no_grad()
def ops():
    gsub = ref("GwcNet.submodule")
    asub = ref("ACVNet.submodule")
    psub = ref("PSMNet.submodule")
    out = {}
    # --- volumes
    L, R = rnd(1, 2, 16, 5, 23), rnd(2, 2, 16, 5, 23)
    out.update(gwc_L=L, gwc_R=R, gwc_out=gsub.build_gwc_volume(L, R, 9, 4))
    L8, R8 = rnd(3, 1, 24, 4, 12), rnd(4, 1, 24, 4, 12)
    out.update(gwc8_L=L8, gwc8_R=R8, gwc8_out=ref("IGEVStereo.submodule").build_gwc_volume(L8, R8, 6, 8))
    cl, cr = rnd(5, 2, 3, 5, 23), rnd(6, 2, 3, 5, 23)
    out.update(cat_L=cl, cat_R=cr, catA_out=gsub.build_concat_volume(cl, cr, 9),
               catB_out=asub.build_concat_volume(cl, cr, 9))
    att = rnd(7, 2, 1, 9, 5, 23)
    out.update(att=att, acv_out=torch.softmax(att, dim=2) * out["catB_out"])
    # --- head
    cost = rnd(8, 2, 1, 6, 5, 7) * 3
    import torch.nn.functional as F
    for ac in (False, True):
        up = F.interpolate(cost, [24, 20, 28], mode="trilinear", align_corners=ac)
        out[f"up_{int(ac)}"] = up[:, 0]
        prob = F.softmax(up[:, 0], dim=1)
        out[f"head_{int(ac)}"] = gsub.disparity_regression(prob, 24)
    out["head_cost"] = cost
    out["head_keepdim"] = psub.disparityregression(24)(F.softmax(out["up_0"], dim=1))
    # IGEV style: no upsampling, keepdim
    c4 = rnd(9, 2, 12, 5, 7)
    out.update(sam_cost=c4, sam_out=ref("IGEVStereo.submodule").disparity_regression(F.softmax(c4, dim=1), 12))
    save("ops_volume_head.npz", **out)

    # --- 1-D correlation (RAFT)
    out = {}
    corr_mod = ref("RAFTStereo.corr")
    f1, f2 = rnd(10, 2, 8, 3, 32), rnd(11, 2, 8, 3, 32)
    blk = corr_mod.CorrBlock1D(f1, f2, num_levels=4, radius=4)
    g = torch.Generator().manual_seed(12)
    coords = torch.arange(32.0).view(1, 1, 1, 32).repeat(2, 2, 3, 1)
    coords[:, 0] += (torch.rand(2, 3, 32, generator=g) - 0.5) * 48  # includes out-of-range taps
    out.update(f1=f1, f2=f2, coords=coords, corr=corr_mod.CorrBlock1D.corr(f1, f2)[:, :, :, 0],
               lookup=blk(coords))
    for i, p in enumerate(blk.corr_pyramid):
        out[f"pyr{i}"] = p.reshape(2, 3, 32, -1)
    # --- IGEV geometry
    geo_mod = ref("IGEVStereo.geometry")
    m1, m2, gv = rnd(13, 1, 8, 3, 16), rnd(14, 1, 8, 3, 16), rnd(15, 1, 4, 8, 3, 16)
    geo = geo_mod.Combined_Geo_Encoding_Volume(m1, m2, gv, num_levels=2, radius=4)
    disp = torch.rand(1, 1, 3, 16, generator=g) * 10 - 1
    cx = torch.arange(16.0).view(1, 1, 1, 16).repeat(1, 1, 3, 1)
    out.update(geo_m1=m1, geo_m2=m2, geo_vol=gv, geo_disp=disp, geo_coords=cx, geo_out=geo(disp, cx))
    save("ops_corr.npz", **out)


def _load_synth(model, seed=0, calib=None, calib_name=None):
    """Name-keyed synthetic weights; with ``calib`` (a left/right pair) the BatchNorm running
    statistics are then replaced by the batch statistics of one reference forward in train mode
    (momentum=None -> exact batch mean / unbiased var), so that every BN really normalises and
    activations stay O(1) like in a trained network.  The calibrated statistics are saved as
    ``bn_calib_<name>.npz`` and re-used by tests and bench.py."""
    sd = synth_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd, strict=True)
    if calib is not None:
        bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        for m in bns:
            m.reset_running_stats()
            m.momentum = None
        model.train()
        model(*calib)
        for m in bns:
            m.momentum = 0.1
            m.num_batches_tracked.zero_()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        save(f"bn_calib_{calib_name}.npz", **{k: v for k, v in sd.items()
                                               if k.endswith("running_mean") or k.endswith("running_var")})
    model.eval()
    return sd


def _keys(sd):
    return {k: list(v.shape) for k, v in sd.items()}


@torch.no_grad()
def blocks():
    """One hourglass of each flavour: conv / strided conv / transposed conv / BN fold / residual order."""
    out, meta = {}, {}
    hg = ref("GwcNet.gwcnet").hourglass(8)
    sd = _load_synth(hg)
    x = rnd(20, 1, 8, 8, 12, 16)
    out.update(gwc_hg_x=x, gwc_hg_y=hg(x))
    meta["gwc_hg"] = dict(keys=_keys(sd), checksum=state_checksum(sd))
    ph = ref("PSMNet.stackhourglass").hourglass(8)
    sd = _load_synth(ph)
    y, pre, post = ph(x, None, None)
    y2, pre2, post2 = ph(x, pre, post)
    out.update(psm_hg_y=y, psm_hg_pre=pre, psm_hg_post=post, psm_hg_y2=y2, psm_hg_pre2=pre2, psm_hg_post2=post2)
    meta["psm_hg"] = dict(keys=_keys(sd), checksum=state_checksum(sd))
    save("blocks.npz", **out)
    json.dump(meta, open(os.path.join(HERE, "blocks.json"), "w"))


@torch.no_grad()
def models():
    meta = {}
    # GwcNet_GC, d=32, 64x128 pair  (volume [1,64,8,16,32])
    calib = synth_pair(2, 64, 128, seed=100, shift=3)
    net = ref("GwcNet.gwcnet").GwcNet_GC(32)
    sd = _load_synth(net, calib=calib, calib_name="gwcnet_gc")
    left, right = synth_pair(1, 64, 128, seed=0, shift=5)
    cap = {}
    net.classif3.register_forward_hook(lambda m, i, o: cap.__setitem__("cost3", o))
    net.dres0.register_forward_hook(lambda m, i, o: cap.__setitem__("dres0", o))
    disp = net(left, right)
    save("gwcnet_gc.npz", disp=disp, cost3=cap["cost3"], dres0=cap["dres0"])
    meta["gwcnet_gc"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=32, shape=[1, 64, 128], shift=5)
    net = ref("GwcNet.gwcnet").GwcNet_G(32)
    sd = _load_synth(net, calib=calib, calib_name="gwcnet_g")
    save("gwcnet_g.npz", disp=net(left, right))
    meta["gwcnet_g"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=32, shape=[1, 64, 128], shift=5)
    # PSMNet, maxdisp=32, 256x256 (SPP needs H/4 >= 64)
    net = ref("PSMNet.stackhourglass").PSMNet(32)
    sd = _load_synth(net, calib=synth_pair(2, 256, 256, seed=101, shift=4), calib_name="psmnet")
    left, right = synth_pair(1, 256, 256, seed=1, shift=7)
    cap = {}
    net.classif3.register_forward_hook(lambda m, i, o: cap.__setitem__("c3", o))
    disp = net(left, right)
    save("psmnet.npz", disp=disp, classif3=cap["c3"][:, :, :, ::4, ::4])
    meta["psmnet"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=32, shape=[1, 256, 256], shift=7)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


@torch.no_grad()
def raft():
    """RAFT-Stereo (config 4 family): 1x3x64x128, 6 iterations, default args (3 GRU levels, corr 4 levels r=4)."""
    net = ref("RAFTStereo.raft_stereo").RAFTStereo()
    sd = _load_synth(net)
    left, right = synth_pair(1, 64, 128, seed=2, shift=3)
    out = net(left, right, iters=6)
    save("raft_stereo.npz", disp=out)
    meta = json.load(open(os.path.join(HERE, "models.json")))
    meta["raft_stereo"] = dict(keys=_keys(sd), checksum=state_checksum(sd), shape=[1, 64, 128], shift=3, iters=6)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


@torch.no_grad()
def acv():
    """ACVNet (config 5): op-level fixtures for the block attention (three padding cases, including the reference's
    '-0:' mask quirk) and the depthwise patch convs, plus the whole model at maxdisp 64 on a 64x144 pair
    (1/16-scale volume 4x4x9 -> W padded to 12 inside the attention)."""
    sub = ref("ACVNet.submodule")
    out = {}
    for tag, (D, H, W) in (("nopad", (4, 4, 8)), ("padr", (4, 4, 6)), ("padrb", (8, 6, 7))):
        ab = sub.attention_block(channels_3d=32, num_heads=4, block=(4, 4, 4))
        sd = _load_synth(ab, seed=7)
        x = rnd(30, 2, 32, D, H, W)
        out.update({f"att_{tag}_x": x, f"att_{tag}_y": ab(x)})
        for k, v in sd.items():
            out[f"att_{tag}_{k}"] = v
    vol = rnd(31, 1, 6, 3, 9, 11)
    for dil in (1, 2, 3):
        conv = torch.nn.Conv3d(6, 6, kernel_size=(1, 3, 3), stride=1, dilation=dil, groups=6, padding=(0, dil, dil), bias=False)
        conv.weight.copy_(rnd(32 + dil, 6, 1, 1, 3, 3))
        out.update({f"patch_w{dil}": conv.weight, f"patch_y{dil}": conv(vol)})
    out["patch_x"] = vol
    save("ops_acv.npz", **out)
    net = ref("ACVNet.acv").ACVNet(64)
    sd = _load_synth(net, calib=synth_pair(2, 64, 144, seed=102, shift=3), calib_name="acvnet")
    left, right = synth_pair(1, 64, 144, seed=3, shift=5)
    cap = {}
    net.classif2.register_forward_hook(lambda m, i, o: cap.__setitem__("cost2", o))
    net.classif_att_.register_forward_hook(lambda m, i, o: cap.__setitem__("att", o))
    disp = net(left, right)
    save("acvnet.npz", disp=disp, cost2=cap["cost2"], att=cap["att"])
    meta = json.load(open(os.path.join(HERE, "models.json")))
    meta["acvnet"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=64, shape=[1, 64, 144], shift=5)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


@torch.no_grad()
def igev():
    """IGEV cost-volume stage (config 5 component level; IGEVStereo() itself needs timm weights): the reference's
    corr_stem / corr_feature_att / cost_agg (hourglass(8)) / classifier run exactly as igev_stereo.py:205-213, plus one
    geometry-encoding lookup at the initial disparity (geometry.py)."""
    ig, sub, geo_mod = ref("IGEVStereo.igev_stereo"), ref("IGEVStereo.submodule"), ref("IGEVStereo.geometry")

    class Stage(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.corr_stem = sub.BasicConv(8, 8, is_3d=True, kernel_size=3, stride=1, padding=1)
            self.corr_feature_att = sub.FeatureAtt(8, 96)
            self.cost_agg = ig.hourglass(8)
            self.classifier = torch.nn.Conv3d(8, 1, 3, 1, 1, bias=False)

        def forward(self, ml, mr, feats, max_disp):
            vol = sub.build_gwc_volume(ml, mr, max_disp // 4, 8)
            vol = self.corr_feature_att(self.corr_stem(vol), feats[0])
            geo = self.cost_agg(vol, feats)
            prob = torch.softmax(self.classifier(geo).squeeze(1), dim=1)
            return sub.disparity_regression(prob, max_disp // 4), geo

    net = Stage()
    sd = _load_synth(net, seed=11)
    h, w, max_disp = 16, 32, 64
    ml, mr = rnd(40, 1, 96, h, w) * 0.3, rnd(41, 1, 96, h, w) * 0.3
    feats = [rnd(42, 1, 96, h, w), rnd(43, 1, 64, h // 2, w // 2), rnd(44, 1, 192, h // 4, w // 4), rnd(45, 1, 160, h // 8, w // 8)]
    init_disp, geo = net(ml, mr, feats, max_disp)
    fn = geo_mod.Combined_Geo_Encoding_Volume(ml.float(), mr.float(), geo.float(), radius=4, num_levels=2)
    coords = torch.arange(w).float().reshape(1, 1, w, 1).repeat(1, h, 1, 1)
    geo_feat = fn(init_disp, coords)
    save("igev_stage.npz", ml=ml, mr=mr, f0=feats[0], f1=feats[1], f2=feats[2], f3=feats[3], init_disp=init_disp, geo=geo,
         geo_feat=geo_feat)
    meta = json.load(open(os.path.join(HERE, "blocks.json")))
    meta["igev_stage"] = dict(keys=_keys(sd), checksum=state_checksum(sd), max_disp=max_disp, seed=11)
    json.dump(meta, open(os.path.join(HERE, "blocks.json"), "w"))


@torch.no_grad()
def cascade():
    """CFNet / PCWNet aggregation blocks: Mish hourglass, hourglassup (2 and 3 levels), align_corners=True head,
    disparity_variance."""
    import torch.nn.functional as F
    cf, pc, cfsub = ref("CFNet.cfnet"), ref("PCWNet.pcwnet"), ref("CFNet.submodule")
    out = {}
    meta = json.load(open(os.path.join(HERE, "blocks.json")))
    x = rnd(50, 1, 8, 8, 16, 16) * 0.5
    f4, f5, f6 = rnd(51, 1, 16, 4, 8, 8) * 0.5, rnd(52, 1, 16, 2, 4, 4) * 0.5, rnd(53, 1, 16, 1, 2, 2) * 0.5
    out.update(x=x, f4=f4, f5=f5, f6=f6)
    hg = cf.hourglass(8)
    sd = _load_synth(hg, seed=21)
    out["cf_hg_y"] = hg(x)
    meta["cf_hg"] = dict(keys=_keys(sd), checksum=state_checksum(sd), seed=21)
    up2 = cf.hourglassup(8)
    sd = _load_synth(up2, seed=22)
    out["cf_up_y"] = up2(x, f4, f5)
    meta["cf_up"] = dict(keys=_keys(sd), checksum=state_checksum(sd), seed=22)
    up3 = pc.hourglassup(8)
    sd = _load_synth(up3, seed=23)
    out["pc_up_y"] = up3(x, f4, f5, f6)
    meta["pc_up"] = dict(keys=_keys(sd), checksum=state_checksum(sd), seed=23)
    # head with align_corners=True (cfnet.py:605-613) + variance (CFNet/submodule.py:127-133)
    cost = rnd(54, 2, 1, 6, 5, 7) * 3
    up = F.interpolate(cost, [24, 20, 28], mode="trilinear", align_corners=True)[:, 0]
    prob = F.softmax(up, dim=1)
    disp = cfsub.disparity_regression(prob, 24)
    out.update(head_cost=cost, head_disp=disp, head_prob=prob, head_var=cfsub.disparity_variance(prob, 24, disp.unsqueeze(1)))
    save("blocks_cascade.npz", **out)
    json.dump(meta, open(os.path.join(HERE, "blocks.json"), "w"))


def train():
    """One PSMNet training step of the reference on CPU (config 3 family, small): train-mode forward (batch-statistic
    BatchNorm), the trainer's loss (smooth-L1 on the three outputs, weights 0.5/0.7/1.0, mask 0 < gt < maxdisp:
    trainer/trainer_torchrun.py:272-278), backward.  Stores the three predictions, the loss and the gradients of a few
    parameters that cover conv / strided conv / transposed conv / classifier / BatchNorm3d / the 2-D extractor (i.e. the
    volume adjoint)."""
    import torch.nn.functional as F
    net = ref("PSMNet.stackhourglass").PSMNet(32)
    z = np.load(os.path.join(HERE, "bn_calib_psmnet.npz"))
    sd = synth_state_dict(net.state_dict(), 0, {k: z[k] for k in z.files})
    net.load_state_dict(sd, strict=True)
    net.train()
    left, right = synth_pair(2, 256, 256, seed=1, shift=7)     # batch 2: the 64x64 SPP branch leaves one value per channel and image
    from stereo_toolbox_b200.synth import synth_gt
    gt = synth_gt(2, 256, 256)
    preds = net(left, right)
    mask = (gt > 0) & (gt < 32)
    loss = sum(w * F.smooth_l1_loss(p.squeeze(1)[mask], gt[mask], reduction="mean") for w, p in zip((0.5, 0.7, 1.0), preds))
    loss.backward()
    names = ["dres0.0.0.weight", "dres0.0.1.weight", "dres0.0.1.bias", "dres1.2.0.weight", "dres2.conv1.0.0.weight",
             "dres2.conv5.0.weight", "dres3.conv6.0.weight", "dres4.conv2.0.weight", "classif1.2.weight", "classif3.0.0.weight",
             "classif3.2.weight", "feature_extraction.lastconv.2.weight", "feature_extraction.firstconv.0.0.weight"]
    params = dict(net.named_parameters())
    sub = lambda p: p.detach()[:, :, ::2, ::2]
    out = {"loss": loss.detach(), "pred1": sub(preds[0]), "pred2": sub(preds[1]), "pred3": sub(preds[2]),
           "rm_dres0": net.dres0[0][1].running_mean.detach().clone(), "rv_dres0": net.dres0[0][1].running_var.detach().clone()}
    for n in names:
        out["grad:" + n] = params[n].grad.detach()
    save("psmnet_train.npz", **out)


@torch.no_grad()
def cfnet():
    """CFNet whole model (eval) on CPU.  The reference's UniformSampler / SpatialTransformer pass ``tensor.get_device()``
    (-1 on CPU) as a device index; for the duration of this run ``Tensor.get_device`` is patched to return the tensor's
    ``torch.device`` instead -- the reference sources are not touched.  maxdisp 64, 64x128 pair."""
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device
    try:
        net = ref("CFNet.cfnet").CFNet(64)
        sd = _load_synth(net, calib=synth_pair(2, 64, 128, seed=103, shift=3), calib_name="cfnet")
        left, right = synth_pair(1, 64, 128, seed=6, shift=5)
        disp = net(left, right)
    finally:
        torch.Tensor.get_device = orig
    save("cfnet.npz", disp=disp)
    meta = json.load(open(os.path.join(HERE, "models.json")))
    meta["cfnet"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=64, shape=[1, 64, 128], shift=5)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


@torch.no_grad()
def pcwnet():
    """PCWNet_GC whole model (eval) on CPU; same ``Tensor.get_device`` patch as cfnet() (PCWNet/submodule.py:130)."""
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device
    try:
        net = ref("PCWNet.pcwnet").PCWNet_GC(64)
        sd = _load_synth(net, calib=synth_pair(2, 64, 128, seed=104, shift=3), calib_name="pcwnet_gc")
        left, right = synth_pair(1, 64, 128, seed=7, shift=5)
        cap = {}
        net.classif3.register_forward_hook(lambda m, i, o: cap.__setitem__("cost3", o))
        disp = net(left, right)
    finally:
        torch.Tensor.get_device = orig
    save("pcwnet_gc.npz", disp=disp, cost3=cap["cost3"])
    meta = json.load(open(os.path.join(HERE, "models.json")))
    meta["pcwnet_gc"] = dict(keys=_keys(sd), checksum=state_checksum(sd), maxdisp=64, shape=[1, 64, 128], shift=5)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


@torch.no_grad()
def igev_model():
    """IGEVStereo whole model (config 5 family).  The reference's constructor asks ``timm_0_5_4`` for a pretrained
    MobileNetV2 (extractor.py:331); timm is absent and there is no network, so for this run the stub module's
    ``create_model`` hands it ``stereo_toolbox_b200.mobilenetv2.MobileNetV2Trunk`` (same stage table and module names
    as timm's mobilenetv2_100).  Everything else -- Feature decoder, stems, context network, cost-volume stage, geometry
    lookup, GRU loop, convex upsampling -- is the reference's own code.  max_disp 64, 64x128 pair, 4 iterations."""
    from stereo_toolbox_b200.mobilenetv2 import MobileNetV2Trunk
    sys.modules["timm_0_5_4"].create_model = lambda *a, **k: MobileNetV2Trunk()
    net = ref("IGEVStereo.igev_stereo").IGEVStereo({"max_disp": 64})
    sd = _load_synth(net, calib=synth_pair(2, 64, 128, seed=105, shift=3), calib_name="igev_stereo")
    # ResidualBlock registers its shortcut norm twice (norm3 and downsample.1 are ONE module, extractor.py:27,49): the
    # name-keyed generator draws two values, load_state_dict keeps the later name's.  The fingerprint is therefore taken
    # over the name-keyed dict the tests rebuild, not over the de-duplicated state_dict() of the loaded model.
    sd = synth_state_dict(net.state_dict(), 0, {k: v for k, v in sd.items() if k.endswith(("running_mean", "running_var"))})
    left, right = synth_pair(1, 64, 128, seed=8, shift=5)
    cap = {}
    net.classifier.register_forward_hook(lambda m, i, o: cap.__setitem__("cost", o))
    def first_delta(m, i, o):          # a hook that returns a value would replace the module's output
        cap.setdefault("delta0", o[2])
    net.update_block.register_forward_hook(first_delta)
    disp = net(left, right, iters=4)
    init_disp, preds = net(left, right, iters=2, test_mode=False)
    save("igev_stereo.npz", disp=disp, cost=cap["cost"], delta0=cap["delta0"], init_disp_up=init_disp, pred_last=preds[-1])
    meta = json.load(open(os.path.join(HERE, "models.json")))
    meta["igev_stereo"] = dict(keys=_keys(sd), checksum=state_checksum(sd), max_disp=64, shape=[1, 64, 128], shift=5, iters=4)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"))


VARIANTS = {   # constructor-argument variants of the iterative models (host-side wiring; hot-path ops unchanged)
    "raft_gru2_slowfast": ("raft", dict(n_gru_layers=2, slow_fast_gru=True)),
    "raft_gru1": ("raft", dict(n_gru_layers=1)),
    "raft_down3_lvl2_r3": ("raft", dict(n_downsample=3, corr_levels=2, corr_radius=3)),
    "raft_instance_ctx": ("raft", dict(context_norm="instance")),
    "igev_gru2_r3": ("igev", dict(max_disp=64, n_gru_layers=2, corr_radius=3)),
    "igev_gru1_d96": ("igev", dict(max_disp=96, n_gru_layers=1)),
}


@torch.no_grad()
def variants():
    """RAFTStereo (args Namespace) / IGEVStereo (args dict) with non-default arguments, 64x128 pair, 3 iterations,
    un-calibrated name-keyed weights.  Stored: the output and the parameter-name -> shape table of every variant."""
    import argparse
    from stereo_toolbox_b200.mobilenetv2 import MobileNetV2Trunk
    sys.modules["timm_0_5_4"].create_model = lambda *a, **k: MobileNetV2Trunk()
    left, right = synth_pair(1, 64, 128, seed=2, shift=3)
    out, meta = {}, {}
    for tag, (family, kw) in VARIANTS.items():
        if family == "raft":
            net = ref("RAFTStereo.raft_stereo").RAFTStereo(argparse.Namespace(**kw))
        else:
            net = ref("IGEVStereo.igev_stereo").IGEVStereo(dict(kw))
        sd = _load_synth(net)
        out[tag] = net(left, right, iters=3)
        meta[tag] = dict(family=family, args=kw, keys=_keys(sd), checksum=state_checksum(sd))
    save("variants.npz", **out)
    json.dump(meta, open(os.path.join(HERE, "variants.json"), "w"))


def raft_train():
    """One RAFT-Stereo training step of the reference on CPU (train mode, BatchNorm frozen as its constructor does,
    3 iterations, default args): the per-iteration predictions, a sequence loss (L1 to a synthetic ground truth,
    gamma = 0.9) and the gradients of parameters on both sides of CorrBlock1D -- the feature network only receives
    gradient through the all-pairs correlation + pyramid + lookup."""
    from stereo_toolbox_b200.synth import synth_gt
    net = ref("RAFTStereo.raft_stereo").RAFTStereo()
    _load_synth(net)
    net.train()
    net.freeze_bn()
    left, right = synth_pair(1, 64, 128, seed=2, shift=3)
    gt = synth_gt(1, 64, 128)[:, None] * 0.25
    preds = net(left, right, iters=3)
    loss = sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean() for i, p in enumerate(preds))
    loss.backward()
    names = ["fnet.conv1.weight", "fnet.layer2.0.conv1.weight", "fnet.conv2.weight", "cnet.conv1.weight",
             "update_block.encoder.convc1.weight", "update_block.gru08.convz.weight", "update_block.flow_head.conv2.weight",
             "update_block.mask.2.weight", "context_zqr_convs.0.weight"]
    params = dict(net.named_parameters())
    out = {"loss": loss.detach()}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach()[:, :, ::2, ::2]
    for n in names:          # large tensors: every k-th element, k = numel // 20000 (tests apply the same rule)
        flat = params[n].grad.detach().flatten()
        out["grad:" + n] = flat[::max(1, flat.numel() // 20000)]
    save("raft_train.npz", **out)


def igev_train():
    """One IGEV-Stereo training step of the reference on CPU (train mode: batch-statistic BatchNorm, test_mode False,
    2 iterations, max_disp 64; timm stub as in igev_model): upsampled initial disparity, per-iteration predictions, an L1
    loss, and gradients of parameters that only receive gradient through hot-path pieces (volume, 3-D stage incl. the k4
    transposed convs and feature gates, soft-argmin, all-pairs correlation + geometry lookup)."""
    from stereo_toolbox_b200.mobilenetv2 import MobileNetV2Trunk
    from stereo_toolbox_b200.synth import synth_gt
    sys.modules["timm_0_5_4"].create_model = lambda *a, **k: MobileNetV2Trunk()
    net = ref("IGEVStereo.igev_stereo").IGEVStereo({"max_disp": 64})
    z = np.load(os.path.join(HERE, "bn_calib_igev_stereo.npz"))
    net.load_state_dict(synth_state_dict(net.state_dict(), 0, {k: z[k] for k in z.files}), strict=True)
    net.train()
    left, right = synth_pair(2, 64, 128, seed=8, shift=5)
    gt = synth_gt(2, 64, 128)[:, None] * 0.25
    init_disp, preds = net(left, right, iters=2)
    loss = (init_disp - gt).abs().mean() + sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean() for i, p in enumerate(preds))
    loss.backward()
    names = ["corr_stem.conv.weight", "cost_agg.conv1.0.conv.weight", "cost_agg.conv3_up.conv.weight",
             "cost_agg.conv1_up.conv.weight", "cost_agg.feature_att_8.feat_att.1.weight", "classifier.weight", "desc.weight",
             "conv.conv.weight", "feature.conv_stem.weight", "update_block.encoder.convc1.weight", "spx.0.weight"]
    params = dict(net.named_parameters())
    out = {"loss": loss.detach(), "init_disp": init_disp.detach()[:, :, ::2, ::2]}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach()[:, :, ::2, ::2]
    for n in names:
        flat = params[n].grad.detach().flatten()
        out["grad:" + n] = flat[::max(1, flat.numel() // 20000)]
    save("igev_train.npz", **out)


def _grad_sample(params, prefixes):
    """First conv-like weight under each prefix; large tensors sampled every k-th element (k = numel // 20000)."""
    out = {}
    for pre in prefixes:
        name = next(n for n, p in params.items() if n.startswith(pre) and n.endswith("weight") and p.dim() >= 4 and p.grad is not None)
        flat = params[name].grad.detach().flatten()
        out["grad:" + name] = flat[::max(1, flat.numel() // 20000)]
    return out


def pcwnet_train():
    """One PCWNet_GC training step of the reference on CPU (train mode, batch 2, maxdisp 64; the ``Tensor.get_device``
    patch of pcwnet()): the six predictions (pcwnet.py:480), a smooth-L1 loss over them, gradients of one weight per
    sub-network."""
    import torch.nn.functional as F
    from stereo_toolbox_b200.synth import synth_gt
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device
    try:
        net = ref("PCWNet.pcwnet").PCWNet_GC(64)
        z = np.load(os.path.join(HERE, "bn_calib_pcwnet_gc.npz"))
        net.load_state_dict(synth_state_dict(net.state_dict(), 0, {k: z[k] for k in z.files}), strict=True)
        net.train()
        left, right = synth_pair(2, 64, 128, seed=7, shift=5)
        gt = synth_gt(2, 64, 128)
        preds = net(left, right)
        mask = (gt > 0) & (gt < 64)
        loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
        loss.backward()
    finally:
        torch.Tensor.get_device = orig
    out = {"loss": loss.detach()}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach()[:, ::2, ::2]
    out.update(_grad_sample(dict(net.named_parameters()),
                            ["dres0.0.0.", "dres1.", "combine1.conv1.", "combine1.conv9.", "dres3.conv5.", "classif0.",
                             "classif4.", "refinenet3.", "dispupsample.", "feature_extraction.firstconv.",
                             "feature_extraction.layer4."]))
    save("pcwnet_train.npz", **out)


def cfnet_train():
    """One CFNet training step of the reference on CPU (train mode, batch 2, maxdisp 64; ``Tensor.get_device`` patch as in
    cfnet()): the nine predictions (cfnet.py:651), a smooth-L1 loss over them, gradients of one weight per sub-network."""
    import torch.nn.functional as F
    from stereo_toolbox_b200.synth import synth_gt
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device
    try:
        net = ref("CFNet.cfnet").CFNet(64)
        z = np.load(os.path.join(HERE, "bn_calib_cfnet.npz"))
        net.load_state_dict(synth_state_dict(net.state_dict(), 0, {k: z[k] for k in z.files}), strict=True)
        net.train()
        left, right = synth_pair(2, 64, 128, seed=6, shift=5)
        gt = synth_gt(2, 64, 128)
        preds = net(left, right)
        mask = (gt > 0) & (gt < 64)
        loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
        loss.backward()
    finally:
        torch.Tensor.get_device = orig
    out = {"loss": loss.detach()}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach()[:, ::2, ::2]
    out.update(_grad_sample(dict(net.named_parameters()),
                            ["dres0.0.0.", "dres0_6.", "combine1.conv1.", "dres3.conv5.", "classif0.", "classif2.",
                             "confidence0_s3.", "confidence2_s3.conv6.", "confidence_classifmid_s3.", "confidence0_s2.",
                             "confidence3_s2.", "confidence_classif0_s2.", "feature_extraction.firstconv."]))
    save("cfnet_train.npz", **out)


def acvnet_train():
    """One ACVNet training step of the reference on CPU (train mode, batch 2, maxdisp 64, 64x144 pair: the 1/16-scale
    attention runs in the one-sided padding case): the four predictions (acv.py:235), a smooth-L1 loss, gradients of one
    weight per sub-network incl. the patch convs and the attention block."""
    import torch.nn.functional as F
    from stereo_toolbox_b200.synth import synth_gt
    net = ref("ACVNet.acv").ACVNet(64)
    z = np.load(os.path.join(HERE, "bn_calib_acvnet.npz"))
    net.load_state_dict(synth_state_dict(net.state_dict(), 0, {k: z[k] for k in z.files}), strict=True)
    net.train()
    left, right = synth_pair(2, 64, 144, seed=3, shift=5)
    gt = synth_gt(2, 64, 144)
    preds = net(left, right)
    mask = (gt > 0) & (gt < 64)
    loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
    loss.backward()
    out = {"loss": loss.detach()}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach()[:, ::2, ::2]
    params = dict(net.named_parameters())
    out.update(_grad_sample(params, ["patch.", "patch_l2.", "dres1_att_.0.0.", "dres2_att_.conv2.", "dres2_att_.attention_block.final1x1.",
                                     "classif_att_.2.", "concatconv.0.0.", "dres0.0.0.", "dres2.conv5.", "dres3.attention_block.final1x1.",
                                     "classif2.2.", "feature_extraction.firstconv."]))
    for n in ("dres2_att_.attention_block.qkv_3d.weight", "dres3.attention_block.qkv_3d.bias"):
        out["grad:" + n] = params[n].grad.detach().flatten()
    save("acvnet_train.npz", **out)


@torch.no_grad()
def upsampling():
    """the two learned convex upsamplers, op level (RAFTStereo.upsample_flow at factor 4 and 8,
    IGEVStereo context_upsample)."""
    import argparse
    out = {}
    for f in (4, 8):
        net = ref("RAFTStereo.raft_stereo").RAFTStereo(argparse.Namespace(n_downsample={4: 2, 8: 3}[f]))
        flow, mask = rnd(60 + f, 2, 2, 5, 7), rnd(61 + f, 2, 9 * f * f, 5, 7)
        out.update({f"raft_flow{f}": flow, f"raft_mask{f}": mask, f"raft_up{f}": net.upsample_flow(flow, mask)})
    disp, wts = rnd(70, 2, 1, 5, 7), torch.softmax(rnd(71, 2, 9, 20, 28), dim=1)
    out.update(igev_disp=disp, igev_w=wts, igev_up=ref("IGEVStereo.submodule").context_upsample(disp, wts))
    save("ops_upsampling.npz", **out)


@torch.no_grad()
def sampled():
    """CFNet's sampled cascade volume, op level -- the reference's cost_volume_generator ('gwc' and
    'concat') + sample channel exactly as cfnet.py:545-550 concatenates them (``Tensor.get_device`` patch as in cfnet()).
    Samples include values that push w - sample outside the row on both sides."""
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device
    try:
        net = ref("CFNet.cfnet").CFNet(64)
        gl, gr = rnd(80, 2, 16, 5, 23), rnd(81, 2, 16, 5, 23)
        cl, cr = rnd(82, 2, 3, 5, 23), rnd(83, 2, 3, 5, 23)
        g = torch.Generator().manual_seed(84)
        samples = torch.randint(-3, 26, (2, 6, 5, 23), generator=g)
        cat_v, _ = net.cost_volume_generator(cl, cr, samples, "concat")
        gwc_v, smp = net.cost_volume_generator(gl, gr, samples, "gwc", 4)
        vol = torch.cat((gwc_v, cat_v, smp), dim=1)
    finally:
        torch.Tensor.get_device = orig
    save("ops_sampled.npz", gl=gl, gr=gr, cl=cl, cr=cr, samples=samples.float(), vol=vol)


if __name__ == "__main__":
    torch.set_num_threads(8)
    # order matters: models() starts models.json afresh, the later generators add their entries to it
    which = sys.argv[1:] or ["ops", "blocks", "models", "raft", "acv", "igev", "cascade", "train", "cfnet", "pcwnet", "igev_model",
                             "variants", "upsampling", "sampled", "raft_train", "igev_train", "pcwnet_train", "cfnet_train", "acvnet_train"]
    for w in which:
        globals()[w]()
