"""IGEV-Stereo drop-in (reference: models/IGEVStereo/igev_stereo.py:92-255, extractor.py, update.py, submodule.py).

Same constructor -- ``IGEVStereo(args=None, imagenet_norm=False)`` with ``args`` a **dict** merged key by key
(igev_stereo.py:96-110) -- same ``forward(image1, image2, iters=None, flow_init=None, test_mode=None)``, same eval
return ``[B,1,H,W]`` and the reference's parameter names, so an IGEV-Stereo checkpoint loads with
``load_state_dict(strict=True)``.

Split of the work:
* hot path (libstb200.so, inherited from ``igev.IGEVCostVolume.stage``): group-wise correlation volume (8 groups of
  96 channels), ``corr_stem`` / ``corr_feature_att`` / ``cost_agg`` / ``classifier`` (3-D convs, k4-s2 transposed convs,
  feature gates), soft-argmin at 1/4 resolution, and ``Combined_Geo_Encoding_Volume`` -- all-pairs correlation, the two
  pyramids and the 2-level x 9-tap x (8+1)-channel lookup executed in every GRU iteration (one kernel launch each);
* torch glue ("next"): MobileNetV2 feature network, context network, ConvGRU update
  block, learned convex upsampling.

The reference takes its MobileNetV2 from ``timm_0_5_4`` and downloads ImageNet weights in the constructor
(extractor.py:331).  timm is not a dependency here: ``mobilenetv2.MobileNetV2Trunk`` has timm's module names, is
randomly initialised, and is overwritten by the checkpoint the user loads (IGEV checkpoints contain ``feature.*``).
"""
from __future__ import annotations

import argparse

import torch
import torch.nn as nn
import torch.nn.functional as F

from .igev import BasicConv, IGEVCostVolume
from .mobilenetv2 import MobileNetV2Trunk
from .raft_stereo import ConvGRU, ResidualBlock, _Trunk, _interp, _pool2x, glue_channels_last


# ------------------------------------------------------------------------------------------ 2-D building blocks
class BasicConv_IN(nn.Module):
    """Conv2d / ConvTranspose2d (no bias) + InstanceNorm2d (no parameters) + LeakyReLU(0.01): submodule.py:79-107."""

    def __init__(self, in_channels, out_channels, deconv=False, IN=True, relu=True, **kwargs):
        super().__init__()
        self.relu, self.use_in = relu, IN
        self.conv = (nn.ConvTranspose2d if deconv else nn.Conv2d)(in_channels, out_channels, bias=False, **kwargs)
        self.IN = nn.InstanceNorm2d(out_channels)

    def forward(self, x):
        x = self.conv(x)
        if self.use_in:
            x = self.IN(x)
        return F.leaky_relu(x, 0.01) if self.relu else x


class _Up2x(nn.Module):
    """Shared body of Conv2x / Conv2x_IN (submodule.py:38-76, 110-148), 2-D, concat=True, keep_concat=True: a stride-2
    (de)conv, nearest resize to the skip tensor if the sizes differ, channel concat, 3x3 conv on 2*out channels."""

    def _build(self, block, in_channels, out_channels, deconv, **norm):
        self.conv1 = block(in_channels, out_channels, deconv, kernel_size=4 if deconv else 3, stride=2, padding=1)
        self.conv2 = block(out_channels * 2, out_channels * 2, False, kernel_size=3, stride=1, padding=1, **norm)

    def forward(self, x, rem):
        x = self.conv1(x)
        if x.shape != rem.shape:
            x = F.interpolate(x, size=rem.shape[-2:], mode="nearest")
        return self.conv2(torch.cat((x, rem), 1))


class Conv2x_IN(_Up2x):
    def __init__(self, in_channels, out_channels, deconv=False):
        super().__init__()
        self._build(BasicConv_IN, in_channels, out_channels, deconv)


class Conv2x(_Up2x):
    def __init__(self, in_channels, out_channels, deconv=False):
        super().__init__()
        self._build(lambda i, o, d, **kw: BasicConv(i, o, deconv=d, **kw), in_channels, out_channels, deconv)


def _in_head(cin, cout, stride):
    """BasicConv_IN + Conv2d + InstanceNorm2d + ReLU (stem_2 / stem_4 / spx_4: igev_stereo.py:123-140)."""
    return nn.Sequential(BasicConv_IN(cin, cout, kernel_size=3, stride=stride, padding=1),
                         nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.InstanceNorm2d(cout), nn.ReLU())


class Feature(nn.Module):
    """extractor.py:327-362: MobileNetV2 stages regrouped into block0..4 (1/2 .. 1/32) + a 3-level U-Net decoder.
    Returns [x4 (48 ch), x8 (64), x16 (192), x32 (160)]."""

    def __init__(self):
        super().__init__()
        trunk = MobileNetV2Trunk()
        cut = (0, 1, 2, 3, 5, 6)
        self.conv_stem, self.bn1, self.act1 = trunk.conv_stem, trunk.bn1, trunk.act1
        for i in range(5):
            setattr(self, f"block{i}", nn.Sequential(*trunk.blocks[cut[i]:cut[i + 1]]))
        self.deconv32_16 = Conv2x_IN(160, 96, deconv=True)
        self.deconv16_8 = Conv2x_IN(192, 32, deconv=True)
        self.deconv8_4 = Conv2x_IN(64, 24, deconv=True)
        self.conv4 = BasicConv_IN(48, 48, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        x2 = self.block0(self.act1(self.bn1(self.conv_stem(x))))
        x4 = self.block1(x2)
        x8 = self.block2(x4)
        x16 = self.block3(x8)
        x32 = self.block4(x16)
        x16 = self.deconv32_16(x32, x16)
        x8 = self.deconv16_8(x16, x8)
        x4 = self.conv4(self.deconv8_4(x8, x4))
        return [x4, x8, x16, x32]


class MultiBasicEncoder(_Trunk):
    """Context network, extractor.py:198-296 (RAFT-Stereo's, with heads named after their resolution 1/4, 1/8, 1/16)."""

    def __init__(self, output_dim=[128], norm_fn="batch", dropout=0.0, downsample=3):
        super().__init__(norm_fn, downsample)
        self.layer4 = self._make_layer(128, 2)
        self.layer5 = self._make_layer(128, 2)
        self.outputs04 = nn.ModuleList([nn.Sequential(ResidualBlock(128, 128, norm_fn, 1), nn.Conv2d(128, d[2], 3, padding=1))
                                        for d in output_dim])
        self.outputs08 = nn.ModuleList([nn.Sequential(ResidualBlock(128, 128, norm_fn, 1), nn.Conv2d(128, d[1], 3, padding=1))
                                        for d in output_dim])
        self.outputs16 = nn.ModuleList([nn.Conv2d(128, d[0], 3, padding=1) for d in output_dim])
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else None
        self._init()

    def forward(self, x, num_layers=3):
        x = self.trunk(x)
        outs = [[f(x) for f in self.outputs04]]
        if num_layers >= 2:
            y = self.layer4(x)
            outs.append([f(y) for f in self.outputs08])
        if num_layers >= 3:
            outs.append([f(self.layer5(y)) for f in self.outputs16])
        return tuple(outs)


# ------------------------------------------------------------------------------------------ update block
class DispHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256, output_dim=1):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, output_dim, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class BasicMotionEncoder(nn.Module):
    """update.py:72-92: the 162-channel geometry lookup and the current disparity -> 128 motion channels."""

    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) * (8 + 1)
        self.convc1 = nn.Conv2d(cor_planes, 64, 1)
        self.convc2 = nn.Conv2d(64, 64, 3, padding=1)
        self.convd1 = nn.Conv2d(1, 64, 7, padding=3)
        self.convd2 = nn.Conv2d(64, 64, 3, padding=1)
        self.conv = nn.Conv2d(128, 128 - 1, 3, padding=1)

    def forward(self, disp, corr):
        cor = F.relu(self.convc2(F.relu(self.convc1(corr))))
        dsp = F.relu(self.convd2(F.relu(self.convd1(disp))))
        out = F.relu(self.conv(torch.cat([cor, dsp], dim=1)))
        return torch.cat([out, disp], dim=1)


class BasicMultiUpdateBlock(nn.Module):
    """update.py:115-153: three ConvGRUs (1/16 -> 1/8 -> 1/4), disparity head, 32-channel upsampling feature."""

    def __init__(self, args, hidden_dims=[]):
        super().__init__()
        self.args = args
        self.encoder = BasicMotionEncoder(args)
        self.gru04 = ConvGRU(hidden_dims[2], 128 + hidden_dims[1] * (args.n_gru_layers > 1))
        self.gru08 = ConvGRU(hidden_dims[1], hidden_dims[0] * (args.n_gru_layers == 3) + hidden_dims[2])
        self.gru16 = ConvGRU(hidden_dims[0], hidden_dims[1])
        self.disp_head = DispHead(hidden_dims[2], hidden_dim=256, output_dim=1)
        self.mask_feat_4 = nn.Sequential(nn.Conv2d(hidden_dims[2], 32, 3, padding=1), nn.ReLU(inplace=True))

    def forward(self, net, inp, corr=None, disp=None, iter04=True, iter08=True, iter16=True, update=True):
        if iter16:
            net[2] = self.gru16(net[2], *(inp[2]), _pool2x(net[1]))
        if iter08:
            extra = (_interp(net[2], net[1]),) if self.args.n_gru_layers > 2 else ()
            net[1] = self.gru08(net[1], *(inp[1]), _pool2x(net[0]), *extra)
        if iter04:
            motion = self.encoder(disp, corr)
            extra = (_interp(net[1], net[0]),) if self.args.n_gru_layers > 1 else ()
            net[0] = self.gru04(net[0], *(inp[0]), motion, *extra)
        if not update:
            return net
        return net, self.mask_feat_4(net[0]), self.disp_head(net[0])


def context_upsample(disp_low, up_weights):
    """submodule.py:243-255: each full-resolution pixel is a convex combination (weights [B,9,4h,4w], already
    soft-maxed) of the 3x3 neighbourhood of its 1/4-resolution parent.  Returns [B,4h,4w]."""
    b, c, h, w = disp_low.shape
    taps = F.unfold(disp_low, 3, 1, 1).reshape(b, 9, h, w)
    taps = F.interpolate(taps, (h * 4, w * 4), mode="nearest")
    return (taps * up_weights).sum(1)


# ------------------------------------------------------------------------------------------ the model
class IGEVStereo(IGEVCostVolume):
    def __init__(self, args=None, imagenet_norm=False, precision="fp32"):
        a = argparse.Namespace(hidden_dims=[128] * 3, n_downsample=2, n_gru_layers=3, max_disp=192, valid_iters=32,
                               train_iters=22, precision_dtype="float16", mixed_precision=False, corr_levels=2,
                               corr_radius=4)
        if args is not None:
            for key in args:
                setattr(a, key, args[key])
        # corr_stem / corr_feature_att / cost_agg / classifier + the CUDA backend (igev.IGEVCostVolume)
        super().__init__(max_disp=a.max_disp, corr_levels=a.corr_levels, corr_radius=a.corr_radius, precision=precision)
        self.args = a
        self.imagenet_norm = imagenet_norm
        ctx = a.hidden_dims
        self.cnet = MultiBasicEncoder(output_dim=[a.hidden_dims, ctx], norm_fn="batch", downsample=a.n_downsample)
        self.update_block = BasicMultiUpdateBlock(a, hidden_dims=a.hidden_dims)
        self.context_zqr_convs = nn.ModuleList([nn.Conv2d(ctx[i], a.hidden_dims[i] * 3, 3, padding=1)
                                                for i in range(a.n_gru_layers)])
        self.feature = Feature()
        self.stem_2 = _in_head(3, 32, 2)
        self.stem_4 = _in_head(32, 48, 2)
        self.spx = nn.Sequential(nn.ConvTranspose2d(2 * 32, 9, kernel_size=4, stride=2, padding=1))
        self.spx_2 = Conv2x_IN(24, 32, True)
        self.spx_4 = _in_head(96, 24, 1)
        self.spx_2_gru = Conv2x(32, 32, True)
        self.spx_gru = nn.Sequential(nn.ConvTranspose2d(2 * 32, 9, kernel_size=4, stride=2, padding=1))
        self.conv = BasicConv_IN(96, 96, kernel_size=3, padding=1, stride=1)
        self.desc = nn.Conv2d(96, 96, kernel_size=1, padding=0, stride=1)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def upsample_disp(self, disp, mask_feat_4, stem_2x):
        """igev_stereo.py:157-166: 9 convex weights per full-resolution pixel from the GRU's 32-channel feature and the
        1/2-resolution stem; disparity scaled by 4 with the resolution."""
        logits = self.spx_gru(self.spx_2_gru(mask_feat_4, stem_2x))
        if disp.is_cuda and not (torch.is_grad_enabled() and (disp.requires_grad or logits.requires_grad)):
            from . import ops
            return ops.context_upsample(disp, logits, scale=4.0, softmax=True).unsqueeze(1)   # softmax + x4 + 9 taps in one kernel
        spx_pred = F.softmax(logits, 1)
        return context_upsample(disp * 4.0, spx_pred).unsqueeze(1)

    def _iteration(self, net_list, inp_list, geo_fn, coords, disp):
        """One GRU iteration of igev_stereo.py:237-242: geometry lookup (one CUDA kernel) + update block (torch)."""
        a = self.args
        disp = disp.detach()
        geo_feat = geo_fn(disp, coords)
        net_list, mask_feat_4, delta_disp = self.update_block(net_list, inp_list, geo_feat, disp,
                                                              iter16=a.n_gru_layers == 3, iter08=a.n_gru_layers >= 2)
        return net_list, mask_feat_4, disp + delta_disp

    def _iterate_umma(self, net_list, inp_list, geo_fn, coords, disp, iters):
        """The GRU loop with the update block on the tensor-core 2-D conv path (update_umma.UmmaIgevUpdate, exact 'fp16x2'
        format, hidden states resident in kernel layout).  Opt-in: ``model.update_mode = "umma"`` (inference); with
        ``model.cuda_graph`` one iteration is captured per shape and replayed.  Returns (disp, mask_feat_4 of the last iterate)."""
        from .update_umma import UmmaIgevUpdate
        upd = self.__dict__.get("_umma_update")
        if upd is None:
            upd = self.__dict__["_umma_update"] = UmmaIgevUpdate(self.update_block, self.args)
        upd._prepare()
        net = [upd.to_cl(t) for t in net_list]
        ctx = upd.context(inp_list)

        def one(net, ctx, fn, co, d):
            net, delta = upd.step(net, ctx, fn(d, co), d)
            return net, d + delta

        if not getattr(self, "cuda_graph", False) or iters < 2:
            for _ in range(iters):
                net, disp = one(net, ctx, geo_fn, coords, disp)
            return disp, upd.mask(net[0])
        key = ("umma", tuple(disp.shape), tuple(tuple(t.shape) for t in net), str(disp.device))
        cache = self.__dict__.setdefault("_graph_cache", {})
        hit = cache.get(key)
        if hit is None:
            st = dict(net=[t.clone() for t in net], ctx=[(a.clone(), b.clone()) for a, b in ctx], disp=disp.clone(),
                      coords=coords.clone(), geos=[t.clone() for t in geo_fn._geos], corrs=[t.clone() for t in geo_fn._corrs])
            geo_fn._geos, geo_fn._corrs = st["geos"], st["corrs"]      # the captured lookup reads the static pyramid buffers
            st["geo_fn"] = geo_fn
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                one(st["net"], st["ctx"], geo_fn, st["coords"], st["disp"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                n2, d2 = one(st["net"], st["ctx"], geo_fn, st["coords"], st["disp"])
                for dst, src in zip(st["net"], n2):
                    dst.copy_(src)
                st["disp"].copy_(d2)
            st["graph"] = graph
            cache[key] = hit = st
        else:
            for dst, src in zip(hit["net"], net):
                dst.copy_(src)
            for (da, db), (sa, sb) in zip(hit["ctx"], ctx):
                da.copy_(sa); db.copy_(sb)
            for dst, src in zip(hit["geos"] + hit["corrs"], list(geo_fn._geos) + list(geo_fn._corrs)):
                dst.copy_(src)
            hit["disp"].copy_(disp)
            hit["coords"].copy_(coords)
        for _ in range(iters):
            hit["graph"].replay()
        return hit["disp"].clone(), upd.mask(hit["net"][0])

    def _iterate_graphed(self, net_list, inp_list, geo_fn, coords, disp, iters):
        """Opt-in (``model.cuda_graph = True``, inference): ONE iteration -- lookup kernel + update block + the write-back of
        its outputs into its own inputs -- is captured into a CUDA graph once per input shape and replayed ``iters`` times
        (same scheme as RAFTStereo._iterate_graphed; at small resolutions the loop is launch-bound).  Per call only the
        graph's static inputs (hidden states, context features, the two pyramids, disparity) are refreshed."""
        key = (tuple(disp.shape), tuple(tuple(t.shape) for t in net_list), str(disp.device))
        cache = self.__dict__.setdefault("_graph_cache", {})
        hit = cache.get(key)
        if hit is None:
            st = dict(net=[t.clone() for t in net_list], inp=[[t.clone() for t in lvl] for lvl in inp_list],
                      disp=disp.clone(), coords=coords.clone(),
                      geos=[t.clone() for t in geo_fn._geos], corrs=[t.clone() for t in geo_fn._corrs])
            geo_fn._geos, geo_fn._corrs = st["geos"], st["corrs"]      # the captured lookup reads the static pyramid buffers
            st["geo_fn"] = geo_fn
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                             # warm-up outside capture (cuDNN plans, lazy loading)
                self._iteration(list(st["net"]), st["inp"], geo_fn, st["coords"], st["disp"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                n2, m2, d2 = self._iteration(list(st["net"]), st["inp"], geo_fn, st["coords"], st["disp"])
                for dst, src in zip(st["net"], n2):
                    dst.copy_(src)
                st["disp"].copy_(d2)
            st["graph"], st["mask"] = graph, m2
            cache[key] = hit = st
        else:
            for dst, src in zip(hit["net"], net_list):
                dst.copy_(src)
            for dl, sl in zip(hit["inp"], inp_list):
                for dst, src in zip(dl, sl):
                    dst.copy_(src)
            for dst, src in zip(hit["geos"] + hit["corrs"], list(geo_fn._geos) + list(geo_fn._corrs)):
                dst.copy_(src)
            hit["disp"].copy_(disp)
            hit["coords"].copy_(coords)
        for _ in range(iters):
            hit["graph"].replay()
        return hit["disp"].clone(), hit["mask"]

    def _folded(self, image1):
        """CUDA inference runs on a shadow copy with every eval BatchNorm2d folded into its convolution (glue.py): at
        1152x1920 cuDNN's NCHW BatchNorm-inference passes of the 2-D networks were 21 of 147 ms.  ``model.fold_bn = False``
        turns it off."""
        if not getattr(self, "fold_bn", True) or self.training or not image1.is_cuda or torch.is_grad_enabled():
            return None
        from .glue import inference_shadow
        ex = torch.zeros(1, 3, 64, 128, device=image1.device)
        sh = inference_shadow(self, lambda m: m(ex, ex, iters=1))
        for k in ("update_mode", "cuda_graph", "channels_last"):
            if k in self.__dict__:
                sh.__dict__[k] = self.__dict__[k]
        return sh

    def forward(self, image1, image2, iters=None, flow_init=None, test_mode=None):
        sh = self._folded(image1)
        if sh is not None:
            return sh(image1, image2, iters=iters, flow_init=flow_init, test_mode=test_mode)
        a = self.args
        if iters is None:
            iters = a.train_iters if self.training else a.valid_iters      # igev_stereo.py:172-176
        if test_mode is None:
            test_mode = not self.training                                  # :177-184
        a.mixed_precision = False      # the reference's training autocast is an optimisation, not semantics: fp32 here
        if not self.imagenet_norm:
            mean = torch.tensor([0.485, 0.456, 0.406], device=image1.device).view(1, 3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225], device=image1.device).view(1, 3, 1, 1)
            image1 = 2 * (image1 * std + mean) - 1.0
            image2 = 2 * (image2 * std + mean) - 1.0

        # NHWC torch glue (raft_stereo.glue_channels_last): cuDNN's tensor-core kernels are NHWC, with NCHW tensors torch brackets
        # the 2-D convs with layout transposes (measured at 1152x1920: 152 -> 140 ms).  Default for CUDA inference (where the
        # forward runs on the folded shadow copy, so the model's own parameters keep their layout); ``channels_last = False`` opts out
        if getattr(self, "channels_last", image1.is_cuda and not self.training and not torch.is_grad_enabled()):
            image1, image2 = glue_channels_last(self, image1, image2)

        # ---- torch glue: 2-D features (igev_stereo.py:195-204)
        features_left = self.feature(image1)
        features_right = self.feature(image2)
        stem_2x = self.stem_2(image1)
        stem_4x = self.stem_4(stem_2x)
        stem_4y = self.stem_4(self.stem_2(image2))
        features_left[0] = torch.cat((features_left[0], stem_4x), 1)
        features_right[0] = torch.cat((features_right[0], stem_4y), 1)
        match_left = self.desc(self.conv(features_left[0]))
        match_right = self.desc(self.conv(features_right[0]))

        # ---- hot path: volume + 3-D aggregation + soft-argmin + geometry-encoding pyramids (:205-213, :229-230)
        init_disp, geo_fn, _ = self.stage(match_left, match_right, features_left)

        if not test_mode:                                            # :217-221
            spx_pred = F.softmax(self.spx(self.spx_2(self.spx_4(features_left[0]), stem_2x)), 1)

        cnet_list = self.cnet(image1, num_layers=a.n_gru_layers)
        net_list = [torch.tanh(x[0]) for x in cnet_list]
        inp_list = [torch.relu(x[1]) for x in cnet_list]
        inp_list = [list(conv(i).split(conv.out_channels // 3, dim=1)) for i, conv in zip(inp_list, self.context_zqr_convs)]

        b, _, h, w = match_left.shape
        coords = torch.arange(w, device=match_left.device).float().reshape(1, 1, w, 1).repeat(b, h, 1, 1)
        disp = init_disp
        disp_preds = []
        # update_mode: "auto" (default) = the tensor-core update block whenever it applies (CUDA inference), "torch" = torch/cuDNN
        if (getattr(self, "update_mode", "auto") in ("umma", "auto") and test_mode and not self.training and disp.is_cuda
                and not torch.is_grad_enabled() and iters >= 1 and all(d % 16 == 0 for d in a.hidden_dims)):
            disp, mask_feat_4 = self._iterate_umma(net_list, inp_list, geo_fn, coords, disp.detach(), iters)
            return self.upsample_disp(disp, mask_feat_4, stem_2x)
        if (getattr(self, "cuda_graph", False) and test_mode and not self.training and iters > 1 and disp.is_cuda
                and not torch.is_grad_enabled()):
            disp, mask_feat_4 = self._iterate_graphed(net_list, inp_list, geo_fn, coords, disp, iters)
            return self.upsample_disp(disp, mask_feat_4, stem_2x)
        for itr in range(iters):
            net_list, mask_feat_4, disp = self._iteration(net_list, inp_list, geo_fn, coords, disp)   # hot path: one lookup kernel
            if test_mode and itr < iters - 1:
                continue                                             # only the last iterate is upsampled
            disp_preds.append(self.upsample_disp(disp, mask_feat_4, stem_2x))
        if test_mode:
            return disp_preds[-1]
        return context_upsample(init_disp * 4.0, spx_pred.float()).unsqueeze(1), disp_preds
