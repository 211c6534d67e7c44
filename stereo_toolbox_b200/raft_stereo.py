"""RAFT-Stereo drop-in (reference: models/RAFTStereo/raft_stereo.py:26-188, extractor.py, update.py).

Same constructor (``RAFTStereo(args=None, imagenet_norm=False)``, ``args`` a Namespace merged via ``vars()``),
same ``forward(image1, image2, iters=None, flow_init=None, test_mode=False)`` and state-dict names.  The hot
path of the iterative model -- all-pairs 1-D correlation, its 1x2 average pyramid and the 4-level x 9-tap
lookup executed every GRU iteration (``CorrBlock1D``, RAFTStereo/corr.py:110-156) -- runs in libstb200.so, and so do, for
CUDA inference, the update block (every conv on the tcgen05 2-D conv path, ``update_umma.py``; ``update_mode``) and the
learned convex upsampling; the encoders stay in torch (rank 2).
"""
from __future__ import annotations

import argparse

import torch
import torch.nn as nn
import torch.nn.functional as F

from .functional import CorrBlock1D


def glue_channels_last(model: nn.Module, *images):
    """Opt-in (``model.channels_last = True``): run the torch glue (2-D encoders, ConvGRUs, heads) on channels-last
    tensors.  cuDNN's tensor-core convolution kernels are NHWC; fed NCHW tensors torch brackets every conv with
    nchw<->nhwc transposes (measured on the GwcNet extractor: ~46 % of its GPU time, profiles/ncu_launches_r01.txt).
    Same arithmetic, so results agree to summation order.  Only the 4-D weights of 2-D convs are re-laid (``Module.to(
    memory_format=channels_last)`` would also touch the 5-D Conv3d weights of the hot path and raise); the hot-path ops
    make their inputs contiguous themselves (ops._f32c).  Returns the images in channels-last layout."""
    if not model.__dict__.get("_glue_cl_done"):
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
        model.__dict__["_glue_cl_done"] = True
    return tuple(t.contiguous(memory_format=torch.channels_last) for t in images)


def _norm(kind: str, ch: int):
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(ch // 8, ch)
    return nn.Sequential()


class ResidualBlock(nn.Module):
    """extractor.py:5-58: two 3x3 convs with norm + ReLU, 1x1-conv shortcut when the shape changes."""

    def __init__(self, in_planes, planes, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride, 1)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _norm(norm_fn, planes), _norm(norm_fn, planes)
        self.downsample = None
        if stride != 1 or in_planes != planes:
            self.norm3 = _norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        return self.relu((x if self.downsample is None else self.downsample(x)) + y)


class _Trunk(nn.Module):
    """conv1 (7x7) + layer1..3 shared by BasicEncoder and MultiBasicEncoder (extractor.py:122-160, 198-230)."""

    def __init__(self, norm_fn, downsample):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = nn.GroupNorm(8, 64) if norm_fn == "group" else _norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(3, 64, 7, 1 + (downsample > 2), 3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = 64
        self.layer1 = self._make_layer(64, 1)
        self.layer2 = self._make_layer(96, 1 + (downsample > 1))
        self.layer3 = self._make_layer(128, 1 + (downsample > 0))

    def _make_layer(self, dim, stride):
        seq = nn.Sequential(ResidualBlock(self.in_planes, dim, self.norm_fn, stride),
                            ResidualBlock(dim, dim, self.norm_fn, 1))
        self.in_planes = dim
        return seq

    def _init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def trunk(self, x):
        x = self.relu1(self.norm1(self.conv1(x)))
        return self.layer3(self.layer2(self.layer1(x)))


class BasicEncoder(_Trunk):
    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0, downsample=3):
        super().__init__(norm_fn, downsample)
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else None
        self._init()

    def forward(self, x, dual_inp=False):
        is_list = isinstance(x, (tuple, list))
        if is_list:
            n = x[0].shape[0]
            x = torch.cat(x, dim=0)
        x = self.conv2(self.trunk(x))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return x.split(n, dim=0) if is_list else x


class MultiBasicEncoder(_Trunk):
    def __init__(self, output_dim=[128], norm_fn="batch", dropout=0.0, downsample=3):
        super().__init__(norm_fn, downsample)
        self.layer4 = self._make_layer(128, 2)
        self.layer5 = self._make_layer(128, 2)
        self.outputs08 = nn.ModuleList([nn.Sequential(ResidualBlock(128, 128, norm_fn, 1), nn.Conv2d(128, d[2], 3, padding=1))
                                        for d in output_dim])
        self.outputs16 = nn.ModuleList([nn.Sequential(ResidualBlock(128, 128, norm_fn, 1), nn.Conv2d(128, d[1], 3, padding=1))
                                        for d in output_dim])
        self.outputs32 = nn.ModuleList([nn.Conv2d(128, d[0], 3, padding=1) for d in output_dim])
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else None
        self._init()

    def forward(self, x, dual_inp=False, num_layers=3):
        x = self.trunk(x)
        v = x
        if dual_inp:
            x = x[: x.shape[0] // 2]
        outs = [[f(x) for f in self.outputs08]]
        if num_layers >= 2:
            y = self.layer4(x)
            outs.append([f(y) for f in self.outputs16])
        if num_layers >= 3:
            outs.append([f(self.layer5(y)) for f in self.outputs32])
        return tuple(outs) + ((v,) if dual_inp else ())


# ------------------------------------------------------------------------------------------ update block
class FlowHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256, output_dim=2):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, output_dim, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class ConvGRU(nn.Module):
    def __init__(self, hidden_dim, input_dim, kernel_size=3):
        super().__init__()
        p = kernel_size // 2
        self.convz = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=p)
        self.convr = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=p)
        self.convq = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=p)

    def forward(self, h, cz, cr, cq, *x_list):
        x = torch.cat(x_list, dim=1)
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(self.convz(hx) + cz)
        r = torch.sigmoid(self.convr(hx) + cr)
        q = torch.tanh(self.convq(torch.cat([r * h, x], dim=1)) + cq)
        return (1 - z) * h + z * q


class BasicMotionEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1)
        self.convc1 = nn.Conv2d(cor_planes, 64, 1)
        self.convc2 = nn.Conv2d(64, 64, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        self.convf2 = nn.Conv2d(64, 64, 3, padding=1)
        self.conv = nn.Conv2d(128, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc2(F.relu(self.convc1(corr))))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)


def _pool2x(x):
    return F.avg_pool2d(x, 3, stride=2, padding=1)


def _interp(x, dest):
    return F.interpolate(x, dest.shape[2:], mode="bilinear", align_corners=True)


class BasicMultiUpdateBlock(nn.Module):
    def __init__(self, args, hidden_dims=[]):
        super().__init__()
        self.args = args
        self.encoder = BasicMotionEncoder(args)
        self.gru08 = ConvGRU(hidden_dims[2], 128 + hidden_dims[1] * (args.n_gru_layers > 1))
        self.gru16 = ConvGRU(hidden_dims[1], hidden_dims[0] * (args.n_gru_layers == 3) + hidden_dims[2])
        self.gru32 = ConvGRU(hidden_dims[0], hidden_dims[1])
        self.flow_head = FlowHead(hidden_dims[2], hidden_dim=256, output_dim=2)
        factor = 2 ** args.n_downsample
        self.mask = nn.Sequential(nn.Conv2d(hidden_dims[2], 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, (factor ** 2) * 9, 1))

    def forward(self, net, inp, corr=None, flow=None, iter08=True, iter16=True, iter32=True, update=True):
        if iter32:
            net[2] = self.gru32(net[2], *(inp[2]), _pool2x(net[1]))
        if iter16:
            extra = (_interp(net[2], net[1]),) if self.args.n_gru_layers > 2 else ()
            net[1] = self.gru16(net[1], *(inp[1]), _pool2x(net[0]), *extra)
        if iter08:
            motion = self.encoder(flow, corr)
            extra = (_interp(net[1], net[0]),) if self.args.n_gru_layers > 1 else ()
            net[0] = self.gru08(net[0], *(inp[0]), motion, *extra)
        if not update:
            return net
        return net, 0.25 * self.mask(net[0]), self.flow_head(net[0])


def coords_grid(batch, ht, wd, device):
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


class RAFTStereo(nn.Module):
    def __init__(self, args=None, imagenet_norm=False):
        super().__init__()
        self.args = argparse.Namespace(hidden_dims=[128] * 3, corr_implementation="reg", shared_backbone=False,
                                       corr_levels=4, corr_radius=4, n_downsample=2, context_norm="batch",
                                       slow_fast_gru=False, n_gru_layers=3, train_iters=22, valid_iters=32,
                                       mixed_precision=False)
        if args is not None:
            for key, value in vars(args).items():
                setattr(self.args, key, value)
        self.imagenet_norm = imagenet_norm
        a = self.args
        ctx = a.hidden_dims
        self.cnet = MultiBasicEncoder(output_dim=[a.hidden_dims, ctx], norm_fn=a.context_norm, downsample=a.n_downsample)
        self.update_block = BasicMultiUpdateBlock(a, hidden_dims=a.hidden_dims)
        self.context_zqr_convs = nn.ModuleList([nn.Conv2d(ctx[i], a.hidden_dims[i] * 3, 3, padding=1)
                                                for i in range(a.n_gru_layers)])
        if a.shared_backbone:
            self.conv2 = nn.Sequential(ResidualBlock(128, 128, "instance", stride=1), nn.Conv2d(128, 256, 3, padding=1))
        else:
            self.fnet = BasicEncoder(output_dim=256, norm_fn="instance", downsample=a.n_downsample)
        self.freeze_bn()

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def upsample_flow(self, flow, mask):
        """Convex 9-tap upsampling (raft_stereo.py:81-93)."""
        N, D, H, W = flow.shape
        f = 2 ** self.args.n_downsample
        if flow.is_cuda and not (torch.is_grad_enabled() and (flow.requires_grad or mask.requires_grad)) and f in (2, 4, 8):
            from . import ops
            return ops.convex_upsample(flow, mask, f)          # one kernel: softmax + 9-tap combination + pixel shuffle (csrc/upsample2d.cu)
        mask = torch.softmax(mask.contiguous().view(N, 1, 9, f, f, H, W), dim=2)      # (channels-last glue: NCHW order first)
        up = F.unfold(f * flow, [3, 3], padding=1).view(N, D, 9, 1, 1, H, W)
        up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
        return up.reshape(N, D, f * H, f * W)

    def _iteration(self, net_list, inp_list, corr_fn, coords0, coords1):
        """One GRU iteration of raft_stereo.py:153-182: pyramid lookup (one CUDA kernel) + update block (torch)."""
        a = self.args
        coords1 = coords1.detach()
        corr = corr_fn(coords1)
        flow = coords1 - coords0
        if a.n_gru_layers == 3 and a.slow_fast_gru:
            net_list = self.update_block(net_list, inp_list, iter32=True, iter16=False, iter08=False, update=False)
        if a.n_gru_layers >= 2 and a.slow_fast_gru:
            net_list = self.update_block(net_list, inp_list, iter32=a.n_gru_layers == 3, iter16=True, iter08=False,
                                         update=False)
        net_list, up_mask, delta_flow = self.update_block(net_list, inp_list, corr, flow,
                                                          iter32=a.n_gru_layers == 3, iter16=a.n_gru_layers >= 2)
        delta_flow[:, 1] = 0.0                    # stereo: project the update onto the epipolar line
        return net_list, up_mask, coords1 + delta_flow

    def _iterate_graphed(self, net_list, inp_list, corr_fn, coords0, coords1, iters):
        """The 32-iteration loop is launch-bound in eager mode (~60 small kernels per iteration at 1/4 resolution): ONE
        iteration -- lookup kernel + update block + the write-back of its outputs into its own inputs -- is captured
        into a CUDA graph once per input shape and replayed ``iters`` times per call.
        Per call only the graph's static inputs (hidden states, context features, correlation pyramid, coordinates) are
        refreshed.  Same kernels, same order, so the result is bit-identical to the eager loop.
        Opt-in: ``model.cuda_graph = True``."""
        key = (tuple(coords1.shape), tuple(tuple(t.shape) for t in net_list), str(coords1.device))
        cache = self.__dict__.setdefault("_graph_cache", {})
        hit = cache.get(key)
        if hit is None:
            st = dict(net=[t.clone() for t in net_list], inp=[[t.clone() for t in lvl] for lvl in inp_list],
                      coords=coords1.clone(), coords0=coords0.clone(), levels=[t.clone() for t in corr_fn._levels])
            corr_fn._levels = st["levels"]                # the captured lookup reads the static pyramid buffers
            st["corr_fn"] = corr_fn
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up outside capture (cuDNN plans, lazy kernel loading);
                self._iteration(list(st["net"]), st["inp"], corr_fn, st["coords0"], st["coords"])   # update_block mutates the list
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                n2, m2, c2 = self._iteration(list(st["net"]), st["inp"], corr_fn, st["coords0"], st["coords"])
                for dst, src in zip(st["net"], n2):
                    dst.copy_(src)
                st["coords"].copy_(c2)
            st["graph"], st["mask"] = graph, m2
            cache[key] = hit = st
        else:
            for dst, src in zip(hit["net"], net_list):
                dst.copy_(src)
            for dl, sl in zip(hit["inp"], inp_list):
                for dst, src in zip(dl, sl):
                    dst.copy_(src)
            for dst, src in zip(hit["levels"], corr_fn._levels):
                dst.copy_(src)
            hit["coords"].copy_(coords1)
        for _ in range(iters):
            hit["graph"].replay()
        return hit["coords"].clone(), hit["mask"]

    def _iterate_umma(self, net_list, inp_list, corr_fn, coords0, coords1, iters):
        """The GRU loop with the update block on the tensor-core 2-D conv path (update_umma.UmmaRaftUpdate: every conv of
        the block on tcgen05 in the exact 'fp16x2' format, hidden states resident in that layout).
        The default for CUDA inference (``model.update_mode = "auto"``; ``"torch"`` selects the torch block).  With ``model.cuda_graph`` one iteration is
        captured once per shape and replayed.  Returns (coords1, up_mask of the last iterate)."""
        from .update_umma import UmmaRaftUpdate
        upd = self.__dict__.get("_umma_update")
        if upd is None:
            upd = self.__dict__["_umma_update"] = UmmaRaftUpdate(self.update_block, self.args)
        upd._prepare()
        net = [upd.to_cl(t) for t in net_list]
        ctx = upd.context(inp_list)

        def one(net, ctx, fn, c0, c1):
            net, delta = upd.step(net, ctx, fn(c1), c1 - c0)
            delta[:, 1] = 0.0                     # stereo: project the update onto the epipolar line (raft_stereo.py:172)
            return net, c1 + delta

        if not getattr(self, "cuda_graph", False) or iters < 2:
            for _ in range(iters):
                net, coords1 = one(net, ctx, corr_fn, coords0, coords1)
            return coords1, upd.mask(net[0])
        key = ("umma", tuple(coords1.shape), tuple(tuple(t.shape) for t in net), str(coords1.device))
        cache = self.__dict__.setdefault("_graph_cache", {})
        hit = cache.get(key)
        if hit is None:
            st = dict(net=[t.clone() for t in net], ctx=[(a.clone(), b.clone()) for a, b in ctx], coords=coords1.clone(),
                      coords0=coords0.clone(), levels=[t.clone() for t in corr_fn._levels])
            corr_fn._levels = st["levels"]                # the captured lookup reads the static pyramid buffers
            st["corr_fn"] = corr_fn
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up outside capture (kernel plans, lazy module loading)
                one(st["net"], st["ctx"], corr_fn, st["coords0"], st["coords"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                n2, c2 = one(st["net"], st["ctx"], corr_fn, st["coords0"], st["coords"])
                for dst, src in zip(st["net"], n2):
                    dst.copy_(src)
                st["coords"].copy_(c2)
            st["graph"] = graph
            cache[key] = hit = st
        else:
            for dst, src in zip(hit["net"], net):
                dst.copy_(src)
            for (da, db), (sa, sb) in zip(hit["ctx"], ctx):
                da.copy_(sa); db.copy_(sb)
            for dst, src in zip(hit["levels"], corr_fn._levels):
                dst.copy_(src)
            hit["coords"].copy_(coords1)
            hit["coords0"].copy_(coords0)
        for _ in range(iters):
            hit["graph"].replay()
        return hit["coords"].clone(), upd.mask(hit["net"][0])

    def _folded(self, image1):
        """CUDA inference runs on a shadow copy with every eval BatchNorm2d folded into its convolution (glue.py): the context
        encoder's BatchNorm passes are pure memory traffic.  ``model.fold_bn = False`` turns it off."""
        if not getattr(self, "fold_bn", True) or self.training or not image1.is_cuda or torch.is_grad_enabled():
            return None
        from .glue import inference_shadow
        ex = torch.zeros(1, 3, 64, 128, device=image1.device)
        sh = inference_shadow(self, lambda m: m(ex, ex, iters=1))
        for k in ("update_mode", "cuda_graph", "channels_last"):
            if k in self.__dict__:
                sh.__dict__[k] = self.__dict__[k]
        return sh

    def forward(self, image1, image2, iters=None, flow_init=None, test_mode=False):
        sh = self._folded(image1)
        if sh is not None:
            return sh(image1, image2, iters=iters, flow_init=flow_init, test_mode=test_mode)
        a = self.args
        if iters is None:
            iters = a.train_iters if self.training else a.valid_iters       # raft_stereo.py:99-103
        if not self.imagenet_norm:
            mean = torch.tensor([0.485, 0.456, 0.406], device=image1.device).view(1, 3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225], device=image1.device).view(1, 3, 1, 1)
            image1 = 2 * (image1 * std + mean) - 1.0
            image2 = 2 * (image2 * std + mean) - 1.0
        if getattr(self, "channels_last", False):
            image1, image2 = glue_channels_last(self, image1, image2)
        if a.shared_backbone:
            *cnet_list, x = self.cnet(torch.cat((image1, image2), dim=0), dual_inp=True, num_layers=a.n_gru_layers)
            fmap1, fmap2 = self.conv2(x).split(x.shape[0] // 2, dim=0)
        else:
            cnet_list = self.cnet(image1, num_layers=a.n_gru_layers)
            fmap1, fmap2 = self.fnet([image1, image2])
        net_list = [torch.tanh(x[0]) for x in cnet_list]
        inp_list = [torch.relu(x[1]) for x in cnet_list]
        inp_list = [list(conv(i).split(conv.out_channels // 3, dim=1)) for i, conv in zip(inp_list, self.context_zqr_convs)]

        # raft_stereo.py:138-149: "reg" / "reg_cuda" / "alt" / "alt_cuda" are four implementations of ONE function (the
        # alternates trade memory for recomputation: correlation of pooled features == pooled correlation, both linear);
        # here all of them are the all-pairs kernels of csrc/corr1d.cu (33.5 MB of pyramid at 512x1024 does not need "alt")
        if a.corr_implementation not in ("reg", "reg_cuda", "alt", "alt_cuda"):
            raise ValueError(f"unknown corr_implementation {a.corr_implementation!r}")
        # ---- hot path: all-pairs correlation + pyramid (once), lookup (every iteration)
        corr_fn = CorrBlock1D(fmap1.float(), fmap2.float(), radius=a.corr_radius, num_levels=a.corr_levels)

        n, _, h, w = net_list[0].shape
        coords0 = coords_grid(n, h, w, image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init
        flow_up = None
        if self.training:
            # raft_stereo.py:105-107,152-188: every iterate is upsampled and returned.  Gradients reach the feature maps
            # through the differentiable CorrBlock1D (autograd.py: forward kernels of csrc/corr1d.cu, composed adjoints);
            # the update block is torch.  fp32 throughout (the reference's autocast is an optimisation, not semantics).
            preds = []
            for _ in range(iters):
                net_list, up_mask, coords1 = self._iteration(net_list, inp_list, corr_fn, coords0, coords1)
                preds.append(-self.upsample_flow(coords1 - coords0, up_mask)[:, :1])
            return preds
        # update_mode: "auto" (default) = "umma" whenever it applies (CUDA inference, no slow_fast_gru, hidden widths that are
        # whole 16-channel blocks), "torch" = the reference-style torch/cuDNN update block, "umma" = insist
        mode = getattr(self, "update_mode", "auto")
        umma_ok = image1.is_cuda and not a.slow_fast_gru and all(d % 16 == 0 for d in a.hidden_dims) and not torch.is_grad_enabled()
        if mode == "umma" and not umma_ok:
            raise RuntimeError("update_mode='umma' needs CUDA tensors, torch.no_grad(), slow_fast_gru=False and hidden_dims % 16 == 0")
        if mode in ("umma", "auto") and umma_ok:
            coords1, up_mask = self._iterate_umma(net_list, inp_list, corr_fn, coords0, coords1, iters)
            return -self.upsample_flow(coords1 - coords0, up_mask)[:, :1]
        if getattr(self, "cuda_graph", False) and flow_init is None and iters > 1:
            coords1, up_mask = self._iterate_graphed(net_list, inp_list, corr_fn, coords0, coords1, iters)
            flow_up = self.upsample_flow(coords1 - coords0, up_mask)[:, :1]
            return -flow_up
        for itr in range(iters):
            net_list, up_mask, coords1 = self._iteration(net_list, inp_list, corr_fn, coords0, coords1)
            if itr < iters - 1:
                continue
            flow_up = self.upsample_flow(coords1 - coords0, up_mask)[:, :1]
        return -flow_up
