"""GwcNet_G / GwcNet_GC drop-ins (reference: models/GwcNet/gwcnet.py:108-232).

Same constructor, same ``forward(left, right)`` contract, same state-dict names/shapes; the
region between the 2-D feature extractor and the returned disparity -- volume build, dres0/1,
three hourglasses, classif3, trilinear upsample + softmax + regression -- runs in libstb200.so.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .aggregation import convbn_3d, deconvbn_3d, make_backend
from .features2d import GwcFeatures


class hourglass(nn.Module):
    """Parameter container of GwcNet/gwcnet.py:68-93; ``run`` is forward :95-105 on a backend."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(convbn_3d(c, c * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv3 = nn.Sequential(convbn_3d(c * 2, c * 4, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(c * 4, c * 4, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = deconvbn_3d(c * 4, c * 2)
        self.conv6 = deconvbn_3d(c * 2, c)
        self.redir1 = convbn_3d(c, c, 1, 1, 0)
        self.redir2 = convbn_3d(c * 2, c * 2, 1, 1, 0)

    def run(self, be, x):
        c1 = be.conv(self.conv1[0], x, "relu")
        c2 = be.conv(self.conv2[0], c1, "relu")
        c3 = be.conv(self.conv3[0], c2, "relu")
        c4 = be.conv(self.conv4[0], c3, "relu")
        r2 = be.conv(self.redir2, c2)
        c5 = be.conv(self.conv5, c4, "relu", residual=r2)      # relu(conv5(conv4) + redir2(conv2))
        r1 = be.conv(self.redir1, x)
        return be.conv(self.conv6, c5, "relu", residual=r1)    # relu(conv6(conv5) + redir1(x))


def _classif():
    return nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                         nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))


class GwcNet(nn.Module):
    def __init__(self, maxdisp, use_concat_volume=False, precision="auto"):
        super().__init__()
        self.maxdisp = maxdisp
        self.use_concat_volume = use_concat_volume
        self.num_groups = 40
        self.concat_channels = 12 if use_concat_volume else 0
        self.feature_extraction = GwcFeatures(use_concat_volume, 12)
        self.dres0 = nn.Sequential(convbn_3d(self.num_groups + self.concat_channels * 2, 32, 3, 1, 1),
                                   nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.dres4 = hourglass(32)
        self.classif0, self.classif1, self.classif2, self.classif3 = _classif(), _classif(), _classif(), _classif()
        self.feature_tf32 = None
        self.feature_mode = None
        self.set_precision(precision)

    def set_precision(self, precision: str):
        """'fp32' = exact CUDA-core path; 'fp16x2' = exact tensor-core path (operand-split fp16, <= 1e-3 px, ~12x faster);
        'fp16' / 'bf16' = single 16-bit tcgen05 path.  'auto' (the constructor default): 'fp16x2' from the first CUDA
        inference forward on -- a drop-in user gets the fast path that meets the fp32 bar without calling anything --
        and 'fp32' until then."""
        self._auto_precision = precision == "auto"
        if self._auto_precision:
            precision = "fp32"
        self.precision = precision
        self._be = make_backend(precision)
        return self

    def _resolve_auto_precision(self, x):
        if getattr(self, "_auto_precision", False) and x.is_cuda and not self.training:
            self.set_precision("fp16x2")

    def aggregate(self, fl, fr, height, width, be=None, all_heads=False):
        """The hot path: features -> disparity [B,H,W] (training: the reference's list of four)."""
        be = be or self._be
        if "_cl" in fl:        # tensor-core extractor handed over its channels-last outputs
            vol = be.volume_from_cl(fl["_cl"], self.maxdisp // 4, self.num_groups)
        else:
            vol = be.volume_gwc_concat(fl["gwc_feature"], fr["gwc_feature"], fl.get("concat_feature"),
                                       fr.get("concat_feature"), self.maxdisp // 4, self.num_groups)
        c = be.conv(self.dres0[0], vol, "relu")
        cost0 = be.conv(self.dres0[2], c, "relu")
        c = be.conv(self.dres1[0], cost0, "relu")
        cost0 = be.conv(self.dres1[2], c, "none", residual=cost0)
        out1 = self.dres2.run(be, cost0)
        out2 = self.dres3.run(be, out1)
        out3 = self.dres4.run(be, out2)
        if all_heads:       # training: [pred0..pred3] from classif0(cost0), classif1(out1), ... (gwcnet.py:191-216)
            preds = []
            for cls, t in ((self.classif0, cost0), (self.classif1, out1), (self.classif2, out2), (self.classif3, out3)):
                c = be.conv(cls[2], be.conv(cls[0], t, "relu"))
                preds.append(be.head(c, self.maxdisp, height, width, align_corners=False))
            return preds
        c = be.conv(self.classif3[0], out3, "relu")
        cost3 = be.conv(self.classif3[2], c)
        self._last_cost = cost3
        return be.head(cost3, self.maxdisp, height, width, align_corners=False)

    def _features(self, left, right):
        """2-D extractor: outside the hot path, runs through torch/cuDNN.  feature_mode:
          'fp32' exact fp32 convs (what the CPU reference computes; default for precision='fp32', which promises
                 <=1e-3 px), 'tf32' torch's default for convs (what the reference itself does on a GPU),
          'tf32_cl' same arithmetic as 'tf32' but channels-last activations: cuDNN's TF32 kernels are NHWC, and with
                 NCHW tensors ~46% of the extractor's GPU time is nchwToNhwc/nhwcToNchw transposes (profiles/ncu_launches_r01.txt),
          'fp16' channels-last fp16 autocast,
          'umma' the extractor's Conv2d+BN(+ReLU,+residual) layers on the tcgen05 conv kernel in fp16/bf16
                 (features_umma.py).  None = 'fp32' for precision fp32, else 'umma'.
        feature_tf32 (legacy knob): False forces 'fp32'."""
        mode = self.feature_mode
        if self.feature_tf32 is False:
            mode = "fp32"
        if mode is None:
            mode = "fp32" if self.precision == "fp32" else "umma"
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and mode != "fp32"
        try:
            if mode == "umma":
                from .features_umma import UmmaGwcFeatures
                if getattr(self, "_fe_umma", None) is None or self._fe_umma.precision != self._be.name:
                    self._fe_umma = UmmaGwcFeatures(self._be.name)
                self._fe_umma.prof = self._be.prof          # per-layer brackets ("conv2d_umma") when a profiler is attached
                head = (self.concatconv[0], self.concatconv[2]) if hasattr(self, "concatconv") else None
                direct = type(self) is GwcNet and os.environ.get("STB_VOLUME_FROM_NCHW", "0") != "1"
                return self._fe_umma(self.feature_extraction, left, right, concat_head=head, channels_last_out=direct)
            with self._be.prof.bracket("torch_features2d", 0.0, 0.0):
                if mode == "fp16":
                    with torch.autocast("cuda", dtype=torch.float16):
                        fl = self.feature_extraction(left.contiguous(memory_format=torch.channels_last))
                        fr = self.feature_extraction(right.contiguous(memory_format=torch.channels_last))
                    cvt = lambda t: t.float().contiguous()
                    fl = {k: cvt(v) for k, v in fl.items()} if isinstance(fl, dict) else cvt(fl)
                    fr = {k: cvt(v) for k, v in fr.items()} if isinstance(fr, dict) else cvt(fr)
                elif mode == "tf32_cl":
                    if not getattr(self, "_fe_channels_last", False):
                        self.feature_extraction.to(memory_format=torch.channels_last)
                        self._fe_channels_last = True
                    fl = self.feature_extraction(left.contiguous(memory_format=torch.channels_last))
                    fr = self.feature_extraction(right.contiguous(memory_format=torch.channels_last))
                else:
                    fl = self.feature_extraction(left)
                    fr = self.feature_extraction(right)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return fl, fr

    def forward(self, left, right):
        if self.training:
            return self._forward_train(left, right)
        self._resolve_auto_precision(left)
        fl, fr = self._features(left, right)
        return self.aggregate(fl, fr, left.shape[2], left.shape[3])

    def _forward_train(self, left, right):
        """Training step forward (exact fp32): torch 2-D extractor (+concatconv) with autograd, cost-volume path on
        TrainBackend (forward and backward in libstb200.so).  Returns [pred0, pred1, pred2, pred3] (gwcnet.py:216)."""
        from .aggregation import train_backend_for
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            fl = self.feature_extraction(left)
            fr = self.feature_extraction(right)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return self.aggregate(fl, fr, left.shape[2], left.shape[3], be=train_backend_for(self), all_heads=True)


def GwcNet_G(d=192, **kw):
    return GwcNet(d, use_concat_volume=False, **kw)


def GwcNet_GC(d=192, **kw):
    return GwcNet(d, use_concat_volume=True, **kw)
