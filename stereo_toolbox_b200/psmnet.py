"""PSMNet drop-in (reference: models/PSMNet/stackhourglass.py:53-161).

Same constructor / ``forward(left, right)`` / state-dict names; eval output [B,1,H,W].
The inline concat-volume loop (:111-120), dres0..4, classif1..3 (cumulative) and the
trilinear+softmax+regression head run in libstb200.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .aggregation import convbn_3d, deconvbn_3d, make_backend
from .features2d import PsmFeatures


class hourglass(nn.Module):
    """Parameter container of PSMNet/stackhourglass.py:10-29; ``run`` is forward :31-50."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(convbn_3d(c, c * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = convbn_3d(c * 2, c * 2, 3, 1, 1)
        self.conv3 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(c * 2, c * 2, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = deconvbn_3d(c * 2, c * 2)
        self.conv6 = deconvbn_3d(c * 2, c)

    def run(self, be, x, presqu, postsqu, skip):
        """Returns (conv6(...) + skip, pre, post); ``skip`` is the ``+cost0`` of :126-132 fused in."""
        out = be.conv(self.conv1[0], x, "relu")
        pre = be.conv(self.conv2, out, "relu", residual=postsqu)          # relu(conv2 + postsqu) / relu(conv2)
        out = be.conv(self.conv3[0], pre, "relu")
        out = be.conv(self.conv4[0], out, "relu")
        post = be.conv(self.conv5, out, "relu", residual=presqu if presqu is not None else pre)
        out = be.conv(self.conv6, post, "none", residual=skip)
        return out, pre, post


def _classif():
    return nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                         nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))


class PSMNet(nn.Module):
    def __init__(self, maxdisp=192, precision="auto"):
        super().__init__()
        self.maxdisp = maxdisp
        self.feature_extraction = PsmFeatures()
        self.dres0 = nn.Sequential(convbn_3d(64, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                   convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.dres4 = hourglass(32)
        self.classif1, self.classif2, self.classif3 = _classif(), _classif(), _classif()
        self.feature_tf32 = None
        self.feature_mode = None
        self.set_precision(precision)

    def set_precision(self, precision: str):
        """'fp32' | 'fp16x2' | 'fp16' | 'bf16'; 'auto' (constructor default): 'fp16x2', the exact tensor-core path, from the
        first CUDA inference forward on, 'fp32' until then (see GwcNet.set_precision)."""
        self._auto_precision = precision == "auto"
        if self._auto_precision:
            precision = "fp32"
        self.precision = precision
        self._be = make_backend(precision)
        return self

    def aggregate(self, fl, fr, height, width, be=None, all_heads=False):
        be = be or self._be
        vol = be.volume_concat(fl, fr, self.maxdisp // 4, mask_left=True)
        c = be.conv(self.dres0[0], vol, "relu")
        cost0 = be.conv(self.dres0[2], c, "relu")
        c = be.conv(self.dres1[0], cost0, "relu")
        cost0 = be.conv(self.dres1[2], c, "none", residual=cost0)
        out1, pre1, post1 = self.dres2.run(be, cost0, None, None, cost0)
        out2, pre2, post2 = self.dres3.run(be, out1, pre1, post1, cost0)
        out3, pre3, post3 = self.dres4.run(be, out2, pre1, post2, cost0)
        cost1 = be.conv(self.classif1[2], be.conv(self.classif1[0], out1, "relu"))
        cost2 = be.conv(self.classif2[2], be.conv(self.classif2[0], out2, "relu"), residual=cost1)
        cost3 = be.conv(self.classif3[2], be.conv(self.classif3[0], out3, "relu"), residual=cost2)
        self._last_cost = cost3
        if all_heads:       # training: [pred1, pred2, pred3], each [B,1,H,W] (stackhourglass.py:139-159)
            return [be.head(c, self.maxdisp, height, width, align_corners=False).unsqueeze(1) for c in (cost1, cost2, cost3)]
        return be.head(cost3, self.maxdisp, height, width, align_corners=False).unsqueeze(1)

    def _features(self, left, right):
        """2-D extractor: outside the hot path, runs through torch/cuDNN.  feature_mode:
          'fp32' exact fp32 convs (what the CPU reference computes; default for precision='fp32', which promises
                 <=1e-3 px), 'tf32' torch's default for convs (what the reference itself does on a GPU),
          'tf32_cl' same arithmetic as 'tf32' but channels-last activations: cuDNN's TF32 kernels are NHWC, and with
                 NCHW tensors ~46% of the extractor's GPU time is nchwToNhwc/nhwcToNchw transposes (profiles/ncu_launches_r01.txt),
          'fp16' channels-last fp16 autocast,
          'umma' trunk + lastconv on the tcgen05 conv kernel in the storage format of the precision, the SPP branches
                 (pooled to a handful of pixels) in fp32 torch (features_umma.UmmaGwcFeatures.psm).
        None = 'fp32' for precision fp32, 'umma' for 'fp16x2' (exact: the whole model stays inside the fp32 bar), else 'tf32_cl'.
        feature_tf32 (legacy knob): False forces 'fp32'."""
        mode = self.feature_mode
        if self.feature_tf32 is False:
            mode = "fp32"
        if mode is None:
            mode = "fp32" if self.precision == "fp32" else ("umma" if self.precision == "fp16x2" else "tf32_cl")
        if mode == "umma":
            from .features_umma import UmmaGwcFeatures
            if getattr(self, "_fe_umma", None) is None or self._fe_umma.precision != self._be.name:
                self._fe_umma = UmmaGwcFeatures(self._be.name)
            self._fe_umma.prof = self._be.prof
            return self._fe_umma.psm(self.feature_extraction, left, right)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and mode != "fp32"
        try:
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
        if getattr(self, "_auto_precision", False) and left.is_cuda:
            self.set_precision("fp16x2")
        fl, fr = self._features(left, right)
        return self.aggregate(fl, fr, left.shape[2], left.shape[3])

    def _forward_train(self, left, right):
        """Training step forward (exact fp32): the 2-D extractor is ordinary torch (train-mode BatchNorm2d, autograd),
        the cost-volume path runs on TrainBackend -- forward and backward in libstb200.so.  Returns the reference's
        list [pred1, pred2, pred3] (stackhourglass.py:159)."""
        from .aggregation import train_backend_for
        prec = getattr(self, "train_precision", "fp32")
        if getattr(self, "train_features", "fp32") == "amp" and prec in ("bf16", "fp16"):
            # the reference's own mixed-precision recipe for the torch part (trainer/trainer_torchrun.py:219,274: the whole step
            # runs under torch.amp.autocast): the 2-D extractor's cuDNN convs in the training dtype on channels-last
            # tensors, BatchNorm2d statistics in fp32 (autocast policy); opt-in, ``model.train_features = "amp"``
            if not getattr(self, "_fe_channels_last", False):
                self.feature_extraction.to(memory_format=torch.channels_last)
                self._fe_channels_last = True
            with torch.autocast("cuda", dtype=torch.bfloat16 if prec == "bf16" else torch.float16):
                fl = self.feature_extraction(left.contiguous(memory_format=torch.channels_last))
                fr = self.feature_extraction(right.contiguous(memory_format=torch.channels_last))
            fl, fr = fl.float(), fr.float()
            return self.aggregate(fl, fr, left.shape[2], left.shape[3], be=train_backend_for(self), all_heads=True)
        if getattr(self, "train_features", "fp32") == "tf32":
            # torch's own default for convolutions on a GPU (cudnn.allow_tf32), i.e. what the reference trainer computes
            # without amp; channels-last so that cuDNN's NHWC tensor-core kernels run without layout transposes
            if not getattr(self, "_fe_channels_last", False):
                self.feature_extraction.to(memory_format=torch.channels_last)
                self._fe_channels_last = True
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = True
            try:
                fl = self.feature_extraction(left.contiguous(memory_format=torch.channels_last))
                fr = self.feature_extraction(right.contiguous(memory_format=torch.channels_last))
            finally:
                torch.backends.cudnn.allow_tf32 = prev
            return self.aggregate(fl, fr, left.shape[2], left.shape[3], be=train_backend_for(self), all_heads=True)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = prev and self.feature_mode not in (None, "fp32") and self.feature_tf32 is not False
        try:
            fl = self.feature_extraction(left)
            fr = self.feature_extraction(right)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        return self.aggregate(fl, fr, left.shape[2], left.shape[3], be=train_backend_for(self), all_heads=True)
