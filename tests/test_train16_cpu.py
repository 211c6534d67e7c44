"""Host logic of the mixed-precision training backend (stereo_toolbox_b200/train16.py) on CPU.

``Umma16TrainBackend`` has two kernel-backed primitives -- ``_raw_conv`` (tcgen05 conv, used for the forward AND, through
an adjoint module, for the data gradient) and ``_wgrad``.  Here they are replaced by torch stand-ins that keep the
backend's tensor conventions (channels-last storage dtype, zero-padded channels, fp32 classifier output); everything
else -- the autograd Function, adjoint construction (flipped taps / ConvTranspose3d with inferred output_padding /
Conv3d for transposed layers), channel padding of gradients, BatchNorm3d batch statistics on the 16-bit tensors, the
fp32 volume / head boundary -- is the product's own code and is compared with the REFERENCE's training step
(tests/golden/psmnet_train.npz: predictions, loss, parameter gradients)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import GOLDEN, golden_state, load_golden, load_meta
from oracle import ref_ops as R


def _make_backend(dtype_name):
    from stereo_toolbox_b200.train16 import Umma16TrainBackend

    class TorchPrimitives(Umma16TrainBackend):
        """_raw_conv / _wgrad by torch (fp32 math on the 16-bit operands, like the kernels: fp32 accumulation)."""

        def _raw_conv(self, conv, x16):
            tr = isinstance(conv, nn.ConvTranspose3d)
            cin = conv.weight.shape[0] if tr else conv.weight.shape[1]
            assert x16.dtype == self.dtype and x16.is_contiguous() and x16.shape[-1] >= cin
            assert x16.shape[-1] in (16, 32, 64) or x16.shape[-1] % 64 == 0            # what the kernel accepts
            assert not x16[..., cin:].any()                                            # padding channels are zero
            x = x16[..., :cin].permute(0, 4, 1, 2, 3).float()
            w = conv.weight.detach().to(self.dtype).float()
            if tr:
                y = F.conv_transpose3d(x, w, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
            else:
                y = F.conv3d(x, w, stride=conv.stride, padding=conv.padding)
            y = y.permute(0, 2, 3, 4, 1).contiguous()
            return y if y.shape[-1] < 8 else y.to(self.dtype)

        def _wgrad(self, conv, x16, gy):
            tr = isinstance(conv, nn.ConvTranspose3d)
            cin = conv.weight.shape[0] if tr else conv.weight.shape[1]
            x = x16[..., :cin].permute(0, 4, 1, 2, 3).float()
            g = gy.permute(0, 4, 1, 2, 3).float()
            w = conv.weight.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                if tr:
                    y = F.conv_transpose3d(x, w, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
                else:
                    y = F.conv3d(x, w, stride=conv.stride, padding=conv.padding)
                (gw,) = torch.autograd.grad(y, w, g)
            return gw

        # the fp32 boundary ops of autograd.py need the CUDA library: differentiable oracle forms here
        def volume_concat(self, l, r, maxdisp4, mask_left=True, att_prob=None):
            return self._to_cl(R.build_concat_volume(l, r, maxdisp4, mask_left))

        def volume_gwc_concat(self, gwc_l, gwc_r, cat_l, cat_r, maxdisp4, groups):
            vol = R.build_gwc_volume(gwc_l, gwc_r, maxdisp4, groups)
            if cat_l is not None:
                vol = torch.cat((vol, R.build_concat_volume(cat_l, cat_r, maxdisp4, True)), 1)
            return self._to_cl(vol)

        def head(self, cost, maxdisp, H, W, align_corners=False):
            B, D, h, w, _ = cost.shape
            return R.upsample_softargmin(cost.reshape(B, D, h, w), maxdisp, H, W, align_corners, False)

    return TorchPrimitives(dtype_name)


@pytest.mark.parametrize("k,s,p,tr,op,dims", [(3, 1, 1, False, 0, (4, 6, 8)), (1, 1, 0, False, 0, (4, 6, 8)),
                                               (3, 2, 1, False, 0, (4, 6, 8)), (3, 2, 1, True, 1, (2, 3, 4))])
def test_raw_conv_gradients_match_torch(k, s, p, tr, op, dims):
    """_RawConvFn: data gradient through the adjoint module, weight gradient through _wgrad -- vs torch autograd of the
    same convolution on the same 16-bit operands."""
    from stereo_toolbox_b200.train16 import _RawConvFn
    be = _make_backend("bf16")
    torch.manual_seed(0)
    cin, cout = 16, 32
    conv = (nn.ConvTranspose3d(cin, cout, k, s, p, output_padding=op, bias=False) if tr
            else nn.Conv3d(cin, cout, k, s, p, bias=False))
    x = torch.randn(2, *dims, cin).to(torch.bfloat16).requires_grad_(True)
    y = _RawConvFn.apply(x, conv.weight, be, conv)
    gy = torch.randn_like(y.float()).to(y.dtype)
    y.backward(gy)
    # torch reference on the same rounded operands
    x2 = x.detach().float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    w2 = conv.weight.detach().to(torch.bfloat16).float().requires_grad_(True)
    y2 = (F.conv_transpose3d(x2, w2, stride=s, padding=p, output_padding=op) if tr else F.conv3d(x2, w2, stride=s, padding=p))
    y2.backward(gy.float().permute(0, 4, 1, 2, 3))
    torch.testing.assert_close(y.float(), y2.permute(0, 2, 3, 4, 1).to(torch.bfloat16).float(), rtol=0, atol=0)
    torch.testing.assert_close(x.grad.float(), x2.grad.permute(0, 2, 3, 4, 1), rtol=2e-2, atol=2e-2)   # gx stored in bf16
    torch.testing.assert_close(conv.weight.grad, w2.grad, rtol=1e-4, atol=1e-4)


def test_padded_input_channels_and_classifier_shapes():
    """(i) an input whose channel count was zero-padded to a legal K width (GwcNet_G's 40-group volume -> 64): the data
    gradient comes back in the padded shape with zeros in the padding; (ii) the 1-channel classifier: fp32 output, fp32
    output gradient re-entering the kernel as a 16-channel 16-bit operand."""
    from stereo_toolbox_b200.train16 import _RawConvFn
    be = _make_backend("bf16")
    torch.manual_seed(1)
    conv = nn.Conv3d(40, 32, 3, 1, 1, bias=False)
    vol = torch.randn(1, 40, 4, 6, 8)
    x = be._to_cl(vol).requires_grad_(True)
    assert x.shape == (1, 4, 6, 8, 64) and not x[..., 40:].any()
    y = _RawConvFn.apply(x, conv.weight, be, conv)
    y.float().sum().backward()
    assert x.grad.shape == x.shape and not x.grad[..., 40:].any() and x.grad[..., :40].abs().sum() > 0
    cls = nn.Conv3d(32, 1, 3, 1, 1, bias=False)
    h = torch.randn(1, 4, 6, 8, 32).to(torch.bfloat16).requires_grad_(True)
    c = _RawConvFn.apply(h, cls.weight, be, cls)
    assert c.dtype == torch.float32 and c.shape == (1, 4, 6, 8, 1)
    c.square().sum().backward()
    h2 = h.detach().float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    w2 = cls.weight.detach().to(torch.bfloat16).float().requires_grad_(True)
    F.conv3d(h2, w2, padding=1).square().sum().backward()
    torch.testing.assert_close(h.grad.float(), h2.grad.permute(0, 2, 3, 4, 1), rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(cls.weight.grad, w2.grad, rtol=1e-4, atol=1e-4)


def test_adjoint_modules_are_cached_and_follow_weight_updates():
    be = _make_backend("bf16")
    conv = nn.Conv3d(16, 32, 3, 1, 1, bias=False)
    a1 = be._adjoint(conv, (1, 4, 4, 4, 16), (1, 4, 4, 4, 32))
    assert be._adjoint(conv, (1, 4, 4, 4, 16), (1, 4, 4, 4, 32)) is a1
    with torch.no_grad():
        conv.weight.mul_(2.0)                       # an optimizer step bumps the version
    a2 = be._adjoint(conv, (1, 4, 4, 4, 16), (1, 4, 4, 4, 32))
    assert a2 is a1                                 # same module (stable id for the kernel-plan cache), refreshed in place
    torch.testing.assert_close(a2.weight, conv.weight.detach().flip(2, 3, 4).transpose(0, 1))


@pytest.mark.parametrize("prec,pred_tol,cos_min", [("bf16", 0.15, 0.93), ("fp16", 0.02, 0.99)])
def test_psmnet_16bit_training_step_vs_reference(prec, pred_tol, cos_min):
    """One PSMNet training step on the 16-bit training backend (torch primitives) vs the reference's fp32 step:
    predictions within the storage error, loss within 1 %, parameter gradients aligned.  The two storage formats bracket
    the wiring: with fp16's 11-bit mantissa every checked gradient has cosine >= 0.994 to the reference's, so what is
    left at bf16 (0.95-0.98 on the deepest layers, back-propagated through ~25 layers of 8-bit-mantissa storage) is
    rounding noise of the format, not a wiring error."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_state_dict, synth_pair, synth_gt
    g = load_golden("psmnet_train.npz")
    meta = load_meta("models.json")["psmnet"]
    tmpl = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
            for k, s in meta["keys"].items()}
    z = np.load(f"{GOLDEN}/bn_calib_psmnet.npz")
    net = S.PSMNet(32)
    net.load_state_dict(synth_state_dict(tmpl, 0, {k: z[k] for k in z.files}), strict=True)
    net.train()
    net.train_precision = prec
    net.__dict__["_train16"] = _make_backend(prec)
    left, right = synth_pair(2, 256, 256, seed=1, shift=7)
    gt = synth_gt(2, 256, 256)
    preds = net(left, right)
    assert len(preds) == 3 and all(p.shape == (2, 1, 256, 256) for p in preds)
    mask = (gt > 0) & (gt < 32)
    loss = sum(w * F.smooth_l1_loss(p.squeeze(1)[mask], gt[mask], reduction="mean") for w, p in zip((0.5, 0.7, 1.0), preds))
    loss.backward()
    for i, p in enumerate(preds):
        assert (p.detach()[:, :, ::2, ::2] - g[f"pred{i + 1}"]).abs().mean().item() < pred_tol
    assert abs(loss.item() - g["loss"].item()) < 0.01 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten(), g["grad:" + name].flatten()
        cos = F.cosine_similarity(got, want, dim=0).item()
        assert cos > cos_min, (name, cos)
        assert 0.9 < (got.norm() / want.norm()).item() < 1.1, name


def test_gwcnet_g_16bit_training_matches_the_fp32_training_path():
    """GwcNet_G (40-group volume zero-padded to 64 channels, 1x1x1 redir convs, four heads) on the 16-bit training backend
    (torch primitives, fp16 storage) vs the same drop-in on the exact-path stand-in (tests/oracle_backend.py
    OracleTrainBackend): four predictions within the storage error, gradients aligned."""
    import stereo_toolbox_b200 as S
    from oracle_backend import oracle_hot_path
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    sd, meta = golden_state("gwcnet_g")
    left, right = synth_pair(2, 64, 128, seed=0, shift=5)
    gt = synth_gt(2, 64, 128)
    mask = (gt > 0) & (gt < meta["maxdisp"])

    def step(prec):
        net = S.GwcNet_G(meta["maxdisp"])
        net.load_state_dict(sd, strict=True)
        net.train()
        if prec != "fp32":
            net.train_precision = prec
            net.__dict__["_train16"] = _make_backend(prec)
        preds = net(left, right)
        loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
        loss.backward()
        return net, preds, loss

    with oracle_hot_path():
        import stereo_toolbox_b200.aggregation as agg
        from oracle_backend import OracleTrainBackend
        old, agg.TrainBackend = agg.TrainBackend, OracleTrainBackend          # train_backend_for() builds agg.TrainBackend
        try:
            ref_net, ref_preds, ref_loss = step("fp32")
        finally:
            agg.TrainBackend = old
    net, preds, loss = step("fp16")
    assert len(preds) == len(ref_preds) == 4                                   # gwcnet.py:216
    for p, q in zip(preds, ref_preds):
        assert (p.detach() - q.detach()).abs().mean().item() < 0.02
    assert abs(loss.item() - ref_loss.item()) < 0.01 * abs(ref_loss.item())
    ref_params = dict(ref_net.named_parameters())
    for name, p in net.named_parameters():
        if p.grad is None or p.dim() < 4 or not name.startswith(("dres", "classif")):
            continue
        cos = F.cosine_similarity(p.grad.flatten(), ref_params[name].grad.flatten(), dim=0).item()
        assert cos > 0.98, (name, cos)


def test_kernel_plans_build_for_every_forward_and_adjoint_conv():
    """Host-side half of the GPU path: for every conv flavour of PSMNet / GwcNet the tcgen05 plan (weight tiles, tap / class
    tables: aggregation_umma.UmmaPlan, pure host code) builds for the forward conv AND for the adjoint module that
    ``Umma16TrainBackend._adjoint`` derives from it -- none falls back to the CUDA-core companion."""
    from stereo_toolbox_b200.aggregation_umma import UmmaPlan
    from stereo_toolbox_b200.train16 import Umma16TrainBackend
    be = Umma16TrainBackend("bf16")
    cases = [(nn.Conv3d(64, 32, 3, 1, 1, bias=False), (1, 8, 16, 16, 64), (1, 8, 16, 16, 32), nn.Conv3d),
             (nn.Conv3d(32, 64, 3, 2, 1, bias=False), (1, 8, 16, 16, 32), (1, 4, 8, 8, 64), nn.ConvTranspose3d),
             (nn.ConvTranspose3d(64, 32, 3, 2, 1, output_padding=1, bias=False), (1, 4, 8, 8, 64), (1, 8, 16, 16, 32), nn.Conv3d),
             (nn.Conv3d(32, 32, 1, 1, 0, bias=False), (1, 8, 16, 16, 32), (1, 8, 16, 16, 32), nn.Conv3d),
             (nn.Conv3d(32, 1, 3, 1, 1, bias=False), (1, 8, 16, 16, 32), (1, 8, 16, 16, 1), nn.Conv3d),
             (nn.Conv3d(128, 128, 3, 1, 1, bias=False), (1, 2, 4, 4, 128), (1, 2, 4, 4, 128), nn.Conv3d),
             (nn.Conv3d(64, 128, 3, 2, 1, bias=False), (1, 4, 8, 8, 64), (1, 2, 4, 4, 128), nn.ConvTranspose3d),
             (nn.ConvTranspose3d(128, 64, 3, 2, 1, output_padding=1, bias=False), (1, 2, 4, 4, 128), (1, 4, 8, 8, 64), nn.Conv3d)]
    for conv, xs, ys in [(c, x, y) for c, x, y, _ in cases]:
        assert UmmaPlan(conv, None, xs[-1], torch.bfloat16).umma_ok
    for conv, xs, ys, kind in cases:
        adj = be._adjoint(conv, xs, ys)
        assert type(adj) is kind
        if kind is nn.ConvTranspose3d:
            assert adj.output_padding[0] == 1
        g_channels = be._as_operand(torch.zeros(ys, dtype=torch.bfloat16)).shape[-1]      # the classifier's 1 -> 16
        assert UmmaPlan(adj, None, g_channels, torch.bfloat16).umma_ok
