"""Tensor-core (tcgen05) path: channels-last bf16 volume builder, implicit-GEMM conv family and the
bf16 drop-in models, against the oracle.  Tolerances: bf16 storage (2^-9 relative per tensor) with fp32
accumulation; model level <=1e-2 px EPE (BASELINE.json north_star)."""
import pytest
import torch
import torch.nn as nn

from conftest import load_golden, golden_state
from oracle import ref_ops as R
from oracle import ref_models as M

pytestmark = pytest.mark.gpu


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def bf(t):
    return t.to(torch.bfloat16).float()


def test_layout_roundtrip():
    from stereo_toolbox_b200.aggregation_umma import to_channels_last_bf16, from_channels_last_bf16
    x = rnd(0, 2, 40, 3, 5, 7)
    cl = to_channels_last_bf16(x.cuda(), 48)
    assert cl.shape == (2, 3, 5, 7, 48)
    torch.testing.assert_close(cl.float().cpu()[..., :40], bf(x).permute(0, 2, 3, 4, 1), rtol=0, atol=0)
    assert cl[..., 40:].abs().max().item() == 0
    torch.testing.assert_close(from_channels_last_bf16(cl, 40).cpu(), bf(x), rtol=0, atol=0)


@pytest.mark.parametrize("G,Cg,Cc,W,D", [(40, 320, 12, 45, 12), (40, 320, 0, 33, 8), (0, 0, 32, 40, 16), (8, 96, 0, 20, 6)])
def test_volume_channels_last(G, Cg, Cc, W, D):
    from stereo_toolbox_b200.aggregation_umma import UmmaBackend, from_channels_last_bf16
    B, H = 2, 5
    be = UmmaBackend()
    if G:
        gl, gr = rnd(1, B, Cg, H, W), rnd(2, B, Cg, H, W)
        cl, cr = (rnd(3, B, Cc, H, W), rnd(4, B, Cc, H, W)) if Cc else (None, None)
        vol = be.volume_gwc_concat(gl.cuda(), gr.cuda(), None if cl is None else cl.cuda(),
                                   None if cr is None else cr.cuda(), D, G)
        want = R.build_gwc_volume(gl, gr, D, G)
        if Cc:
            want = torch.cat((want, R.build_concat_volume(cl, cr, D, True)), 1)
    else:
        cl, cr = rnd(3, B, Cc, H, W), rnd(4, B, Cc, H, W)
        vol = be.volume_concat(cl.cuda(), cr.cuda(), D, True)
        want = R.build_concat_volume(cl, cr, D, True)
    ct = want.shape[1]
    got = from_channels_last_bf16(vol, ct).cpu()
    torch.testing.assert_close(got, bf(want), rtol=1e-2, atol=1e-3)
    if vol.shape[-1] > ct:
        assert vol[..., ct:].abs().max().item() == 0


UCONVS = [
    # cin, cout, k, stride, pad, transposed, act, residual, (D,H,W)
    (32, 32, 3, 1, 1, False, "relu", False, (6, 9, 37)),
    (64, 32, 3, 1, 1, False, "relu", False, (5, 8, 31)),
    (32, 64, 3, 1, 1, False, "none", True, (4, 17, 30)),
    (48, 32, 3, 1, 1, False, "relu", False, (4, 6, 20)),
    (16, 16, 3, 1, 1, False, "leaky", False, (4, 6, 20)),
    (32, 1, 3, 1, 1, False, "none", False, (6, 9, 37)),
    (32, 32, 1, 1, 0, False, "none", False, (4, 10, 33)),
    (64, 32, 3, 2, 1, True, "relu", True, (3, 5, 17)),
    (64, 64, 3, 1, 1, False, "relu", False, (4, 6, 20)),
    (128, 128, 3, 1, 1, False, "relu", False, (3, 6, 20)),    # companion kernel (weights not resident yet)
    (32, 64, 3, 2, 1, False, "relu", False, (6, 10, 22)),     # strided conv: companion kernel
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,tr,act,res,dims", UCONVS)
def test_conv_family_bf16(cin, cout, k, stride, pad, tr, act, res, dims):
    from stereo_toolbox_b200.aggregation_umma import UmmaBackend, to_channels_last_bf16, from_channels_last_bf16
    D, H, W = dims
    B = 2
    x = bf(rnd(1, B, cin, D, H, W))
    if tr:
        conv = nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=pad, output_padding=1, bias=False)
    else:
        conv = nn.Conv3d(cin, cout, k, stride, pad, bias=False)
    bn = nn.BatchNorm3d(cout)
    with torch.no_grad():
        conv.weight.copy_(rnd(2, *conv.weight.shape) * (2.0 / (cin * k ** 3)) ** 0.5)
        bn.weight.copy_(0.75 + 0.5 * torch.rand(cout)); bn.bias.copy_(0.1 * rnd(3, cout))
        bn.running_mean.copy_(0.1 * rnd(4, cout)); bn.running_var.copy_(0.5 + torch.rand(cout))
    layer = nn.Sequential(conv, bn).eval()
    bnd = dict(weight=bn.weight.detach(), bias=bn.bias.detach(), running_mean=bn.running_mean, running_var=bn.running_var)
    want0 = R.conv3d_bn_act(x, conv.weight.detach(), bnd, stride, pad, "none", None, tr, 1 if tr else 0)
    resid = bf(rnd(5, *want0.shape)) if res else None
    want = R.conv3d_bn_act(x, conv.weight.detach(), bnd, stride, pad, act, resid, tr, 1 if tr else 0)
    be = UmmaBackend()
    layer = layer.cuda()
    xcl = to_channels_last_bf16(x.cuda())
    rcl = None if resid is None else to_channels_last_bf16(resid.cuda())
    got = be.conv(layer, xcl, act, rcl)
    assert got.shape == (B,) + tuple(want.shape[2:]) + (cout,)
    if got.dtype == torch.float32:
        got = got.permute(0, 4, 1, 2, 3).cpu()
    else:
        got = from_channels_last_bf16(got).cpu()
    # weights are rounded to bf16 on the tensor-core path: allow ~2^-8 relative of the typical magnitude
    err = (got - want).abs()
    scale = want.abs().mean().item() + 1e-3
    assert err.max().item() < 0.06 * max(1.0, want.abs().max().item()), f"max err {err.max().item()}"
    assert err.mean().item() < 0.01 * scale + 2e-3, f"mean err {err.mean().item()} (scale {scale})"


def _pair(meta):
    from stereo_toolbox_b200.synth import synth_pair
    b, h, w = meta["shape"]
    return synth_pair(b, h, w, seed=1 if h == 256 else 0, shift=meta["shift"])


@pytest.mark.parametrize("key", ["gwcnet_gc", "gwcnet_g"])
def test_gwcnet_golden_bf16(key):
    import stereo_toolbox_b200 as S
    g = load_golden(f"{key}.npz")
    sd, meta = golden_state(key)
    net = (S.GwcNet_GC if key == "gwcnet_gc" else S.GwcNet_G)(meta["maxdisp"], precision="bf16")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = _pair(meta)
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    epe = (disp - g["disp"]).abs().mean().item()
    print(f"{key} bf16 EPE vs reference: {epe:.4e} px")
    if "cost3" in g:
        c = net._last_cost.cpu().permute(0, 4, 1, 2, 3)
        print("cost3 max abs err", (c - g["cost3"]).abs().max().item(), "of range", g["cost3"].abs().max().item())
    assert epe < 1e-2, f"EPE vs reference {epe}"


def test_psmnet_golden_bf16():
    import stereo_toolbox_b200 as S
    g = load_golden("psmnet.npz")
    sd, meta = golden_state("psmnet")
    net = S.PSMNet(meta["maxdisp"], precision="bf16")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = _pair(meta)
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    epe = (disp - g["disp"]).abs().mean().item()
    print(f"psmnet bf16 EPE vs reference: {epe:.4e} px")
    assert disp.shape == g["disp"].shape
    assert epe < 1e-2, f"EPE vs reference {epe}"
