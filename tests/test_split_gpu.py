"""The exact tensor-core path ('fp16x2': operand-split fp16 storage, three tcgen05 MMAs per K-step) against the fp32
oracle.  Bar: fp32-level accuracy -- per layer <= 2e-6 of the output scale, model level <= 1e-3 px EPE vs the reference
(BASELINE.json north_star, fp32 bar)."""
import pytest
import torch
import torch.nn as nn

from conftest import load_golden, golden_state
from oracle import ref_ops as R
from oracle import ref_models as M

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_split_layout_roundtrip_and_pack():
    """Device layout kernels agree bit for bit with the host-side split_pack (the format the weights are packed in)."""
    from stereo_toolbox_b200.aggregation_umma import to_channels_last, from_channels_last, split_pack, split_unpack
    x = rnd(0, 2, 40, 3, 5, 7) * 3.0
    cl = to_channels_last(x.cuda(), 64, torch.float16, split=True)
    assert cl.shape == (2, 3, 5, 7, 128) and cl.dtype == torch.float16
    xp = torch.zeros(2, 3, 5, 7, 64)
    xp[..., :40] = x.permute(0, 2, 3, 4, 1)
    want = split_pack(xp)
    assert torch.equal(cl.cpu(), want)
    back = from_channels_last(cl, 40, split=True).cpu()
    assert torch.equal(back, split_unpack(want)[..., :40].permute(0, 4, 1, 2, 3))
    # 22 mantissa bits while lo is a normal fp16; below that lo is an fp16 subnormal (spacing 2^-24)
    assert ((back - x).abs() <= 2.0 ** -21 * x.abs() + 2.0 ** -24).all()
    # 3-channel image path (small-C kernel)
    img = rnd(1, 4, 3, 16, 24)
    cl = to_channels_last(img.cuda(), 16, torch.float16, split=True)
    ip = torch.zeros(4, 16, 24, 16)
    ip[..., :3] = img.permute(0, 2, 3, 1)
    assert torch.equal(cl.cpu(), split_pack(ip))


@pytest.mark.parametrize("G,Cg,Cc,W,D", [(40, 320, 12, 45, 12), (40, 320, 0, 33, 8), (0, 0, 32, 40, 16), (8, 96, 0, 20, 6)])
def test_split_volume(G, Cg, Cc, W, D):
    from stereo_toolbox_b200.aggregation_umma import UmmaBackend, from_channels_last
    B, H = 2, 5
    be = UmmaBackend("fp16x2")
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
    assert vol.shape[-1] % 32 == 0
    got = from_channels_last(vol, ct, split=True).cpu()
    torch.testing.assert_close(got, want, rtol=2e-6, atol=2e-6)
    full = from_channels_last(vol, split=True).cpu()
    if full.shape[1] > ct:
        assert full[:, ct:].abs().max().item() == 0


SCONVS = [
    # cin, cout, k, stride, pad, transposed, act, residual, (D,H,W)
    (32, 32, 3, 1, 1, False, "relu", False, (6, 9, 37)),
    (64, 32, 3, 1, 1, False, "relu", False, (5, 8, 31)),        # two K-chunks (fp32 workspace)
    (32, 64, 3, 1, 1, False, "none", True, (4, 17, 30)),
    (64, 32, 3, 1, 1, False, "relu", True, (4, 6, 65)),
    (16, 16, 3, 1, 1, False, "leaky", False, (4, 6, 20)),       # one (hi, lo) slice pair per row, ragged 16-channel epilogue
    (32, 1, 3, 1, 1, False, "none", False, (6, 9, 37)),         # classifier, fp32 output
    (32, 32, 1, 1, 0, False, "none", False, (4, 10, 33)),
    (64, 32, 3, 2, 1, True, "relu", True, (3, 5, 17)),          # merged transposed conv, paired stores
    (64, 64, 3, 1, 1, False, "relu", False, (4, 6, 20)),
    (128, 128, 3, 1, 1, False, "relu", True, (3, 6, 20)),
    (128, 64, 3, 2, 1, True, "relu", True, (2, 3, 9)),
    (32, 64, 3, 2, 1, False, "relu", False, (6, 10, 22)),       # strided conv: parity sub-tiles
    (64, 128, 3, 2, 1, False, "relu", False, (4, 8, 70)),
    (32, 32, 3, 1, 1, False, "mish", True, (3, 20, 64)),
    (64, 64, 1, 1, 0, False, "none", True, (4, 10, 33)),        # k1, both K-chunks accumulated in TMEM (3-D pseudo-depth chunks)
    (64, 32, 3, 2, 1, True, "relu", True, (9, 13, 70)),         # merged transposed conv on 16-channel slices, several tiles / depth chunks
    (64, 32, 3, 2, 1, True, "none", False, (3, 5, 17)),         # same plan through the generic epilogue
    (128, 64, 3, 2, 1, True, "relu", True, (7, 11, 40)),        # 4 K-chunks as two passes of two chunks (fp32 partial between them), 4 slices
    (192, 32, 3, 2, 1, True, "relu", True, (3, 5, 17)),         # three passes: the middle one reads AND writes the partial
    (128, 64, 3, 2, 1, True, "mish", False, (3, 5, 17)),        # last pass through the generic epilogue (PCWNet's Mish)
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,tr,act,res,dims", SCONVS)
def test_split_conv_family(cin, cout, k, stride, pad, tr, act, res, dims):
    from stereo_toolbox_b200.aggregation_umma import UmmaBackend, to_channels_last, from_channels_last
    D, H, W = dims
    B = 2
    x = rnd(1, B, cin, D, H, W)
    opad = 1 if (tr and k == 3) else 0
    if tr:
        conv = nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=pad, output_padding=opad, bias=False)
    else:
        conv = nn.Conv3d(cin, cout, k, stride, pad, bias=False)
    bn = nn.BatchNorm3d(cout)
    with torch.no_grad():
        conv.weight.copy_(rnd(2, *conv.weight.shape) * (2.0 / (cin * k ** 3)) ** 0.5)
        bn.weight.copy_(0.75 + 0.5 * torch.rand(cout)); bn.bias.copy_(0.1 * rnd(3, cout))
        bn.running_mean.copy_(0.1 * rnd(4, cout)); bn.running_var.copy_(0.5 + torch.rand(cout))
    layer = nn.Sequential(conv, bn).eval()
    bnd = dict(weight=bn.weight.detach(), bias=bn.bias.detach(), running_mean=bn.running_mean, running_var=bn.running_var)
    want0 = R.conv3d_bn_act(x.double(), conv.weight.detach().double(), {k_: v.double() for k_, v in bnd.items()}, stride, pad,
                            "none", None, tr, opad)
    resid = rnd(5, *want0.shape) if res else None
    want = R.conv3d_bn_act(x.double(), conv.weight.detach().double(), {k_: v.double() for k_, v in bnd.items()}, stride, pad,
                           act, None if resid is None else resid.double(), tr, opad).float()
    be = UmmaBackend("fp16x2")
    layer = layer.cuda()
    xcl = to_channels_last(x.cuda(), None, torch.float16, split=True)
    rcl = None if resid is None else to_channels_last(resid.cuda(), None, torch.float16, split=True)
    got = be.conv(layer, xcl, act, rcl)
    if got.dtype == torch.float32:
        assert got.shape == (B,) + tuple(want.shape[2:]) + (cout,)
        got = got.permute(0, 4, 1, 2, 3).cpu()
    else:
        assert got.shape == (B,) + tuple(want.shape[2:]) + (2 * cout,)
        got = from_channels_last(got, split=True).cpu()
    err = (got - want).abs()
    scale = max(1.0, want.abs().max().item())
    print(f"[fp16x2] {cin}->{cout} k{k} s{stride} tr={tr} {act}: max err {err.max().item():.3e}, mean {err.mean().item():.3e} (scale {scale:.2f})")
    # fp32-level: operands carry 22 bits, accumulation is fp32 in TMEM, the stored output carries 22 bits
    assert err.max().item() < 4e-6 * scale and err.mean().item() < 4e-7 * scale


def _pair(meta):
    from stereo_toolbox_b200.synth import synth_pair
    b, h, w = meta["shape"]
    return synth_pair(b, h, w, seed=1 if h == 256 else 0, shift=meta["shift"])


@pytest.mark.parametrize("feature_mode", ["fp32", "umma"])
def test_gwcnet_gc_golden_fp16x2(feature_mode):
    """Drop-in GwcNet_GC on the exact tensor-core path vs the REFERENCE's output (fixture): the fp32 bar, with the 2-D
    extractor exact (torch fp32) and on the same split tcgen05 kernel."""
    import stereo_toolbox_b200 as S
    g = load_golden("gwcnet_gc.npz")
    sd, meta = golden_state("gwcnet_gc")
    left, right = _pair(meta)
    net = S.GwcNet_GC(meta["maxdisp"], precision="fp16x2")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.feature_mode = feature_mode
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    epe = (disp - g["disp"]).abs().mean().item()
    cost_err = (net._last_cost.view(-1).cpu() - g["cost3"].reshape(-1)).abs().mean().item() if "cost3" in g else float("nan")
    print(f"GwcNet_GC fp16x2 features={feature_mode}: EPE vs reference {epe:.3e} px, pre-softmax cost err {cost_err:.3e}")
    assert epe < 1e-3


@pytest.mark.parametrize("feature_mode", ["fp32", "umma"])
def test_psmnet_golden_fp16x2(feature_mode):
    """'umma': the SPP extractor's trunk and lastconv on the split tcgen05 kernel too (features_umma.UmmaGwcFeatures.psm), the
    model's default on this precision."""
    import stereo_toolbox_b200 as S
    g = load_golden("psmnet.npz")
    sd, meta = golden_state("psmnet")
    left, right = _pair(meta)
    net = S.PSMNet(meta["maxdisp"], precision="fp16x2")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    if feature_mode == "fp32":
        net.feature_mode = "fp32"
    else:
        assert net.feature_mode is None                   # default -> 'umma' on fp16x2
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    epe = (disp - g["disp"]).abs().mean().item()
    print(f"PSMNet fp16x2 features={feature_mode}: EPE vs reference {epe:.3e} px")
    assert epe < 1e-3
