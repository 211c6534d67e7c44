"""Parity of the CUDA operators (through the C ABI) against the golden fixtures and the oracle."""
import pytest
import torch

from conftest import load_golden
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return t.to(DEV)


def close(a, b, rtol=1e-5, atol=1e-5):
    torch.testing.assert_close(a.cpu(), b, rtol=rtol, atol=atol)


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_volumes_golden():
    import stereo_toolbox_b200 as S
    g = load_golden("ops_volume_head.npz")
    close(S.build_gwc_volume(cu(g["gwc_L"]), cu(g["gwc_R"]), 9, 4), g["gwc_out"])
    close(S.build_gwc_volume(cu(g["gwc8_L"]), cu(g["gwc8_R"]), 6, 8), g["gwc8_out"])
    close(S.build_concat_volume(cu(g["cat_L"]), cu(g["cat_R"]), 9), g["catA_out"], 0, 0)       # bit-exact copy
    close(S.build_concat_volume_unmasked(cu(g["cat_L"]), cu(g["cat_R"]), 9), g["catB_out"], 0, 0)
    from stereo_toolbox_b200 import ops
    prob = ops.softmax_d(cu(g["att"]))
    close(ops.concat_volume(cu(g["cat_L"]), cu(g["cat_R"]), 9, False, att_prob=prob), g["acv_out"])
    close(S.groupwise_correlation(cu(g["gwc_L"]), cu(g["gwc_R"]), 4), g["gwc_out"][:, :, 0])


@pytest.mark.parametrize("B,C,G,H,W,D", [(1, 320, 40, 7, 52, 12), (2, 96, 8, 5, 33, 48), (1, 40, 40, 3, 8, 16),
                                         (1, 64, 4, 2, 4, 9)])
def test_gwc_volume_oracle(B, C, G, H, W, D):
    import stereo_toolbox_b200 as S
    L, Rr = rnd(1, B, C, H, W), rnd(2, B, C, H, W)
    close(S.build_gwc_volume(cu(L), cu(Rr), D, G), R.build_gwc_volume(L, Rr, D, G), 1e-5, 2e-5)


def test_volume_into_slice():
    from stereo_toolbox_b200.aggregation import Fp32Backend
    gl, gr, cl, cr = rnd(1, 2, 80, 6, 20), rnd(2, 2, 80, 6, 20), rnd(3, 2, 12, 6, 20), rnd(4, 2, 12, 6, 20)
    vol = Fp32Backend().volume_gwc_concat(cu(gl), cu(gr), cu(cl), cu(cr), 8, 40)
    want = torch.cat((R.build_gwc_volume(gl, gr, 8, 40), R.build_concat_volume(cl, cr, 8, True)), 1)
    close(vol, want, 1e-5, 2e-5)


def test_head_golden():
    import stereo_toolbox_b200 as S
    g = load_golden("ops_volume_head.npz")
    cost = cu(g["head_cost"])
    close(S.upsample_softargmin(cost, 24, 20, 28, False), g["head_0"], 1e-5, 5e-5)
    close(S.upsample_softargmin(cost, 24, 20, 28, True), g["head_1"], 1e-5, 5e-5)
    close(S.upsample_softargmin(cost, 24, 20, 28, False, keepdim=True), g["head_keepdim"], 1e-5, 5e-5)
    close(S.upsample_softargmin(cu(g["sam_cost"]), 12, 5, 7, keepdim=True), g["sam_out"], 1e-5, 5e-5)
    prob = torch.softmax(g["up_0"], 1)
    close(S.disparity_regression(cu(prob), 24), g["head_0"], 1e-5, 5e-5)
    close(S.disparityregression(24)(cu(prob)), g["head_keepdim"], 1e-5, 5e-5)


@pytest.mark.parametrize("align", [False, True])
def test_head_oracle_x4(align):
    import stereo_toolbox_b200 as S
    cost = rnd(5, 2, 1, 48, 12, 39) * 2
    got = S.upsample_softargmin(cu(cost), 192, 48, 156, align)
    want = R.upsample_softargmin(cost, 192, 48, 156, align)
    assert (got.cpu() - want).abs().max().item() < 1e-3      # values up to 191; fp32 index math vs fp64 oracle


@pytest.mark.parametrize("align,outd", [(False, 48), (True, 48), (False, 40)])
def test_head_extreme_isolated_peak(align, outd):
    """An isolated logit far (> 700) above its neighbours: the interpolated bins never hit the knot itself, so a softmax
    shifted by the KNOT maximum would flush every exp to zero (0/0).  F.softmax -- and the kernels -- shift by the maximum of
    the interpolated bins.  outd=40: the generic kernel (not a x4 upsampling)."""
    import stereo_toolbox_b200 as S
    cost = rnd(6, 1, 1, 12, 5, 9)
    cost[:, :, 7] += 2500.0
    cost[0, 0, 3, 2, 4] = 9000.0
    got = S.upsample_softargmin(cu(cost), outd, 20, 36, align).cpu()
    want = R.upsample_softargmin(cost, outd, 20, 36, align)
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 2e-3


CONVS = [
    # cin, cout, k, stride, pad, transposed, outpad, act, residual
    (8, 16, 3, 1, 1, False, 0, "relu", False),
    (16, 8, 3, 2, 1, False, 0, "relu", False),
    (8, 8, 1, 1, 0, False, 0, "none", False),
    (16, 8, 3, 2, 1, True, 1, "relu", True),
    (12, 40, 3, 1, 1, False, 0, "mish", True),
    (8, 1, 3, 1, 1, False, 0, "none", True),
    (16, 8, 4, 2, 1, True, 0, "leaky", False),
    (64, 32, 3, 1, 1, False, 0, "relu", False),
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,tr,op,act,res", CONVS)
def test_conv3d_family_oracle(cin, cout, k, stride, pad, tr, op, act, res):
    from stereo_toolbox_b200 import ops
    x = rnd(1, 2, cin, 6, 9, 37)
    w = rnd(2, *((cin, cout) if tr else (cout, cin)), k, k, k) * (2.0 / (cin * k ** 3)) ** 0.5
    bn = dict(weight=0.75 + 0.5 * torch.rand(cout), bias=0.1 * rnd(3, cout), running_mean=0.1 * rnd(4, cout),
              running_var=0.5 + torch.rand(cout))
    want0 = R.conv3d_bn_act(x, w, bn, stride, pad, "none", None, tr, op)
    resid = rnd(5, *want0.shape) if res else None
    want = R.conv3d_bn_act(x, w, bn, stride, pad, act, resid, tr, op)
    got = ops.conv3d_bn_act(cu(x), cu(w), tuple(cu(bn[k_]) for k_ in ("weight", "bias", "running_mean", "running_var")),
                            stride, pad, act, None if resid is None else cu(resid), tr, op)
    assert got.shape == want.shape
    close(got, want, 1e-4, 1e-4)


def test_corr_golden():
    import stereo_toolbox_b200 as S
    g = load_golden("ops_corr.npz")
    blk = S.CorrBlock1D(cu(g["f1"]), cu(g["f2"]), num_levels=4, radius=4)
    assert len(blk.corr_pyramid) == 5
    close(S.CorrBlock1D.corr(cu(g["f1"]), cu(g["f2"]))[:, :, :, 0], g["corr"])
    for i, p in enumerate(blk.corr_pyramid):
        assert p.shape == (2 * 3 * 32, 1, 1, 32 >> i)
        close(p.reshape(2, 3, 32, -1), g[f"pyr{i}"])
    out = blk(cu(g["coords"]))
    assert out.shape == g["lookup"].shape and out.dtype == torch.float32
    close(out, g["lookup"], 1e-5, 3e-5)


def test_corr_oracle_raft_shape():
    import stereo_toolbox_b200 as S
    f1, f2 = rnd(1, 1, 256, 6, 100), rnd(2, 1, 256, 6, 100)
    blk = S.CorrBlock1D(cu(f1), cu(f2), 4, 4)
    corr = R.corr1d(f1, f2, True)
    close(blk._levels[0], corr, 1e-4, 1e-4)
    coords = torch.arange(100.0).view(1, 1, 1, 100).repeat(1, 2, 6, 1)
    coords[:, 0] += (torch.rand(1, 6, 100, generator=torch.Generator().manual_seed(3)) - 0.5) * 150
    want = R.corr_lookup(R.corr_pyramid(corr, 4), coords[:, 0], 4, 4)
    close(blk(cu(coords)), want, 1e-4, 2e-4)


def test_geo_golden():
    import stereo_toolbox_b200 as S
    g = load_golden("ops_corr.npz")
    geo = S.Combined_Geo_Encoding_Volume(cu(g["geo_m1"]), cu(g["geo_m2"]), cu(g["geo_vol"]), num_levels=2, radius=4)
    out = geo(cu(g["geo_disp"]), cu(g["geo_coords"]))
    assert out.shape == g["geo_out"].shape
    close(out, g["geo_out"], 1e-5, 3e-5)


def test_torch_ops_registration():
    from stereo_toolbox_b200 import ops
    ops.register_torch_ops()
    L, Rr = rnd(1, 1, 16, 4, 12), rnd(2, 1, 16, 4, 12)
    v = torch.ops.stb200.gwc_volume(cu(L), cu(Rr), 5, 4)
    close(v, R.build_gwc_volume(L, Rr, 5, 4), 1e-5, 2e-5)
    m = torch.ops.stb200.gwc_volume(L.to("meta"), Rr.to("meta"), 5, 4)
    assert m.shape == v.shape


@pytest.mark.parametrize("tag", ["nopad", "padr", "padrb"])
def test_block_attention_golden(tag):
    """stb_block_attention against the reference attention_block outputs (fixture), fp32 NCDHW; the qkv Linear and
    final1x1 go through the conv family as 1x1x1 convolutions with bias."""
    from stereo_toolbox_b200.acvnet import attention_block
    from stereo_toolbox_b200.aggregation import Fp32Backend
    g = load_golden("ops_acv.npz")
    ab = attention_block(32, num_heads=4, block=(4, 4, 4))
    ab.load_state_dict({k: g[f"att_{tag}_{k}"] for k in ("qkv_3d.weight", "qkv_3d.bias", "final1x1.weight", "final1x1.bias")})
    ab = ab.cuda()
    with torch.no_grad():
        y = ab.run(Fp32Backend(), cu(g[f"att_{tag}_x"]))
    close(y, g[f"att_{tag}_y"], 1e-4, 1e-5)


@pytest.mark.parametrize("D,H,W,heads,C", [(12, 24, 78, 16, 128), (4, 5, 3, 2, 32), (8, 9, 13, 8, 64)])
def test_block_attention_oracle(D, H, W, heads, C):
    """KITTI 1/16-scale shape (pad_r = 2, pad_b = 0) and ragged small shapes, kernel alone vs the oracle core."""
    from stereo_toolbox_b200 import ops
    qw, qb = rnd(1, 3 * C, C) * C ** -0.5, rnd(2, 3 * C) * 0.1
    x = rnd(3, 2, C, D, H, W)
    eye = torch.eye(C)
    want = R.block_attention(x, qw, qb, eye.view(C, C, 1, 1, 1), torch.zeros(C), heads)
    qkv = torch.einsum("oc,bcdhw->bodhw", qw, x) + qb.view(1, -1, 1, 1, 1)
    got = ops.block_attention(cu(qkv.contiguous()), cu(qb), heads, (4, 4, 4))
    close(got, want, 1e-4, 1e-5)
    got_cl = ops.block_attention(cu(qkv.permute(0, 2, 3, 4, 1).contiguous()), cu(qb), heads, (4, 4, 4), channels_last=True)
    close(got_cl.permute(0, 4, 1, 2, 3), want, 1e-4, 1e-5)


def test_patch_dw_golden_and_slices():
    from stereo_toolbox_b200 import ops
    g = load_golden("ops_acv.npz")
    x = cu(g["patch_x"])
    for dil in (1, 2, 3):
        close(ops.patch_dw(x, cu(g[f"patch_w{dil}"]), dil), g[f"patch_y{dil}"])
    # channel slices with different dilations into one buffer (ACVNet/acv.py:170-173)
    out = torch.full_like(x, float("nan"))
    ops.patch_dw(x, cu(g["patch_w1"][:2]), 1, out=out, c_off=0)
    ops.patch_dw(x, cu(g["patch_w2"][2:4]), 2, out=out, c_off=2)
    ops.patch_dw(x, cu(g["patch_w3"][4:]), 3, out=out, c_off=4)
    want = torch.cat((g["patch_y1"][:, :2], g["patch_y2"][:, 2:4], g["patch_y3"][:, 4:]), 1)
    close(out, want)


def test_ops_follow_the_tensor_device():
    """model.to('cuda:1') without torch.cuda.set_device (the reference's speed_and_memory_test(model, device='cuda:1')
    pattern): kernels must launch on the tensors' device and stream, not on the current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import stereo_toolbox_b200 as S
    assert torch.cuda.current_device() == 0
    L, Rr = rnd(1, 1, 80, 6, 24), rnd(2, 1, 80, 6, 24)
    vol = S.build_gwc_volume(L.to("cuda:1"), Rr.to("cuda:1"), 8, 40)
    assert vol.device.index == 1 and torch.cuda.current_device() == 0
    close(vol, R.build_gwc_volume(L, Rr, 8, 40), 1e-5, 2e-5)
    with pytest.raises(Exception):
        S.build_gwc_volume(L.to("cuda:0"), Rr.to("cuda:1"), 8, 40)


def test_transposed_conv_with_kernel_smaller_than_stride():
    """ConvTranspose3d(k=1, s=2): seven of the eight output-parity classes receive no tap; they hold act(shift + residual)."""
    from stereo_toolbox_b200 import ops
    x, w = rnd(1, 1, 4, 3, 4, 5), rnd(2, 4, 6, 1, 1, 1)
    res = rnd(3, 1, 6, 5, 7, 9)
    bn = dict(weight=torch.rand(6) + 0.5, bias=rnd(4, 6), running_mean=rnd(5, 6) * 0.1, running_var=torch.rand(6) + 0.5)
    got = ops.conv3d_bn_act(cu(x), cu(w), tuple(cu(bn[k]) for k in ("weight", "bias", "running_mean", "running_var")),
                            stride=2, padding=0, act="relu", residual=cu(res), transposed=True)
    want = R.conv3d_bn_act(x, w, bn, 2, 0, "relu", res, True, 0)
    close(got, want, 1e-5, 1e-5)


def test_convex_upsampling_kernels_golden_and_oracle():
    """RAFTStereo.upsample_flow (x4 and x8) and IGEV context_upsample against the reference's outputs
    (fixture made by tests/golden/make_golden.py from the reference's own functions) and the oracle at a RAFT-sized shape."""
    from stereo_toolbox_b200 import ops
    g = load_golden("ops_upsampling.npz")
    for f in (4, 8):
        got = ops.convex_upsample(cu(g[f"raft_flow{f}"]), cu(g[f"raft_mask{f}"]), f)
        assert got.shape == g[f"raft_up{f}"].shape
        close(got, g[f"raft_up{f}"], 1e-5, 1e-5)
    close(ops.context_upsample(cu(g["igev_disp"]), cu(g["igev_w"])), g["igev_up"], 1e-5, 1e-5)
    # fused softmax + scale (what IGEVStereo.upsample_disp asks for)
    logits = rnd(3, 2, 9, 20, 28) * 2
    want = R.context_upsample(g["igev_disp"] * 4.0, torch.softmax(logits, 1))
    close(ops.context_upsample(cu(g["igev_disp"]), cu(logits), scale=4.0, softmax=True), want, 1e-5, 2e-5)
    flow, mask = rnd(4, 1, 2, 33, 70) * 5, rnd(5, 1, 144, 33, 70) * 3          # odd sizes: tails of the 128-wide blocks
    close(ops.convex_upsample(cu(flow), cu(mask), 4), R.convex_upsample(flow, mask, 4), 1e-5, 2e-5)
    flow, mask = rnd(6, 1, 1, 9, 11), rnd(7, 1, 36, 9, 11)
    close(ops.convex_upsample(cu(flow), cu(mask), 2), R.convex_upsample(flow, mask, 2), 1e-5, 1e-5)


@pytest.mark.parametrize("W", [16, 20, 18])
def test_patch_dw_vectorised_and_scalar_kernels_agree_with_torch(W):
    """W % 4 == 0 takes the four-outputs-per-thread kernel (one aligned 16-byte load left / centre / right of the quad per
    filter row), anything else the scalar one: both against torch's depthwise dilated Conv3d, incl. images narrower than a
    dilated tap reach."""
    from stereo_toolbox_b200 import ops
    g = torch.Generator().manual_seed(W)
    x = torch.randn(2, 6, 3, 7, W, generator=g)
    for dil in (1, 2, 3):
        wt = torch.randn(6, 1, 1, 3, 3, generator=g)
        want = torch.nn.functional.conv3d(x, wt, padding=(0, dil, dil), dilation=(1, dil, dil), groups=6)
        close(ops.patch_dw(cu(x), cu(wt), dil), want, 1e-5, 1e-6)
