"""Size-independent properties of the hot-path kernels at BASELINE.json's FULL shapes (KITTI 1/4-resolution volume
48x96x312 from 384x1248; RAFT 1/4-resolution 128x256 from 512x1024), where the CPU oracle is too slow to be the checker:
exact copies / exact zeros, linearity, power-of-two scaling (commutes with 16-bit rounding, hence bit-exact),
shift invariance of the soft-argmin, symmetry of the all-pairs correlation, pyramid = pairwise averages, integer-coordinate
lookups = pyramid entries.  The kernels are the confirmed ones (bench.py runs them at these sizes); the tests themselves were
written after the round-1 GPU budget was spent."""
import pytest
import torch


pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

B, H4, W4, D4 = 2, 96, 312, 48


def _rand(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).cuda()


def test_gwc_volume_zero_region_linearity_and_checksum():
    import stereo_toolbox_b200 as S
    L1, L2, R = _rand(1, B, 320, H4, W4), _rand(2, B, 320, H4, W4), _rand(3, B, 320, H4, W4)
    v1, v2 = S.build_gwc_volume(L1, R, D4, 40), S.build_gwc_volume(L2, R, D4, 40)
    assert v1.shape == (B, 40, D4, H4, W4)
    for d in (1, 7, 47):
        assert not v1[:, :, d, :, :d].any()                                   # w < d: exact zeros (GwcNet/submodule.py:57-61)
    v12 = S.build_gwc_volume(2.0 * L1 - 0.5 * L2, R, D4, 40)
    torch.testing.assert_close(v12, 2.0 * v1 - 0.5 * v2, rtol=1e-4, atol=1e-4)   # linear in the left features
    torch.testing.assert_close(v1[:, :, 0].mean(1), (L1 * R).mean(1), rtol=1e-4, atol=1e-5)   # d = 0: mean of group means
    d = 5                                                                      # plane d = plane 0 of the shifted pair
    torch.testing.assert_close(v1[:, :, d, :, d:], S.build_gwc_volume(L1[..., d:].contiguous(), R[..., :-d].contiguous(), 1, 40)[:, :, 0],
                               rtol=1e-5, atol=1e-6)


def test_concat_volume_is_an_exact_copy():
    import stereo_toolbox_b200 as S
    L, R = _rand(4, B, 12, H4, W4), _rand(5, B, 12, H4, W4)
    for builder, masked in ((S.build_concat_volume, True), (S.build_concat_volume_unmasked, False)):
        v = builder(L, R, D4)
        assert v.shape == (B, 24, D4, H4, W4)
        for d in (0, 3, 47):
            assert torch.equal(v[:, :12, d, :, d:], L[..., d:])
            assert torch.equal(v[:, 12:, d, :, d:], R[..., :W4 - d])
            assert not v[:, 12:, d, :, :d].any()
            if d:
                assert (not v[:, :12, d, :, :d].any()) if masked else torch.equal(v[:, :12, d, :, :d], L[..., :d])


def test_softargmin_head_invariances():
    import stereo_toolbox_b200 as S
    cost = _rand(6, B, 1, D4, H4, W4) * 2
    d0 = S.upsample_softargmin(cost, 192, 384, 1248)
    assert d0.shape == (B, 384, 1248) and d0.min() >= 0 and d0.max() <= 191
    torch.testing.assert_close(S.upsample_softargmin(cost + 3.75, 192, 384, 1248), d0, rtol=1e-4, atol=2e-3)   # softmax shift invariance
    flat = S.upsample_softargmin(torch.zeros_like(cost), 192, 384, 1248)
    torch.testing.assert_close(flat, torch.full_like(flat, 95.5), rtol=0, atol=1e-3)                            # uniform -> mean of 0..191
    peak = torch.full_like(cost, -30.0)
    peak[:, :, 20] = 30.0                                                     # all mass on source plane 20 -> bins 80..83
    torch.testing.assert_close(S.upsample_softargmin(peak, 192, 384, 1248), torch.full_like(flat, 81.5), rtol=0, atol=0.51)


def test_corr_symmetry_pyramid_and_integer_lookup():
    import stereo_toolbox_b200 as S
    f1, f2 = _rand(7, 1, 256, 128, 256), _rand(8, 1, 256, 128, 256)
    blk, blk_t = S.CorrBlock1D(f1, f2, num_levels=4, radius=4), S.CorrBlock1D(f2, f1, num_levels=4, radius=4)
    c, ct = blk._levels[0], blk_t._levels[0]
    assert c.shape == (1, 128, 256, 256)
    torch.testing.assert_close(c, ct.transpose(2, 3), rtol=1e-4, atol=1e-4)                   # corr(f1,f2)[w1,w2] = corr(f2,f1)[w2,w1]
    torch.testing.assert_close(c[0, 5, 17, 33], (f1[0, :, 5, 17] * f2[0, :, 5, 33]).sum() / 16.0, rtol=1e-4, atol=1e-4)
    for lo, hi in zip(blk._levels[:-1], blk._levels[1:]):
        torch.testing.assert_close(hi, 0.5 * (lo[..., 0::2] + lo[..., 1::2]), rtol=1e-6, atol=1e-6)
    xs = torch.arange(256.0, device="cuda").view(1, 1, 1, 256).repeat(1, 2, 128, 1)           # identity coordinates
    out = blk(xs)
    assert out.shape == (1, 36, 128, 256)
    w = torch.arange(256, device="cuda")
    for k in range(9):                                                        # level 0, tap k: corr[h, w, w + k - 4], zero outside
        idx = w + k - 4
        ok = (idx >= 0) & (idx < 256)
        want = torch.where(ok, c[0, :, w, idx.clamp(0, 255)], torch.zeros((), device="cuda"))
        torch.testing.assert_close(out[0, k], want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_tcgen05_conv_power_of_two_scaling_is_bit_exact(prec):
    """conv(2x) == 2 conv(x) bit for bit on the 16-bit tensor-core path (no BatchNorm shift, no activation): scaling by a
    power of two commutes with every rounding on the way (operands, fp32 accumulation, 16-bit store)."""
    import torch.nn as nn
    from stereo_toolbox_b200.aggregation_umma import UmmaBackend
    be = UmmaBackend(prec)
    torch.manual_seed(0)
    for conv, dims in ((nn.Conv3d(32, 32, 3, 1, 1, bias=False), (D4, H4, W4)), (nn.Conv3d(32, 64, 3, 2, 1, bias=False), (D4, H4, W4)),
                       (nn.ConvTranspose3d(64, 32, 3, 2, 1, output_padding=1, bias=False), (D4 // 2, H4 // 2, W4 // 2))):
        conv = conv.cuda()
        cin = conv.in_channels
        x = (torch.randn(1, *dims, cin, device="cuda") * 0.5).to(be.dtype)
        y1, y2 = be.conv(conv, x, "none"), be.conv(conv, (x.float() * 2).to(be.dtype), "none")
        normal = y1.float().abs() > 1e-3          # fp16 subnormal results round at a fixed absolute step: scaling is exact only above them
        assert normal.float().mean().item() > 0.9
        assert torch.equal(y2.float()[normal], y1.float()[normal] * 2)
