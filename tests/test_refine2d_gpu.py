"""PCWNet's refinement inputs as kernels (csrc/refine2d.cu: warp at x - disp with the reference's grid arithmetic, the
+-maxdisp 1-D correlation volume incl. the reference's negative-offset slices) against the torch forms of the same
functions in stereo_toolbox_b200/pcwnet.py -- which the CPU suite pins to the reference through the whole-model fixture
(tests/test_host_mirror_cpu.py::test_pcwnet_gc_mirror)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _torch_form(fn, *args):
    """the differentiable torch form: taken whenever an argument requires grad"""
    with torch.enable_grad():
        return fn(*[a.clone().requires_grad_(True) for a in args[:1]], *args[1:]).detach()


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 37, 150), (1, 32, 64, 128), (1, 8, 5, 300)])
def test_warp_kernel_matches_grid_sample(B, C, H, W):
    from stereo_toolbox_b200 import pcwnet as P
    x = rnd(1, B, C, H, W).cuda()
    disp = (torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(2)) * 40.0 - 4.0).cuda()   # incl. out-of-image targets
    with torch.no_grad():
        got = P.warp(x, disp)
    want = _torch_form(P.warp, x, disp)
    # the sampling position is O(W) in fp32 (ulp 1.5e-5 at W = 150): a last-bit difference there moves a value by ~1e-5, and a
    # pixel whose mask weight sits within rounding of the 0.999 threshold may flip -- count those, bound everything
    err = (got - want).abs()
    bad = err > 1e-5
    print(f"warp: max err {err.max().item():.3e}, fraction above 1e-5: {bad.float().mean().item():.2e}")
    assert bad.float().mean().item() < 1e-3, bad.float().mean().item()
    assert err.max().item() < 1e-3


@pytest.mark.parametrize("B,C,H,W,md", [(2, 32, 9, 150, 24), (1, 32, 3, 128, 24), (1, 16, 4, 61, 8), (1, 40, 2, 300, 4)])
def test_corr_volume_kernel_matches_reference_slices(B, C, H, W, md):
    from stereo_toolbox_b200 import pcwnet as P
    left, right = rnd(3, B, C, H, W).cuda(), rnd(4, B, C, H, W).cuda()
    with torch.no_grad():
        got = P.build_correlation_volume(left, right, md)
    want = _torch_form(lambda l, r, m: P.build_correlation_volume(l, r, m), left, right, md)
    assert got.shape == want.shape == (B, 2 * md + 1, H, W)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-6)
