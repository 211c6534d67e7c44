"""IGEV-Stereo training path on CPU: the drop-in model in train mode with the forward kernels answered by the oracle
(tests/oracle_backend.py: OracleTrainBackend for the 3-D stage, oracle ops for correlation / lookup) and the product's
own adjoints for the geometry-encoding lookup (autograd.py), against one training step of the REFERENCE
(tests/golden/igev_train.npz)."""
import torch

from conftest import load_golden, golden_state
from oracle import ref_ops as R
from oracle_backend import oracle_hot_path


def test_geo_lookup_adjoint_matches_autograd_of_the_oracle():
    from stereo_toolbox_b200.functional import Combined_Geo_Encoding_Volume
    torch.manual_seed(0)
    f1 = torch.randn(2, 12, 3, 20, requires_grad=True)
    f2 = torch.randn(2, 12, 3, 20, requires_grad=True)
    geo = torch.randn(2, 8, 10, 3, 20, requires_grad=True)
    disp = torch.rand(2, 1, 3, 20) * 14 - 2                         # some taps outside [0, D) / [0, W)
    coords = torch.arange(20.0).view(1, 1, 20, 1).repeat(2, 3, 1, 1)
    w = torch.randn(2, 162, 3, 20)
    with oracle_hot_path():
        fn = Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
        assert fn._diff
        (fn(disp, coords) * w).sum().backward()
    got = [t.grad.clone() for t in (f1, f2, geo)]
    for t in (f1, f2, geo):
        t.grad = None
    geos, corrs = R.geo_pyramids(f1, f2, geo, 2)
    (R.geo_lookup(geos, corrs, disp[:, 0], coords[..., 0], 4) * w).sum().backward()
    for a, t in zip(got, (f1, f2, geo)):
        torch.testing.assert_close(a, t.grad, rtol=1e-4, atol=1e-5)


def test_igev_training_step_vs_reference():
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    g = load_golden("igev_train.npz")
    sd, meta = golden_state("igev_stereo")
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.IGEVStereo({"max_disp": meta["max_disp"]})
        net.load_state_dict(sd, strict=True)
        net.train()
        left, right = synth_pair(2, 64, 128, seed=8, shift=5)
        gt = synth_gt(2, 64, 128)[:, None] * 0.25
        init_disp, preds = net(left, right, iters=2)                  # igev_stereo.py:254-255
        assert init_disp.shape == (2, 1, 64, 128) and len(preds) == 2
        loss = (init_disp - gt).abs().mean() + sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean()
                                                   for i, p in enumerate(preds))
        loss.backward()
    assert (init_disp.detach()[:, :, ::2, ::2] - g["init_disp"]).abs().mean().item() < 1e-3
    for i, p in enumerate(preds):
        assert (p.detach()[:, :, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)]
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        assert err < 5e-3, (name, err)
