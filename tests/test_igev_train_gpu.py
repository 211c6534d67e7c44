"""IGEV-Stereo training path on the GPU: 3-D stage on aggregation.TrainBackend (forward + backward kernels of
libstb200.so, incl. the k4-s2 transposed convs), geometry-encoding lookup forward kernel + composed adjoint, vs one
training step of the reference (tests/golden/igev_train.npz)."""
import pytest
import torch

from conftest import load_golden, golden_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_igev_training_step_vs_reference():
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    g = load_golden("igev_train.npz")
    sd, meta = golden_state("igev_stereo")
    net = S.IGEVStereo({"max_disp": meta["max_disp"]})
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    left, right = synth_pair(2, 64, 128, seed=8, shift=5)
    gt = (synth_gt(2, 64, 128)[:, None] * 0.25).cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        init_disp, preds = net(left.cuda(), right.cuda(), iters=2)
        loss = (init_disp - gt).abs().mean() + sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean()
                                                   for i, p in enumerate(preds))
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert (init_disp.detach().cpu()[:, :, ::2, ::2] - g["init_disp"]).abs().mean().item() < 1e-3
    for i, p in enumerate(preds):
        assert (p.detach().cpu()[:, :, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3
    assert abs(loss.item() - g["loss"].item()) < 1e-3 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten().cpu(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)]
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        assert err < 2e-2, (name, err)
