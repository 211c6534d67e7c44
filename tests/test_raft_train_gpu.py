"""RAFT-Stereo training path on the GPU: forward kernels of csrc/corr1d.cu + the composed adjoints of autograd.py, vs
torch autograd of the oracle and vs one training step of the reference (tests/golden/raft_train.npz)."""
import pytest
import torch

from conftest import load_golden, golden_state
from oracle import ref_ops as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_corrblock_adjoints_match_autograd_of_the_oracle():
    from stereo_toolbox_b200.functional import CorrBlock1D
    torch.manual_seed(0)
    f1c, f2c = torch.randn(2, 16, 3, 22), torch.randn(2, 16, 3, 22)
    coords = torch.rand(2, 2, 3, 22) * 30 - 4
    w = torch.randn(2, 4 * 9, 3, 22)
    f1, f2 = f1c.cuda().requires_grad_(True), f2c.cuda().requires_grad_(True)
    blk = CorrBlock1D(f1, f2, num_levels=4, radius=4)
    assert blk._diff
    (blk(coords.cuda()) * w.cuda()).sum().backward()
    f1c.requires_grad_(True)
    f2c.requires_grad_(True)
    want = R.corr_lookup(R.corr_pyramid(R.corr1d(f1c, f2c, True), 4), coords[:, 0], 4, 4)
    (want * w).sum().backward()
    torch.testing.assert_close(f1.grad.cpu(), f1c.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(f2.grad.cpu(), f2c.grad, rtol=1e-3, atol=1e-4)


def test_raft_training_step_vs_reference():
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    g = load_golden("raft_train.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    net = S.RAFTStereo()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    net.freeze_bn()
    left, right = synth_pair(1, 64, 128, seed=2, shift=3)
    gt = (synth_gt(1, 64, 128)[:, None] * 0.25).cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        preds = net(left.cuda(), right.cuda(), iters=3)
        loss = sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean() for i, p in enumerate(preds))
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert len(preds) == 3
    for i, p in enumerate(preds):
        assert (p.detach().cpu()[:, :, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3
    assert abs(loss.item() - g["loss"].item()) < 1e-3 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten().cpu(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)]
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        assert err < 2e-2, (name, err)
