"""RAFT-Stereo training path on CPU: the drop-in model in train mode with the forward kernels answered by the oracle
(tests/oracle_backend.py) and the product's own adjoints (autograd.py: _Corr1dFn / _AvgPoolLastFn / _Corr1dLookupFn),
against one training step of the REFERENCE (tests/golden/raft_train.npz)."""
import torch

from conftest import load_golden, golden_state
from oracle import ref_ops as R
from oracle_backend import oracle_hot_path


def test_corrblock_adjoints_match_autograd_of_the_oracle():
    """d(lookup(pyramid(corr(f1, f2))))/d(f1, f2) through the product's autograd Functions vs torch autograd of the
    oracle restatement, including coordinates outside the row (zero taps) and an odd pyramid width."""
    from stereo_toolbox_b200.functional import CorrBlock1D
    torch.manual_seed(0)
    f1 = torch.randn(2, 16, 3, 22, requires_grad=True)
    f2 = torch.randn(2, 16, 3, 22, requires_grad=True)
    coords = torch.rand(2, 2, 3, 22) * 30 - 4                       # some taps fall outside [0, W2)
    w = torch.randn(2, 4 * 9, 3, 22)
    with oracle_hot_path():
        blk = CorrBlock1D(f1, f2, num_levels=4, radius=4)
        assert blk._diff
        (blk(coords) * w).sum().backward()
    g1, g2 = f1.grad.clone(), f2.grad.clone()
    f1.grad = f2.grad = None
    want = R.corr_lookup(R.corr_pyramid(R.corr1d(f1, f2, True), 4), coords[:, 0], 4, 4)
    (want * w).sum().backward()
    torch.testing.assert_close(g1, f1.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g2, f2.grad, rtol=1e-4, atol=1e-5)


def test_raft_training_step_vs_reference():
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    g = load_golden("raft_train.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.RAFTStereo()
        net.load_state_dict(sd, strict=True)
        net.train()
        net.freeze_bn()
        left, right = synth_pair(1, 64, 128, seed=2, shift=3)
        gt = synth_gt(1, 64, 128)[:, None] * 0.25
        preds = net(left, right, iters=3)
        assert isinstance(preds, list) and len(preds) == 3 and preds[0].shape == (1, 1, 64, 128)     # raft_stereo.py:188
        loss = sum(0.9 ** (len(preds) - i - 1) * (p - gt).abs().mean() for i, p in enumerate(preds))
        loss.backward()
    for i, p in enumerate(preds):
        assert (p.detach()[:, :, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)]              # the fixture's sampling rule (make_golden.py raft_train)
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        assert err < 2e-3, (name, err)
