"""PCWNet_GC / CFNet training paths on CPU: the drop-in models in train mode with the 3-D path on the oracle's
TrainBackend stand-in (tests/oracle_backend.py), against one training step of the REFERENCE (tests/golden/*_train.npz:
the full prediction lists, loss, and one weight gradient per sub-network)."""
import torch.nn.functional as F

from conftest import load_golden, golden_state
from oracle_backend import oracle_hot_path


def _step(key, ctor, seed, fixture):
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    g = load_golden(fixture)
    sd, meta = golden_state(key)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = ctor(S, meta)
        net.load_state_dict(sd, strict=True)
        net.train()
        left, right = synth_pair(2, 64, 128, seed=seed, shift=5)
        gt = synth_gt(2, 64, 128)
        preds = net(left, right)
        mask = (gt > 0) & (gt < meta["maxdisp"])
        loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
        loss.backward()
    return g, net, preds, loss


def _check(g, net, preds, loss, n_preds):
    assert isinstance(preds, list) and len(preds) == n_preds
    for i, p in enumerate(preds):
        assert p.shape == (2, 64, 128)
        assert (p.detach()[:, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3, i
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)]
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        cos = F.cosine_similarity(got, want, dim=0).item()
        # (a conv weight directly in front of a train-mode BatchNorm has a near-zero, cancellation-dominated gradient:
        #  dispupsample.0.0 differs by 5e-3 of its tiny maximum at cosine 0.999999; everything else agrees to ~1e-6)
        assert err < 5e-3 or cos > 0.99999, (name, err, cos)


def test_pcwnet_gc_training_step_vs_reference():
    g, net, preds, loss = _step("pcwnet_gc", lambda S, m: S.PCWNet_GC(m["maxdisp"]), 7, "pcwnet_train.npz")
    _check(g, net, preds, loss, 6)                      # pcwnet.py:480


def test_cfnet_training_step_vs_reference():
    g, net, preds, loss = _step("cfnet", lambda S, m: S.CFNet(m["maxdisp"]), 6, "cfnet_train.npz")
    _check(g, net, preds, loss, 9)                      # cfnet.py:651


def _acv_step(**ctor_kw):
    from stereo_toolbox_b200.synth import synth_pair, synth_gt
    sd, meta = golden_state("acvnet")
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.ACVNet(meta["maxdisp"], **ctor_kw)
        net.load_state_dict(sd, strict=True)
        net.train()
        left, right = synth_pair(2, 64, 144, seed=3, shift=5)
        gt = synth_gt(2, 64, 144)
        preds = net(left, right)
        mask = (gt > 0) & (gt < meta["maxdisp"])
        loss = sum(F.smooth_l1_loss(p[mask], gt[mask], reduction="mean") for p in preds)
        loss.backward()
    return net, preds, loss


def test_acvnet_training_step_vs_reference():
    """Four predictions (acv.py:235); gradients through the patch convs, the windowed attention (one-sided padding case)
    and the softmax-weighted concat volume."""
    g = load_golden("acvnet_train.npz")
    net, preds, loss = _acv_step()
    assert isinstance(preds, list) and len(preds) == 4
    for i, p in enumerate(preds):
        assert p.shape == (2, 64, 144)
        assert (p.detach()[:, ::2, ::2] - g[f"pred{i}"]).abs().mean().item() < 1e-3, i
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * abs(g["loss"].item())
    params = dict(net.named_parameters())
    for name in [k[5:] for k in g if k.startswith("grad:")]:
        got, want = params[name].grad.flatten(), g["grad:" + name]
        got = got[::max(1, got.numel() // 20000)] if want.numel() != got.numel() else got
        err = (got - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        cos = F.cosine_similarity(got, want, dim=0).item()
        assert err < 5e-3 or cos > 0.99999, (name, err, cos)


def test_acvnet_training_variants():
    """freeze_attn_weights: three predictions, no gradient into the attention branch (acv.py:164-178, 233);
    attn_weights_only: one prediction (acv.py:235)."""
    net, preds, _ = _acv_step(freeze_attn_weights=True)
    assert len(preds) == 3
    assert all(p.grad is None for n, p in net.named_parameters() if n.startswith(("patch", "dres1_att_", "dres2_att_", "classif_att_")))
    assert net.dres0[0][0].weight.grad is not None and net.concatconv[0][0].weight.grad is not None
    net, preds, _ = _acv_step(attn_weights_only=True)
    assert len(preds) == 1 and net.patch.weight.grad is not None and net.dres0[0][0].weight.grad is None
