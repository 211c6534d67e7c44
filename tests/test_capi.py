"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/stb200.h declares, rejects bad arguments without touching a GPU, and the host mirror keeps
the reference's state-dict layout."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_meta
from stereo_toolbox_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "stb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    h = _lib.lib()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/stb200.h but not exported by libstb200.so"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"
    assert h.stb_version() >= 100
    assert b"bad argument" in h.stb_error_string(-1)


def test_bad_arguments_are_rejected_before_any_launch():
    h = _lib.lib()
    null = ctypes.c_void_p(0)
    assert h.stb_gwc_volume_f32(null, null, null, 1, 8, 4, 4, 2, 4, 4, 0, null) == -1
    one = ctypes.c_void_p(16)
    assert h.stb_gwc_volume_f32(one, one, one, 1, 10, 4, 4, 2, 4, 4, 0, null) == -1      # C % G != 0
    assert h.stb_upsample_softargmin_f32(one, one, 1, 0, 4, 4, 8, 16, 16, 0, null) == -1
    with pytest.raises(_lib.StbError):
        _lib.check(-2, "x")


def test_ops_refuse_cpu_tensors():
    from stereo_toolbox_b200 import build_gwc_volume, CorrBlock1D
    with pytest.raises(_lib.StbError):
        build_gwc_volume(torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4), 2, 4)
    with pytest.raises(_lib.StbError):
        CorrBlock1D(torch.zeros(1, 8, 4, 16), torch.zeros(1, 8, 4, 16))


def test_state_dict_layout_matches_reference():
    import stereo_toolbox_b200 as S
    meta = load_meta("models.json")
    for key, ctor in (("gwcnet_gc", lambda: S.GwcNet_GC(32)), ("gwcnet_g", lambda: S.GwcNet_G(32)),
                      ("psmnet", lambda: S.PSMNet(32)), ("raft_stereo", lambda: S.RAFTStereo()),
                      ("acvnet", lambda: S.ACVNet(64)), ("cfnet", lambda: S.CFNet(64)),
                      ("pcwnet_gc", lambda: S.PCWNet_GC(64)), ("igev_stereo", lambda: S.IGEVStereo({"max_disp": 64}))):
        sd = ctor().state_dict()
        want = meta[key]["keys"]
        assert set(sd) == set(want)
        for k, shape in want.items():
            assert list(sd[k].shape) == shape, k


def test_conv_plan_tap_lists():
    """Tap decomposition of the conv flavours (host logic, no GPU): tap counts and offsets."""
    from stereo_toolbox_b200.ops import ConvPlan
    w = torch.randn(4, 3, 3, 3, 3)
    p = ConvPlan(w, None, 2, 1, False)
    assert len(p.classes) == 1 and p.classes[0][4] == 27 and p.classes[0][5] == 2 and p.out_size(8) == 4
    wt = torch.randn(3, 4, 3, 3, 3)
    p = ConvPlan(wt, None, 2, 1, True, 1)           # ConvTranspose3d k3 s2 p1 op1
    assert sorted(c[4] for c in p.classes) == [1, 2, 2, 2, 4, 4, 4, 8] and p.out_size(6) == 12
    assert sum(c[4] for c in p.classes) == 27
    p = ConvPlan(torch.randn(3, 4, 4, 4, 4), None, 2, 1, True, 0)   # IGEV k4 s2 p1
    assert [c[4] for c in p.classes] == [8] * 8 and p.out_size(6) == 12


def test_training_mode_fails_loudly():
    """Every model trains on the CUDA path only: there is no CPU fallback in train mode either (CPU tensors are refused
    by the first hot-path op), and nothing silently runs something else."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200._lib import StbError
    for net, hw in ((S.GwcNet_GC(32), (64, 128)), (S.PSMNet(32), (256, 256)), (S.ACVNet(32), (64, 128)), (S.CFNet(64), (64, 128)),
                    (S.PCWNet_GC(64), (64, 128))):
        net.train()
        with pytest.raises(StbError):
            net(torch.zeros(2, 3, *hw), torch.zeros(2, 3, *hw))
    with pytest.raises(NotImplementedError):           # the reference's own PCWNet_G cannot run either (pcwnet.py:127-131 vs :493)
        S.PCWNet_G(64).train()(torch.zeros(2, 3, 64, 128), torch.zeros(2, 3, 64, 128))


def test_load_checkpoint_flexible(tmp_path):
    """models/__init__.py:20-51: DDP-prefixed ('module.') checkpoints under a state-dict key, plain checkpoints, extra
    keys ignored, keys absent from the checkpoint keep the model's current values."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_state_dict
    src = S.GwcNet_G(32)
    sd = synth_state_dict(src.state_dict(), 3)
    held_out = "classif3.2.weight"
    ddp = {"module." + k: v for k, v in sd.items() if k != held_out}
    ddp["module.not_in_model.weight"] = torch.zeros(3)
    path = tmp_path / "ckpt.pth"
    torch.save({"model": ddp, "epoch": 7}, path)
    net = S.GwcNet_G(32)
    before = net.state_dict()[held_out].clone()
    out = S.load_checkpoint_flexible(net, str(path), "model")
    assert out is net
    got = net.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], before if k == held_out else v), k
    plain = tmp_path / "plain.pth"
    torch.save(sd, plain)
    net2 = S.load_checkpoint_flexible(S.GwcNet_G(32), str(plain))
    assert all(torch.equal(net2.state_dict()[k], v) for k, v in sd.items())


def test_speed_and_memory_test_keeps_the_reference_signature():
    """evaluation/speed_and_memory_test.py:11: (model, resolution=None, batch_size=1, num_iterations=100, device='cuda:0')."""
    import inspect
    import stereo_toolbox_b200 as S
    params = inspect.signature(S.speed_and_memory_test).parameters
    assert list(params)[:5] == ["model", "resolution", "batch_size", "num_iterations", "device"]
    assert params["resolution"].default is None and params["batch_size"].default == 1
    assert params["num_iterations"].default == 100 and params["device"].default == "cuda:0"
    with pytest.raises(ValueError):
        S.speed_and_memory_test(torch.nn.Identity(), device="cpu")


def test_torch_ops_fake_kernels_give_reference_shapes():
    """torch.ops.stb200.*: every op has a fake (Meta) kernel, so tracing / shape propagation around a patched model works
    without a device; shapes follow the reference functions."""
    from stereo_toolbox_b200 import ops
    ops.register_torch_ops()
    T = torch.ops.stb200
    m = lambda *s: torch.empty(*s, device="meta")
    assert T.gwc_volume(m(2, 320, 6, 20), m(2, 320, 6, 20), 12, 40).shape == (2, 40, 12, 6, 20)
    assert T.concat_volume(m(2, 12, 6, 20), m(2, 12, 6, 20), 12, True).shape == (2, 24, 12, 6, 20)
    assert T.upsample_softargmin(m(2, 1, 12, 6, 20), 48, 24, 80, False).shape == (2, 24, 80)
    assert T.conv3d_bn_act(m(1, 32, 8, 8, 8), m(64, 32, 3, 3, 3), None, None, None, None, 2, 1, "relu", None, False, 0).shape \
        == (1, 64, 4, 4, 4)
    assert T.conv3d_bn_act(m(1, 64, 4, 4, 4), m(64, 32, 3, 3, 3), None, None, None, None, 2, 1, "none", None, True, 1).shape \
        == (1, 32, 8, 8, 8)
    assert T.corr1d(m(1, 256, 8, 16), m(1, 256, 8, 16), True).shape == (1, 8, 16, 16)
    assert T.corr1d_lookup([m(1, 8, 16, 16), m(1, 8, 16, 8)], m(1, 2, 8, 16), 4, 2).shape == (1, 18, 8, 16)
    assert T.avgpool_last(m(1, 8, 16, 16)).shape == (1, 8, 16, 8)
    geos, corrs = [m(1, 8, 16, 8, 12), m(1, 8, 16, 8, 6)], [m(1, 8, 16, 16), m(1, 8, 16, 8)]
    assert T.geo_lookup(geos, corrs, m(1, 1, 8, 16), m(1, 8, 16, 1), 4).shape == (1, 162, 8, 16)     # IGEVStereo/update.py:76
    assert T.feature_gate(m(1, 8, 12, 8, 16), m(1, 8, 8, 16)).shape == (1, 8, 12, 8, 16)
    assert T.softmax_d(m(1, 1, 12, 8, 16)).shape == (1, 1, 12, 8, 16)
    assert T.disparity_regression(m(2, 12, 8, 16), 12, False).shape == (2, 8, 16)
    assert T.disparity_regression(m(2, 12, 8, 16), 12, True).shape == (2, 1, 8, 16)
    assert T.disparity_variance(m(2, 12, 8, 16), 12, m(2, 1, 8, 16)).shape == (2, 1, 8, 16)
    assert T.patch_dw(m(1, 40, 12, 8, 16), m(40, 1, 1, 3, 3), 2).shape == (1, 40, 12, 8, 16)
    assert T.block_attention(m(1, 96, 4, 8, 8), m(96), 4, [4, 4, 4]).shape == (1, 32, 4, 8, 8)


def test_call_device_tracking():
    """ops._p notes the device of every tensor argument; a call may not mix devices, and _lib.call() clears the note
    (the launch itself then runs under torch.cuda.device(<that device>) -- exercised on a 2-GPU box by
    tests/test_gpu_ops.py::test_ops_follow_the_tensor_device)."""
    from stereo_toolbox_b200 import _lib
    _lib._reset_device()
    _lib.note_device(1)
    _lib.note_device(1)
    assert _lib.call_device() == 1
    with pytest.raises(_lib.StbError):
        _lib.note_device(0)
    assert _lib.call_device() is None                       # a refused call leaves no stale note behind
