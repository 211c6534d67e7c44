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
    """PSMNet / GwcNet train on the CUDA path only (no CPU fallback: CPU tensors are refused); models whose training path
    is not built raise NotImplementedError instead of silently running something else."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200._lib import StbError
    net = S.GwcNet_GC(32).train()
    with pytest.raises(StbError):
        net(torch.zeros(2, 3, 64, 128), torch.zeros(2, 3, 64, 128))
    with pytest.raises(NotImplementedError):
        S.ACVNet(32).train()(torch.zeros(1, 3, 64, 128), torch.zeros(1, 3, 64, 128))


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
