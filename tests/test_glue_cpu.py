"""glue.inference_shadow: the BatchNorm2d-folded inference copy of the iterative models (host logic, CPU; the hot path is
answered by the oracle like in tests/test_host_mirror_cpu.py).  The folded copy must reproduce the model's output, must
have folded every BatchNorm2d that directly follows a convolution, and must leave the model itself untouched."""
import copy

import pytest
import torch
import torch.nn as nn

from conftest import golden_state, load_golden
from oracle_backend import oracle_hot_path


@pytest.mark.parametrize("name", ["raft_stereo", "igev_stereo"])
def test_folded_shadow_reproduces_the_model(name):
    from stereo_toolbox_b200.synth import synth_pair
    from stereo_toolbox_b200 import glue
    g = load_golden(f"{name}.npz")
    sd, meta = golden_state(name, calib=False) if name == "raft_stereo" else golden_state(name)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.RAFTStereo() if name == "raft_stereo" else S.IGEVStereo({"max_disp": meta["max_disp"]})
        net.load_state_dict(sd, strict=True)
        net.eval()
        before = {k: v.clone() for k, v in net.state_dict().items()}
        n_bn = sum(isinstance(m, nn.BatchNorm2d) for m in net.modules())
        left, right = synth_pair(1, 64, 128, seed=2 if name == "raft_stereo" else 8, shift=meta["shift"])
        ex = torch.zeros(1, 3, 64, 128)
        with torch.no_grad():
            sh = glue.inference_shadow(net, lambda m: m(ex, ex, iters=1))
            assert glue.inference_shadow(net, None) is sh                       # cached while the state is unchanged
            out = sh(left, right, iters=meta["iters"])
    left_bn = sum(isinstance(m, nn.BatchNorm2d) for m in sh.modules())
    print(f"{name}: {sh._folded_bn_count} of {n_bn} BatchNorm2d folded, {left_bn} left")
    assert sh._folded_bn_count > 0 and left_bn == n_bn - sh._folded_bn_count
    assert left_bn <= 2, "every BatchNorm2d of these models directly follows a convolution"
    assert (out - g["disp"]).abs().mean().item() < 1e-3                         # vs the REFERENCE's output
    after = net.state_dict()
    assert before.keys() == after.keys() and all(torch.equal(before[k], after[k]) for k in before)   # the model is untouched
    assert "_fold_shadow" not in dict(net.named_modules()) and not any(k.startswith("_fold") for k in after)
    # a parameter update invalidates the copy
    with torch.no_grad():
        next(net.parameters()).add_(0.0)
    with oracle_hot_path(), torch.no_grad():
        assert glue.inference_shadow(net, lambda m: m(ex, ex, iters=1)) is not sh
