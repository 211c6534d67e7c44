"""Host-side mirrors of the reference models on CPU: every drop-in model runs with its hot path answered by the oracle
(tests/oracle_backend.py) and is compared with the output the REFERENCE itself produced for the same seeded input and
name-keyed weights (tests/golden/*.npz, tests/golden/make_golden.py).  What this pins without a GPU: constructor
signatures, state-dict layout (strict load of the reference's own keys), the torch glue around the hot path (2-D
extractors, CFNet's uniform sampler + warps, PCWNet's refinement, ACVNet's attention plumbing), layer order, residual
wiring and return conventions.  The CUDA kernels behind the same calls are pinned by the ``-m gpu`` tests."""
import pytest
import torch

from conftest import load_golden, golden_state
from oracle_backend import oracle_hot_path


def _pair(meta, seed):
    from stereo_toolbox_b200.synth import synth_pair
    b, h, w = meta["shape"]
    return synth_pair(b, h, w, seed=seed, shift=meta["shift"])


def _run(ctor, key, seed, **fwd):
    sd, meta = golden_state(key)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = ctor(S, meta)
        net.load_state_dict(sd, strict=True)
        net.eval()
        left, right = _pair(meta, seed)
        with torch.no_grad():
            return net(left, right, **fwd), net


@pytest.mark.parametrize("key", ["gwcnet_gc", "gwcnet_g"])
def test_gwcnet_mirror(key):
    g = load_golden(f"{key}.npz")
    disp, net = _run(lambda S, m: (S.GwcNet_GC if key == "gwcnet_gc" else S.GwcNet_G)(m["maxdisp"]), key, 0)
    assert disp.shape == g["disp"].shape                       # [B,H,W]  (gwcnet.py:224)
    if "cost3" in g:
        torch.testing.assert_close(net._last_cost, g["cost3"], rtol=1e-3, atol=1e-3)
    assert (disp - g["disp"]).abs().mean().item() < 1e-3


def test_psmnet_mirror():
    g = load_golden("psmnet.npz")
    disp, _ = _run(lambda S, m: S.PSMNet(m["maxdisp"]), "psmnet", 1)
    assert disp.shape == g["disp"].shape == (1, 1, 256, 256)   # keepdim regression (PSMNet/submodule.py:53)
    assert (disp - g["disp"]).abs().mean().item() < 1e-3


def test_acvnet_mirror():
    g = load_golden("acvnet.npz")
    disp, _ = _run(lambda S, m: S.ACVNet(m["maxdisp"]), "acvnet", 3)
    assert disp.shape == g["disp"].shape
    assert (disp - g["disp"]).abs().mean().item() < 1e-3


def test_cfnet_mirror():
    """Includes the torch glue the GPU build keeps outside the kernels: uniform sampler, gather warps, sampled volumes."""
    g = load_golden("cfnet.npz")
    disp, _ = _run(lambda S, m: S.CFNet(m["maxdisp"]), "cfnet", 6)
    assert disp.shape == g["disp"].shape
    assert (disp - g["disp"]).abs().mean().item() < 1e-3


def test_pcwnet_gc_mirror():
    g = load_golden("pcwnet_gc.npz")
    disp, _ = _run(lambda S, m: S.PCWNet_GC(m["maxdisp"]), "pcwnet_gc", 7)
    assert disp.shape == g["disp"].shape
    assert (disp - g["disp"]).abs().mean().item() < 1e-3


def test_raft_stereo_mirror():
    """RAFTStereo through the package's own ``functional.CorrBlock1D`` host code (pyramid bookkeeping, level count)."""
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("raft_stereo.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.RAFTStereo()
        net.load_state_dict(sd, strict=True)
        net.eval()
        left, right = synth_pair(1, 64, 128, seed=2, shift=meta["shift"])
        with torch.no_grad():
            out = net(left, right, iters=meta["iters"])
    assert out.shape == g["disp"].shape == (1, 1, 64, 128)
    assert (out - g["disp"]).abs().mean().item() < 1e-3


def test_igev_stereo_mirror():
    """IGEVStereo through ``igev.IGEVCostVolume.stage`` (hourglass wiring, feature gates, layout calls) and
    ``functional.Combined_Geo_Encoding_Volume`` -- the host code the GPU build runs, unlike the coarser stage swap of
    tests/test_oracle_golden.py::test_igev_stereo_mirror_with_oracle_stage."""
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("igev_stereo.npz")
    sd, meta = golden_state("igev_stereo")
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.IGEVStereo({"max_disp": meta["max_disp"]})
        net.load_state_dict(sd, strict=True)
        net.eval()
        left, right = synth_pair(1, 64, 128, seed=8, shift=meta["shift"])
        with torch.no_grad():
            out = net(left, right, iters=meta["iters"])
    assert out.shape == g["disp"].shape == (1, 1, 64, 128)
    assert (out - g["disp"]).abs().mean().item() < 1e-3


def _variant_tags():
    from conftest import load_meta
    return sorted(load_meta("variants.json"))


@pytest.mark.parametrize("tag", _variant_tags())
def test_iterative_model_argument_variants(tag):
    """Non-default constructor arguments of RAFTStereo (Namespace) / IGEVStereo (dict): GRU level count, slow-fast
    schedule, 1/8 resolution, correlation levels / radius, context norm, max_disp -- state-dict layout (strict load of
    the reference's keys) and output vs the reference run with the same arguments."""
    import argparse
    from conftest import load_meta
    from stereo_toolbox_b200.synth import synth_pair, synth_state_dict, state_checksum
    meta = load_meta("variants.json")[tag]
    want = load_golden("variants.npz")[tag]
    template = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
                for k, s in meta["keys"].items()}
    sd = synth_state_dict(template, 0)
    assert abs(state_checksum(sd) - meta["checksum"]) <= 1e-6 * abs(meta["checksum"])
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        if meta["family"] == "raft":
            net = S.RAFTStereo(argparse.Namespace(**meta["args"]))
        else:
            net = S.IGEVStereo(dict(meta["args"]))
        net.load_state_dict(sd, strict=True)
        net.eval()
        left, right = synth_pair(1, 64, 128, seed=2, shift=3)
        with torch.no_grad():
            out = net(left, right, iters=3)
    assert out.shape == want.shape == (1, 1, 64, 128)
    assert (out - want).abs().mean().item() < 1e-3


def test_swap_is_scoped():
    """Outside ``oracle_hot_path`` the product refuses CPU tensors again (no CPU fallback is left behind)."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200._lib import StbError
    with oracle_hot_path():
        pass
    with pytest.raises(StbError):
        S.build_gwc_volume(torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4), 2, 4)
    with pytest.raises(StbError):
        S.GwcNet_G(32).eval()(torch.zeros(1, 3, 64, 128), torch.zeros(1, 3, 64, 128))


@pytest.mark.parametrize("family", ["raft", "igev"])
def test_channels_last_glue_is_numerically_equivalent(family):
    """``model.channels_last = True`` re-lays the torch glue (2-D convs) as NHWC: same arithmetic, same result (to summation
    order) as the default NCHW run, against the reference's output."""
    from stereo_toolbox_b200.synth import synth_pair
    if family == "raft":
        g, (sd, meta), seed = load_golden("raft_stereo.npz"), golden_state("raft_stereo", calib=False), 2
    else:
        g, (sd, meta), seed = load_golden("igev_stereo.npz"), golden_state("igev_stereo"), 8
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = S.RAFTStereo() if family == "raft" else S.IGEVStereo({"max_disp": meta["max_disp"]})
        net.load_state_dict(sd, strict=True)
        net.eval()
        net.channels_last = True
        left, right = synth_pair(1, 64, 128, seed=seed, shift=meta["shift"])
        with torch.no_grad():
            out = net(left, right, iters=meta["iters"])
        assert any(m.weight.is_contiguous(memory_format=torch.channels_last) and not m.weight.is_contiguous()
                   for m in net.modules() if isinstance(m, torch.nn.Conv2d) and m.weight.shape[1] > 1 and m.weight.shape[2] > 1)
        assert all(m.weight.is_contiguous() for m in net.modules() if isinstance(m, torch.nn.Conv3d))
    assert (out - g["disp"]).abs().mean().item() < 1e-3


@pytest.mark.parametrize("key,ctor,seed", [("cfnet", "CFNet", 6), ("pcwnet_gc", "PCWNet_GC", 7)])
def test_channels_last_glue_cascade_models(key, ctor, seed):
    """The same opt-in for the two models with the largest 2-D networks (their gather warps / sampled volumes / refinement
    must not depend on the memory format)."""
    g = load_golden(f"{key}.npz")
    sd, meta = golden_state(key)
    with oracle_hot_path():
        import stereo_toolbox_b200 as S
        net = getattr(S, ctor)(meta["maxdisp"])
        net.load_state_dict(sd, strict=True)
        net.eval()
        net.channels_last = True
        left, right = _pair(meta, seed)
        with torch.no_grad():
            disp = net(left, right)
    assert (disp - g["disp"]).abs().mean().item() < 1e-3
