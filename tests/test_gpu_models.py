"""Model-level parity: drop-in models (CUDA hot path) vs the golden fixtures made by the reference
and vs the oracle restatement.  EPE tolerance from BASELINE.json north_star: <=1e-3 px (fp32)."""
import pytest
import torch

from conftest import load_golden, golden_state
from oracle import ref_models as M

pytestmark = pytest.mark.gpu


def _pair(meta):
    from stereo_toolbox_b200.synth import synth_pair
    b, h, w = meta["shape"]
    return synth_pair(b, h, w, seed=1 if h == 256 else 0, shift=meta["shift"])


@pytest.mark.parametrize("key", ["gwcnet_gc", "gwcnet_g"])
def test_gwcnet_golden_fp32(key):
    import stereo_toolbox_b200 as S
    g = load_golden(f"{key}.npz")
    sd, meta = golden_state(key)
    net = (S.GwcNet_GC if key == "gwcnet_gc" else S.GwcNet_G)(meta["maxdisp"], precision="fp32")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = _pair(meta)
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    assert disp.shape == g["disp"].shape
    if "cost3" in g:
        torch.testing.assert_close(net._last_cost.cpu(), g["cost3"], rtol=2e-3, atol=2e-3)
    epe = (disp - g["disp"]).abs().mean().item()
    assert epe < 1e-3, f"EPE vs reference {epe}"
    assert (disp - g["disp"]).abs().max().item() < 2e-2


def test_psmnet_golden_fp32():
    import stereo_toolbox_b200 as S
    g = load_golden("psmnet.npz")
    sd, meta = golden_state("psmnet")
    net = S.PSMNet(meta["maxdisp"], precision="fp32")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = _pair(meta)
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    assert disp.shape == g["disp"].shape == (1, 1, 256, 256)
    epe = (disp - g["disp"]).abs().mean().item()
    assert epe < 1e-3, f"EPE vs reference {epe}"


def test_gwcnet_oracle_fp32_wider():
    """A different size (W not a multiple of 32, D/4=12) against the oracle restatement."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    sd, _ = golden_state("gwcnet_gc")
    left, right = synth_pair(1, 96, 176, seed=5, shift=9)
    want, aux = M.gwcnet_forward(sd, left, right, 48, True, return_aux=True)
    net = S.GwcNet_GC(48, precision="fp32")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    torch.testing.assert_close(net._last_cost.cpu(), aux["cost3"], rtol=2e-3, atol=2e-3)
    assert (disp - want).abs().mean().item() < 1e-3


def test_raft_stereo_golden():
    """BASELINE config 4 family: RAFT-Stereo with CorrBlock1D (all-pairs corr + pyramid + per-iteration lookup) on the
    CUDA path vs the reference's own output."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("raft_stereo.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    net = S.RAFTStereo()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.update_mode = "torch"              # this test pins the torch / cuDNN update block (the default is the tensor-core one)
    left, right = synth_pair(1, 64, 128, seed=2, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the 2-D encoders/GRUs are torch; keep them fp32 for the comparison
    try:
        with torch.no_grad():
            out = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert out.shape == g["disp"].shape == (1, 1, 64, 128)
    epe = (out - g["disp"]).abs().mean().item()
    assert epe < 1e-3, f"EPE vs reference {epe}"


def test_raft_stereo_cuda_graph_iteration_is_bit_identical():
    """model.cuda_graph = True replays one captured GRU iteration (lookup kernel + update block) instead of launching it
    eagerly: same kernels in the same order, so the output must be bit-identical, and still match the reference."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("raft_stereo.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    net = S.RAFTStereo()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.update_mode = "torch"              # this test pins the torch / cuDNN update block (the default is the tensor-core one)
    left, right = synth_pair(1, 64, 128, seed=2, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            eager = net(left.cuda(), right.cuda(), iters=meta["iters"])
            net.cuda_graph = True
            graphed = net(left.cuda(), right.cuda(), iters=meta["iters"])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert torch.equal(eager, graphed)
    assert (graphed.cpu() - g["disp"]).abs().mean().item() < 1e-3


def test_raft_stereo_oracle_512x1024_shape_slice():
    """A wider pair (W/4 = 96 > one lookup tile, 8 iterations) against the same model run on CPU with the oracle
    CorrBlock1D."""
    import stereo_toolbox_b200 as S
    import stereo_toolbox_b200.raft_stereo as rs
    from oracle import ref_ops as R
    from stereo_toolbox_b200.synth import synth_pair
    sd, meta = golden_state("raft_stereo", calib=False)
    left, right = synth_pair(1, 96, 384, seed=4, shift=11)
    cpu = rs.RAFTStereo()
    cpu.load_state_dict(sd)
    cpu.eval()
    old = rs.CorrBlock1D
    rs.CorrBlock1D = R.CorrBlock1D
    try:
        with torch.no_grad():
            want = cpu(left, right, iters=8)
    finally:
        rs.CorrBlock1D = old
    net = S.RAFTStereo()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    net.update_mode = "torch"              # this test pins the torch / cuDNN update block (the default is the tensor-core one)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            got = net(left.cuda(), right.cuda(), iters=8).cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert (got - want).abs().mean().item() < 1e-3


def test_acvnet_golden_fp32():
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("acvnet.npz")
    sd, meta = golden_state("acvnet")
    net = S.ACVNet(meta["maxdisp"], precision="fp32")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = synth_pair(1, 64, 144, seed=3, shift=meta["shift"])
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    assert disp.shape == g["disp"].shape
    torch.testing.assert_close(net._last_att.cpu(), g["att"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(net._last_cost.cpu(), g["cost2"], rtol=2e-3, atol=2e-3)
    epe = (disp - g["disp"]).abs().mean().item()
    assert epe < 1e-3, f"EPE vs reference {epe}"
    net.attn_weights_only = True
    with torch.no_grad():
        att_disp = net(left.cuda(), right.cuda()).cpu()
    want = M.acvnet_forward(sd, left, right, meta["maxdisp"], attn_weights_only=True)
    assert (att_disp - want).abs().mean().item() < 1e-3


def test_acvnet_oracle_fp32_padded_shape():
    """H/16 and W/16 both ragged against the 4x4x4 attention blocks (pad_b > 0 and pad_r > 0: the masked case)."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    sd, _ = golden_state("acvnet")
    left, right = synth_pair(1, 96, 112, seed=6, shift=4)          # 1/16 scale: 6 x 7
    want = M.acvnet_forward(sd, left, right, 64)
    net = S.ACVNet(64, precision="fp32")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    epe = (disp - want).abs().mean().item()
    assert epe < 1e-3, f"EPE vs oracle {epe}"


def test_cfnet_golden():
    """CFNet whole model (fused 1/8-1/16-1/32 volumes + two cascade stages) vs the reference's own output (generated on CPU,
    tests/golden/make_golden.py cfnet).  The cascade rounds its search ranges to integers (floor / ceil / .long() samples),
    so a last-bit difference upstream can move one sample of one pixel: the fp32 bar is the median / mean error plus a
    bound on the fraction of such pixels.  The 16-bit path is compared where the model is still continuous -- the
    first-stage disparity before any integer sampling -- against the fp32 path; downstream of the integer samplers an
    untrained (flat-distribution) network turns 1e-2 px of upstream difference into whole-sample jumps."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("cfnet.npz")
    sd, meta = golden_state("cfnet")
    left, right = synth_pair(1, 64, 128, seed=6, shift=meta["shift"])
    stage1 = {}
    for precision in ("fp32", "fp16"):
        net = S.CFNet(meta["maxdisp"], precision=precision)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False          # identical fp32 2-D features: the comparison is about the hot path
        try:
            with torch.no_grad():
                disp = net(left.cuda(), right.cuda()).cpu()
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        assert disp.shape == g["disp"].shape == (1, 64, 128) and torch.isfinite(disp).all()
        stage1[precision] = net._last["pred2_s4"].cpu()
        if precision == "fp32":
            err = (disp - g["disp"]).abs()
            assert err.median().item() < 1e-3 and err.mean().item() < 5e-3, (err.median().item(), err.mean().item())
            assert (err > 0.1).float().mean().item() < 0.01
    d = (stage1["fp16"] - stage1["fp32"]).abs().mean().item()
    assert d < 1e-2, f"first-stage disparity (1/8 scale) fp16 vs fp32 path: {d} px"


def test_pcwnet_gc_golden():
    """PCWNet_GC whole model (4-scale volumes fused by the 3-level hourglassup, align_corners=True head, full-resolution
    2-D refinement) vs the reference's own CPU output: pre-softmax cost of classif3 and the final disparity (fp32 path);
    the 16-bit path is compared on the cost-volume stage's disparity (before the 2-D refinement net)."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("pcwnet_gc.npz")
    sd, meta = golden_state("pcwnet_gc")
    left, right = synth_pair(1, 64, 128, seed=7, shift=meta["shift"])
    stage = {}
    for precision in ("fp32", "fp16x2", "fp16"):
        net = S.PCWNet_GC(meta["maxdisp"], precision=precision)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                disp = net(left.cuda(), right.cuda()).cpu()
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        assert disp.shape == g["disp"].shape == (1, 64, 128) and torch.isfinite(disp).all()
        stage[precision] = net._last["pred3"].cpu()
        if precision == "fp32":
            torch.testing.assert_close(net._last_cost.cpu(), g["cost3"], rtol=2e-3, atol=2e-3)
        if precision in ("fp32", "fp16x2"):      # fp16x2: the whole 3-D path (Mish epilogues, 3-level hourglassup) on the exact tensor-core format
            epe = (disp - g["disp"]).abs().mean().item()
            print(f"PCWNet_GC {precision}: EPE vs reference {epe:.3e} px")
            assert epe < 1e-3, f"{precision}: EPE vs reference {epe}"
    d = (stage["fp16"] - stage["fp32"]).abs().mean().item()
    assert d < 1e-2, f"cost-volume stage disparity fp16 vs fp32 path: {d} px"


@pytest.mark.parametrize("key", ["gwcnet_gc", "psmnet", "acvnet"])
def test_default_constructor_runs_the_exact_tensor_core_path(key):
    """precision='auto' (the constructor default of GwcNet / PSMNet / ACVNet): the first CUDA inference forward switches the
    model to 'fp16x2' -- a drop-in user who never calls set_precision gets the tensor-core path, inside the fp32 bar."""
    import stereo_toolbox_b200 as S
    g = load_golden(f"{key}.npz")
    sd, meta = golden_state(key)
    net = {"gwcnet_gc": S.GwcNet_GC, "psmnet": S.PSMNet, "acvnet": S.ACVNet}[key](meta["maxdisp"])
    assert net.precision == "fp32" and net._auto_precision            # nothing resolved before a CUDA forward
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    if key == "acvnet":
        from stereo_toolbox_b200.synth import synth_pair
        left, right = synth_pair(1, 64, 144, seed=3, shift=meta["shift"])
    else:
        left, right = _pair(meta)
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    assert net.precision == "fp16x2" and not net._auto_precision
    epe = (disp - g["disp"]).abs().mean().item()
    print(f"{key} default constructor -> {net.precision}: EPE vs reference {epe:.3e} px")
    assert epe < 1e-3
