"""IGEVStereo whole model (BASELINE config 5 family) on the CUDA hot path vs the reference's own output
(tests/golden/igev_stereo.npz, made by tests/golden/make_golden.py igev_model).

The cost-volume stage and the geometry lookup are the kernels already pinned at block level at exactly these shapes
(tests/test_gpu_blocks.py::test_igev_cost_volume); this file checks them inside the whole drop-in model: features ->
stage -> 4 GRU iterations (one lookup launch each) -> convex upsampling."""
import pytest
import torch

from conftest import load_golden, golden_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _run(precision, **fwd):
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    sd, meta = golden_state("igev_stereo")
    net = S.IGEVStereo({"max_disp": meta["max_disp"]}, precision=precision)
    net.load_state_dict(sd, strict=True)          # the reference's own parameter names (timm-named MobileNetV2 trunk)
    net = net.cuda().eval()
    net.update_mode = "torch"                     # these tests pin the torch / cuDNN update block (the default is the tensor-core one)
    left, right = synth_pair(1, 64, 128, seed=8, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the 2-D networks are torch glue; keep them fp32 for the comparison
    try:
        with torch.no_grad():
            return net(left.cuda(), right.cuda(), **fwd), meta
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_igev_stereo_golden_fp32():
    g = load_golden("igev_stereo.npz")
    sd, meta = golden_state("igev_stereo")
    out, _ = _run("fp32", iters=meta["iters"])
    out = out.cpu()
    assert out.shape == g["disp"].shape == (1, 1, 64, 128)
    epe = (out - g["disp"]).abs().mean().item()
    assert epe < 1e-3, f"EPE vs reference {epe}"          # px, north_star fp32 bar


def test_igev_stereo_golden_fp16x2():
    """The whole model with the cost-volume stage on the exact tensor-core path as well (precision='fp16x2': 48-channel level
    as three K-chunks, k4-s2 transposed convs, split-storage feature gates) and the update block on tcgen05 (default)."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("igev_stereo.npz")
    sd, meta = golden_state("igev_stereo")
    net = S.IGEVStereo({"max_disp": meta["max_disp"]}, precision="fp16x2")
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    left, right = synth_pair(1, 64, 128, seed=8, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    epe = (out - g["disp"]).abs().mean().item()
    print(f"IGEVStereo fp16x2 stage + tcgen05 update block: EPE vs reference {epe:.3e} px")
    assert epe < 1e-3, f"EPE vs reference {epe}"


def test_igev_stereo_train_style_return_fp32():
    """test_mode=False in eval: (upsampled initial disparity, one prediction per iteration) -- igev_stereo.py:254-255."""
    g = load_golden("igev_stereo.npz")
    (init_up, preds), _ = _run("fp32", iters=2, test_mode=False)
    assert len(preds) == 2 and init_up.shape == (1, 1, 64, 128)
    assert (init_up.cpu() - g["init_disp_up"]).abs().mean().item() < 1e-3
    assert (preds[-1].cpu() - g["pred_last"]).abs().mean().item() < 1e-3


def test_igev_stereo_fp16_runs():
    """16-bit tensor-core stage inside the whole model.  The stage's own 16-bit bar (<=1e-2 px on init_disp) is asserted
    at block level; behind it sit 4 GRU iterations of an untrained network, so here the whole-model output is checked
    for shape / finiteness and its distance to the reference is reported, not bounded."""
    g = load_golden("igev_stereo.npz")
    out, meta = _run("fp16", iters=4)
    out = out.float().cpu()
    assert out.shape == g["disp"].shape
    assert torch.isfinite(out).all()
    print(f"IGEVStereo fp16 stage, whole-model EPE vs reference: {(out - g['disp']).abs().mean().item():.4f} px")


def test_igev_stereo_cuda_graph_iteration_matches_eager():
    """model.cuda_graph = True replays one captured GRU iteration (geometry-lookup kernel + update block) -- also on the
    second call, which only refreshes the graph's static inputs.  Our lookup kernel is deterministic, but cuDNN may pick a
    different algorithm for the update block's small convolutions under stream capture (it did on the round-2 hardware run:
    not bit-identical), so the bar is the parity bar, 1e-3 px max, not bit identity."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    sd, meta = golden_state("igev_stereo")
    net = S.IGEVStereo({"max_disp": meta["max_disp"]})
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.update_mode = "torch"
    left, right = synth_pair(1, 64, 128, seed=8, shift=meta["shift"])
    left2, right2 = synth_pair(1, 64, 128, seed=9, shift=3)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            eager = net(left.cuda(), right.cuda(), iters=4)
            eager2 = net(left2.cuda(), right2.cuda(), iters=4)
            net.cuda_graph = True
            graphed = net(left.cuda(), right.cuda(), iters=4)
            graphed2 = net(left2.cuda(), right2.cuda(), iters=4)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    d1, d2 = (eager - graphed).abs().max().item(), (eager2 - graphed2).abs().max().item()
    print(f"IGEV eager vs graph replay: max |diff| {d1:.3e} / {d2:.3e} px")
    assert d1 < 1e-3 and d2 < 1e-3


@pytest.mark.parametrize("graph", [False, True])
def test_igev_stereo_golden_update_on_tensor_cores(graph):
    """model.update_mode = 'umma': the ConvGRU update block on the tcgen05 2-D conv path in the exact 'fp16x2' format
    (update_umma.UmmaIgevUpdate), hidden states resident in kernel layout -- vs the reference's output, fp32 bar."""
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("igev_stereo.npz")
    sd, meta = golden_state("igev_stereo")
    net = S.IGEVStereo({"max_disp": meta["max_disp"]})
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    assert getattr(net, "update_mode", "auto") == "auto"          # the default: tensor-core update block for CUDA inference
    net.cuda_graph = graph
    left, right = synth_pair(1, 64, 128, seed=8, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
            if graph:
                # second call = replay with refreshed static inputs; the torch feature networks in front of the loop are not
                # bit-reproducible between calls (cuDNN algorithm choice, see the test above), so: the parity bar, not equality
                out2 = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
                assert (out - out2).abs().max().item() < 1e-3
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    epe = (out - g["disp"]).abs().mean().item()
    print(f"IGEVStereo, update block on tcgen05 (graph={graph}): EPE vs reference {epe:.3e} px")
    assert epe < 1e-3, f"EPE vs reference {epe}"
