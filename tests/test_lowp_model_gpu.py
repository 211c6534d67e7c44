"""The 16-bit CUDA path against the oracle's 16-bit STORAGE MODEL (oracle/ref_lowp.py): the reference's arithmetic with
operands rounded exactly where the tcgen05 path rounds.  If the kernels add no error of their own, the CUDA output sits
on that model -- i.e. much closer to it than to the fp32 reference, whose distance is the price of the storage format
(tests/test_precision_model_cpu.py).  Written after the round-1 GPU budget was spent; the full-KITTI-shape numbers that
motivated it were measured separately (model on CPU 0.1159 px, CUDA path in bench.py 0.1165 px vs the fp32 reference)."""
import pytest
import torch

from conftest import golden_state
from oracle import ref_lowp, ref_models as M

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("prec,dtype", [("fp16", torch.float16), ("bf16", torch.bfloat16)])
def test_gwcnet_gc_16bit_sits_on_the_storage_model(prec, dtype):
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    sd, meta = golden_state("gwcnet_gc")
    b, h, w = meta["shape"]
    left, right = synth_pair(b, h, w, seed=0, shift=meta["shift"])
    ref32 = M.gwcnet_forward(sd, left, right, meta["maxdisp"], True)
    with ref_lowp.storage_16bit(dtype):
        model = M.gwcnet_forward(sd, left, right, meta["maxdisp"], True)
    net = S.GwcNet_GC(meta["maxdisp"], precision=prec)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.feature_tf32 = False                      # identical fp32 2-D features: the check is about the hot path
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    to_ref = (disp - ref32).abs().mean().item()
    to_model = (disp - model).abs().mean().item()
    model_to_ref = (model - ref32).abs().mean().item()
    print(f"GwcNet_GC {prec}: EPE vs fp32 reference {to_ref:.3e} px (storage model predicts {model_to_ref:.3e}), "
          f"vs 16-bit storage model {to_model:.3e} px")
    # The model predicts the SIZE of the rounding error, not its realisation: the kernel accumulates in a different order
    # than the CPU, values a hair from a 16-bit rounding boundary round the other way, and from there on the two carry
    # independent rounding noise of the same distribution (round-2 hardware run: 5.6e-3 measured / 5.7e-3 predicted for
    # fp16, 4.5e-2 / 4.8e-2 for bf16, and 4.6e-3 / 3.3e-2 between the two realisations).  So: the CUDA error must match
    # the predicted size (the kernels add nothing on top of the storage format) and the two realisations must be no
    # further apart than two independent draws.
    assert abs(to_ref / model_to_ref - 1.0) < 0.35, (to_ref, model_to_ref)
    assert to_model < 1.5 * max(to_ref, model_to_ref), (to_model, to_ref)


def test_head_x4_fast_path_matches_generic(tmp_path):
    """STB_HEAD_X4 (default on; 0 = generic kernel) selects ``upsample_softargmin_x4_kernel`` (outD == 4*D, align_corners=False: compile-time bin
    weights, no per-bin index arithmetic).  It evaluates the generic kernel's expressions in the generic kernel's order,
    so the two must agree to float rounding; the switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    import stereo_toolbox_b200 as S
    from oracle import ref_ops as R
    g = torch.Generator().manual_seed(11)
    cost = torch.randn(2, 1, 12, 9, 21, generator=g) * 3
    torch.save(cost, tmp_path / "cost.pt")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, torch; sys.path.insert(0, %r); import stereo_toolbox_b200 as S; "
            "c = torch.load(%r); torch.save(S.upsample_softargmin(c.cuda(), 48, 36, 84).cpu(), %r)"
            % (root, str(tmp_path / "cost.pt"), str(tmp_path / "x4.pt")))
    env = dict(os.environ, STB_HEAD_X4="0")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
    generic = torch.load(tmp_path / "x4.pt")
    fast = S.upsample_softargmin(cost.cuda(), 48, 36, 84).cpu()
    want = R.upsample_softargmin(cost, 48, 36, 84, False, False)
    torch.testing.assert_close(fast, want, rtol=1e-4, atol=1e-4)          # the oracle, like the generic kernel's test
    torch.testing.assert_close(fast, generic, rtol=0, atol=1e-4)


@pytest.mark.parametrize("switch", ["STB_UMMA_CLS1", "STB_UMMA_T2PAIR"])
def test_optin_instantiations_are_bit_identical(tmp_path, switch):
    """(Both default on since the round-2 hardware confirmation; the subprocess switches one OFF.)  STB_UMMA_CLS1 routes the 32->1 classifiers (kw-merged, fp32 out, no residual) to their own instantiation of
    the tcgen05 conv kernel (three single-column TMEM reads + 2 shuffles instead of three 32-column reads + 64);
    STB_UMMA_T2PAIR=1 routes the 64->32 merged transposed convs to the instantiation that stores the two w-parity classes
    as one contiguous pair.  Same accumulators, same additions in the same order, so the pre-softmax cost must be
    bit-identical.  The switches are read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); "
            "from conftest import golden_state; import stereo_toolbox_b200 as S; "
            "from stereo_toolbox_b200.synth import synth_pair; "
            "sd, meta = golden_state('gwcnet_gc'); net = S.GwcNet_GC(meta['maxdisp'], precision='fp16'); "
            "net.load_state_dict(sd); net = net.cuda().eval(); net.feature_tf32 = False; "
            "l, r = synth_pair(1, 64, 128, seed=0, shift=meta['shift']); "
            "d = net(l.cuda(), r.cuda()); torch.save((net._last_cost.cpu(), d.cpu()), %r)"
            % (root, os.path.join(root, "tests"), str(tmp_path / "cls1.pt")))
    subprocess.run([sys.executable, "-c", "import torch\nwith torch.no_grad():\n    exec(%r)" % code], check=True,
                   env=dict(os.environ, **{switch: "0"}), timeout=300)
    cost1, disp1 = torch.load(tmp_path / "cls1.pt")
    sd, meta = golden_state("gwcnet_gc")
    net = S.GwcNet_GC(meta["maxdisp"], precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    net.feature_tf32 = False
    left, right = synth_pair(1, 64, 128, seed=0, shift=meta["shift"])
    with torch.no_grad():
        disp0 = net(left.cuda(), right.cuda()).cpu()
    assert torch.equal(net._last_cost.cpu(), cost1)
    assert torch.equal(disp0, disp1)
