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
    net = (S.GwcNet_GC if key == "gwcnet_gc" else S.GwcNet_G)(meta["maxdisp"])
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
    net = S.PSMNet(meta["maxdisp"])
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
    net = S.GwcNet_GC(48)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        disp = net(left.cuda(), right.cuda()).cpu()
    torch.testing.assert_close(net._last_cost.cpu(), aux["cost3"], rtol=2e-3, atol=2e-3)
    assert (disp - want).abs().mean().item() < 1e-3
