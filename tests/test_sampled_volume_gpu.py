"""stb_sampled_volume_f32 (csrc/sampled.cu: CFNet's cascade-stage volume in one launch) vs the reference fixture and the
oracle, and inside the whole CFNet (STB_CFNET_SAMPLED=1, read at import: hence the subprocess)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import load_golden, golden_state
from oracle import ref_ops as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_sampled_volume_golden_and_oracle():
    from stereo_toolbox_b200 import ops
    g = load_golden("ops_sampled.npz")
    got = ops.sampled_volume(*(g[k].cuda() for k in ("gl", "gr", "cl", "cr", "samples")), 4).cpu()
    torch.testing.assert_close(got, g["vol"], rtol=1e-5, atol=1e-5)
    # CFNet stage shapes: 320 channels in 40 groups (cpg 8), 12 concat channels, 16 samples, W not a multiple of 4
    gen = torch.Generator().manual_seed(5)
    gl, gr = torch.randn(1, 320, 3, 78, generator=gen), torch.randn(1, 320, 3, 78, generator=gen)
    cl, cr = torch.randn(1, 12, 3, 78, generator=gen), torch.randn(1, 12, 3, 78, generator=gen)
    smp = torch.randint(-2, 50, (1, 16, 3, 78), generator=gen).float()
    got = ops.sampled_volume(gl.cuda(), gr.cuda(), cl.cuda(), cr.cuda(), smp.cuda(), 40).cpu()
    torch.testing.assert_close(got, R.sampled_volume(gl, gr, cl, cr, smp, 40), rtol=1e-5, atol=1e-5)


def test_cfnet_with_the_fused_sampled_volume(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, torch\n"
            f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
            "from conftest import golden_state\n"
            "import stereo_toolbox_b200 as S\n"
            "from stereo_toolbox_b200.synth import synth_pair\n"
            "sd, meta = golden_state('cfnet')\n"
            "net = S.CFNet(meta['maxdisp']); net.load_state_dict(sd); net = net.cuda().eval()\n"
            "l, r = synth_pair(1, 64, 128, seed=6, shift=meta['shift'])\n"
            "with torch.no_grad():\n"
            "    d = net(l.cuda(), r.cuda())\n"
            f"torch.save(d.cpu(), {str(tmp_path / 'cfnet.pt')!r})\n")
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, STB_CFNET_SAMPLED="1"), timeout=300)
    got = torch.load(tmp_path / "cfnet.pt")
    want = load_golden("cfnet.npz")["disp"]
    assert (got - want).abs().median().item() < 1e-3          # integer samplers: isolated whole-sample jumps allowed
