import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _exact_library_math(request):
    """GPU tests compare with CPU fp32 fixtures of the reference: keep the torch glue around the kernels (cuDNN / cuBLAS,
    forward AND backward) in true fp32.  torch's default lets cuDNN convolutions use TF32, which alone moved a 2-D
    refinement-net gradient to cosine 0.44 against the fixture (round-2 hardware run, PCWNet_GC training step)."""
    if "gpu" not in request.keywords or not torch.cuda.is_available():
        yield
        return
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


# Tests that launch kernel code / kernel configurations which have never run on hardware (written after the round-1 GPU
# budget was spent) are opt-in: a deadlocked kernel would take the whole GPU run with it.  tools/trip_r2.sh sets the switch.
import os as _os
unconfirmed_kernels = pytest.mark.skipif(_os.environ.get("STB200_RUN_UNCONFIRMED") != "1",
                                         reason="launches kernels not yet confirmed on hardware; set STB200_RUN_UNCONFIRMED=1")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_meta(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_state(model_key, meta_file="models.json", calib=True):
    """State dict for a golden model fixture: name-keyed synthetic weights (+ calibrated BN stats)."""
    from stereo_toolbox_b200.synth import synth_state_dict, state_checksum
    meta = load_meta(meta_file)[model_key]
    template = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
                for k, s in meta["keys"].items()}
    over = None
    if calib:
        z = np.load(os.path.join(GOLDEN, f"bn_calib_{model_key}.npz"))
        over = {k: z[k] for k in z.files}
    sd = synth_state_dict(template, meta.get("seed", 0), over)
    assert abs(state_checksum(sd) - meta["checksum"]) <= 1e-6 * abs(meta["checksum"]), "synthetic weights drifted"
    return sd, meta
