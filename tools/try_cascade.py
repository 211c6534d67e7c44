"""PCWNet_GC / CFNet: the exact tensor-core path ('fp16x2') against the fp32 CUDA-core path of the same drop-in model at the
KITTI shape (the fp32 path is pinned to the reference by the golden fixtures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_state
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.synth import synth_pair
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for key, ctor, seed in (("pcwnet_gc", S.PCWNet_GC, 7), ("cfnet", S.CFNet, 6)):
    sd, meta = golden_state(key)
    left, right = synth_pair(1, 384, 1248, seed=seed, shift=23)
    outs = {}
    for prec in ("fp32", "fp16x2"):
        net = ctor(192, precision=prec); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
        with torch.no_grad():
            out = net(left.cuda(), right.cuda())
        outs[prec] = (out[-1] if isinstance(out, (list, tuple)) else out).float().cpu()
        del net; torch.cuda.empty_cache()
    d = (outs["fp16x2"] - outs["fp32"]).abs()
    print(f"{key} 384x1248: fp16x2 vs fp32 path: mean {d.mean().item():.3e} px, median {d.median().item():.3e}, max {d.max().item():.3e}, |disp| mean {outs['fp32'].abs().mean().item():.1f}")
