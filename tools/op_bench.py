#!/usr/bin/env python
"""Per-operator timings of the hot path at the BASELINE config shapes, through the
same Python -> C-ABI calls the drop-in models make.

Every launch is bracketed by CUDA events on the launching stream; L2 is flushed (a 512 MB memset) before each
timed launch, so the numbers are cold-L2 like the full bench's.  `bytes` is the ALGORITHMIC (compulsory) traffic
of SURVEY 8d; frac = bytes / time / MEASURED_PEAKS.json hbm_gbs.

  python tools/op_bench.py [--iters 5] [--only name,...] [--json gpurun_out/op_bench.json]
  STB_OPBENCH_NCU=1: one untimed launch per op between cudaProfilerStart/Stop (for ncu --profile-from-start off)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stereo_toolbox_b200 as S                      # noqa: E402
from stereo_toolbox_b200 import ops                  # noqa: E402
from stereo_toolbox_b200.aggregation_umma import UmmaBackend   # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


class Bench:
    def __init__(self, iters):
        self.iters = iters
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
        self.rows = []
        self.ncu = bool(os.environ.get("STB_OPBENCH_NCU"))

    def run(self, name, fn, nbytes, flops=0.0, note=""):
        if self.ncu:
            fn(); torch.cuda.synchronize()
            torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
            return
        for _ in range(2):
            fn()
        ts = []
        for _ in range(self.iters):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        pk = peaks()
        row = {"op": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
               "GBps": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / pk["hbm_gbs"], 3),
               "TFLOPs": round(flops / ms / 1e9, 2), "note": note}
        self.rows.append(row)
        print(json.dumps(row), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    only = set(filter(None, a.only.split(",")))
    want = lambda n: not only or n in only
    torch.cuda.set_device(0)
    bn = Bench(a.iters)
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)

    # ---- K shape: 384x1248 -> 96x312 at 1/4, D/4 = 48, batch 8 (config 2)
    B, H, W, D = 8, 96, 312, 48
    if want("gwc_volume_f32"):
        l, r = rn(B, 320, H, W), rn(B, 320, H, W)
        bn.run("gwc_volume_f32", lambda: ops.gwc_volume(l, r, D, 40), 4.0 * (2 * l.numel() + B * 40 * D * H * W),
               note="K B=8, C=320 G=40, fp32 NCDHW out")
        del l, r
    if want("concat_volume_f32"):
        l, r = rn(B, 12, H, W), rn(B, 12, H, W)
        bn.run("concat_volume_f32", lambda: ops.concat_volume(l, r, D, True), 4.0 * (2 * l.numel() + B * 24 * D * H * W),
               note="K B=8, C=12 variant A")
        l, r = rn(2, 32, 144, 240), rn(2, 32, 144, 240)
        bn.run("concat_volume_f32_S", lambda: ops.concat_volume(l, r, D, True), 4.0 * (2 * l.numel() + 2 * 64 * D * 144 * 240),
               note="S B=2, C=32 (PSMNet)")
        del l, r
    if want("volume_cl16"):
        be = UmmaBackend("fp16")
        gl, gr, cl, cr = rn(B, 320, H, W), rn(B, 320, H, W), rn(B, 12, H, W), rn(B, 12, H, W)
        nb = 4.0 * 2 * (gl.numel() + cl.numel()) + 2.0 * B * D * H * W * 64
        bn.run("volume_cl16_gwc_concat", lambda: be.volume_gwc_concat(gl, gr, cl, cr, D, 40), nb,
               note="K B=8, gwc 40 + concat 2x12 -> NDHWC fp16 64 ch")
        l, r = rn(2, 32, 144, 240), rn(2, 32, 144, 240)
        bn.run("volume_cl16_concat_S", lambda: be.volume_concat(l, r, D, True), 4.0 * 2 * l.numel() + 2.0 * 2 * D * 144 * 240 * 64,
               note="S B=2, concat 2x32 (PSMNet) -> NDHWC fp16")
        del gl, gr, cl, cr, l, r
    if want("head"):
        cost = rn(B, D, H, W)
        bn.run("upsample_softargmin", lambda: ops.upsample_softargmin(cost, 192, 384, 1248, False),
               4.0 * (cost.numel() + B * 384 * 1248), note="K B=8; 736 M exp; reference materialises 6 x 2.9 GB")
        del cost
    # ---- R shape: 512x1024 -> 128x256 at 1/4, C=256 (config 4)
    if want("corr"):
        f1, f2 = rn(1, 256, 128, 256), rn(1, 256, 128, 256)
        blk = [None]

        def build():
            blk[0] = S.CorrBlock1D(f1, f2, 4, 4)
        pyr_bytes = 4.0 * 128 * 256 * (256 + 128 + 64 + 32 + 16)
        bn.run("corr1d_build_pyramid", build, 4.0 * 2 * f1.numel() + pyr_bytes, flops=2.0 * 128 * 256 * 256 * 256,
               note="R B=1 C=256: corr + 4 avg-pool levels")
        coords = torch.arange(256.0, device="cuda").view(1, 1, 1, 256).repeat(1, 2, 128, 1) - 17.3
        bn.run("corr1d_lookup", lambda: blk[0](coords), 4.0 * (2 * 36 * 128 * 256), note="R per GRU iteration, 4 levels x 9 taps")
        f1, f2 = rn(8, 256, 128, 256), rn(8, 256, 128, 256)
        bn.run("corr1d_build_pyramid_B8", build_b8(f1, f2, blk), 4.0 * 2 * f1.numel() + 8 * pyr_bytes,
               flops=8 * 2.0 * 128 * 256 * 256 * 256, note="R B=8")
        coords8 = coords.repeat(8, 1, 1, 1).contiguous()
        bn.run("corr1d_lookup_B8", lambda: blk[0](coords8), 8 * 4.0 * (2 * 36 * 128 * 256), note="R B=8 per iteration")
        del f1, f2
    # ---- M shape: 1152x1920, D=256 -> 288x480, D/4=64 (config 5)
    if want("geo"):
        Hm, Wm, Dm = 288, 480, 64
        f1, f2 = rn(1, 96, Hm, Wm), rn(1, 96, Hm, Wm)
        geo = rn(1, 8, Dm, Hm, Wm)
        enc = [None]

        def build():
            enc[0] = S.Combined_Geo_Encoding_Volume(f1, f2, geo, 2, 4)
        bn.run("geo_encoding_build", build, 4.0 * (2 * f1.numel() + 2 * geo.numel() * 1.5 + Hm * Wm * Wm * 1.5),
               flops=2.0 * 96 * Hm * Wm * Wm, note="M: corr C=96 + geo permute + 2-level pyramids")
        disp = torch.rand(1, 1, Hm, Wm, device="cuda") * 60
        coords = torch.arange(float(Wm), device="cuda").view(1, 1, 1, Wm).repeat(1, 1, Hm, 1)
        bn.run("geo_lookup", lambda: enc[0](disp, coords), 4.0 * 2 * 162 * Hm * Wm, note="M per iteration, 162 ch out")
        del f1, f2, geo
    if want("acv"):
        Hm, Wm, Dm = 288, 480, 64
        l, r = rn(1, 32, Hm, Wm), rn(1, 32, Hm, Wm)
        att = rn(1, 1, Dm, Hm, Wm)

        def acv():
            p = ops.softmax_d(att)
            return ops.concat_volume(l, r, Dm, False, att_prob=p)
        bn.run("acv_softmax_concat_unmasked", acv, 4.0 * (2 * l.numel() + 3 * att.numel() + 64 * Dm * Hm * Wm),
               note="M: softmax_D(att) * concat variant B, C=32, fp32 NCDHW")
    # ---- IGEV / CFNet elementwise pieces (M shape) and the training adjoints (S-crop shape 384x512 -> 96x128, D/4=48)
    if want("gate"):
        Hm, Wm, Dm = 288, 480, 64
        vol = rn(1, 16, Dm, Hm, Wm)
        gl = rn(1, 16, Hm, Wm)
        bn.run("feature_gate_f32", lambda: ops.feature_gate(vol, gl), 4.0 * (2 * vol.numel() + gl.numel()),
               note="M: IGEV FeatureAtt gate, 16 ch at 1/4 res")
        prob = torch.softmax(rn(2, 192, 384, 512), 1)
        dsp = ops.disparity_regression(prob, 192).unsqueeze(1)
        bn.run("disparity_variance", lambda: ops.disparity_variance(prob, 192, dsp), 4.0 * (prob.numel() + 2 * dsp.numel()),
               note="CFNet variance over a full-res probability volume, B=2 384x512")
        del vol, gl, prob, dsp
    if want("train"):
        from stereo_toolbox_b200 import _lib
        from stereo_toolbox_b200.ops import _p, _stream
        Bt, Ht, Wt, Dt = 2, 96, 128, 48
        x, gy = rn(Bt, 32, Dt, Ht, Wt), rn(Bt, 32, Dt, Ht, Wt)
        dw = torch.zeros(3, 3, 3, 32, 32, device="cuda")

        def wgrad():
            dw.zero_()
            _lib.call("stb_conv3d_wgrad_f32", _p(x), _p(gy), _p(dw), Bt, 32, Dt, Ht, Wt, 32, Dt, Ht, Wt, 3, 1, 1, _stream())
        bn.run("conv3d_wgrad_f32_32x32_k3", wgrad, 4.0 * (x.numel() + gy.numel()), flops=2.0 * 27 * 32 * 32 * Bt * Dt * Ht * Wt,
               note="weight gradient of a 32->32 k3 s1 layer, B=2 crop shape (fp32 FMA bound)")
        gvol = rn(Bt, 64, Dt, Ht, Wt)
        gl_, gr_ = torch.empty(Bt, 32, Ht, Wt, device="cuda"), torch.empty(Bt, 32, Ht, Wt, device="cuda")
        bn.run("concat_volume_bwd", lambda: _lib.call("stb_concat_volume_bwd_f32", _p(gvol), _p(gl_), _p(gr_), Bt, 32, Ht, Wt, Dt,
                                                      1, 64, 0, _stream()),
               4.0 * (gvol.numel() + 2 * gl_.numel()), note="adjoint of the PSMNet concat volume, B=2 crop shape")
        cost, gd = rn(Bt, Dt, Ht, Wt), rn(Bt, 384, 512)
        gc = torch.zeros_like(cost)

        def head_bwd():
            gc.zero_()
            _lib.call("stb_upsample_softargmin_bwd_f32", _p(cost), _p(gd), _p(gc), Bt, Dt, Ht, Wt, 192, 384, 512, 0, _stream())
        bn.run("upsample_softargmin_bwd", head_bwd, 4.0 * (2 * cost.numel() + gd.numel()),
               note="adjoint of the fused head, B=2 384x512 (MUFU + atomics bound)")
    if a.json and not bn.ncu:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        json.dump({"peaks": peaks(), "rows": bn.rows}, open(a.json, "w"), indent=1)


def build_b8(f1, f2, blk):
    def f():
        blk[0] = S.CorrBlock1D(f1, f2, 4, 4)
    return f


if __name__ == "__main__":
    main()
