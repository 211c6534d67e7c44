#!/usr/bin/env python
"""The reference's published Table 3 (README.md:182-203: whole-model latency / peak memory, batch 1, RTX 4090, fp32) re-run
with the drop-in models on this GPU, through the reference's own protocol (stereo_toolbox_b200.evaluation
.speed_and_memory_test == evaluation/speed_and_memory_test.py).  Prints a markdown table and one JSON line.

    python tools/table3.py --models gwcnet_gc psmnet --precisions fp16 fp32 --iters 20
Random-init weights (the protocol does not depend on their values); iterative models run 32 iterations (their default).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# RTX 4090 rows of the reference's Table 3: seconds at (480,640) / (736,1280) / (1088,1920)   
PUBLISHED = {"psmnet": (0.0396, 0.1245, 0.2866), "gwcnet_gc": (0.0386, 0.1326, 0.3093), "cfnet": (0.0481, 0.1434, 0.3343),
             "raft": (0.1967, 0.3624, 0.7613), "acvnet": (0.0494, 0.1664, 0.3848), "pcwnet_gc": (0.0888, 0.2769, 0.6419),
             "igev": (0.2363, 0.3501, 0.6741)}


def build(name, precision):
    import stereo_toolbox_b200 as S
    if name == "raft":
        return S.RAFTStereo()
    if name == "igev":
        return S.IGEVStereo(precision=precision)
    ctor = {"psmnet": S.PSMNet, "gwcnet_gc": S.GwcNet_GC, "cfnet": S.CFNet, "acvnet": S.ACVNet, "pcwnet_gc": S.PCWNet_GC}[name]
    return ctor(192, precision=precision)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", nargs="+", default=list(PUBLISHED))
    ap.add_argument("--precisions", nargs="+", default=["fp16", "fp32"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cuda-graph", action="store_true", help="raft / igev: model.cuda_graph = True")
    ap.add_argument("--channels-last", action="store_true", help="raft / igev / cfnet / pcwnet_gc: model.channels_last = True")
    args = ap.parse_args()
    from stereo_toolbox_b200.evaluation import speed_and_memory_test
    rows, out = [], {}
    for name in args.models:
        for prec in (["fp32"] if name == "raft" else args.precisions):
            try:
                net = build(name, prec)
                net.cuda_graph, net.channels_last = args.cuda_graph, args.channels_last
                _, secs, mbs = speed_and_memory_test(net, num_iterations=args.iters, verbose=False)
            except Exception as e:          # keep the table going: one model failing at one size is a finding, not a crash
                rows.append(f"| {name} | {prec} | failed: {type(e).__name__}: {str(e)[:80]} |")
                continue
            out[f"{name}/{prec}"] = {"seconds": secs, "peak_mb": mbs}
            pub = PUBLISHED[name]
            cells = " | ".join(f"{s:.4f} s ({p / s:.1f}x) / {m:.0f} MB" for s, p, m in zip(secs, pub, mbs))
            rows.append(f"| {name} | {prec} | {cells} |")
    print("| model | precision | (480,640): ours (vs RTX 4090 published) / peak | (736,1280) | (1088,1920) |")
    print("|---|---|---|---|---|")
    print("\n".join(rows))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
