#!/usr/bin/env python
"""Turn an ncu launch list of ONE timed bench.py step into (i) the per-family time / DRAM-byte table committed under
profiles/ and (ii) profiles/ncu_traffic_r02.json, the only source bench.py accepts for ``roofline.traffic``.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --profile-from-start off --log-file gpurun_out/r2_launches.csv \
      env STB_CUDA_PROFILER=1 python bench.py --steps 1 --warmup 2 --no-extras --no-train --no-cpu-baseline
  python tools/ncu_traffic.py gpurun_out/r2_launches.csv fp16x2 > profiles/ncu_launches_r02_fp16x2.md

The 2-D extractor and the 3-D aggregation share the conv kernel; the volume builder's launch separates them."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, precision = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    order = []
    for r in rows[1:]:
        try:
            kid, name, metric, unit, val = r[ix["ID"]], r[ix["Kernel Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]]
        except (KeyError, IndexError):
            continue
        if kid not in per:
            per[kid] = {"name": name}
            order.append(kid)
        v = float(val.replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[kid][metric] = v * scale
    fam, seen_volume = {}, False
    for kid in order:
        k = per[kid]
        n = k["name"]
        if "volume_cl" in n:
            seen_volume = True
        if "conv3d_umma_kernel" in n:
            f = "conv3d_umma" if seen_volume else "conv2d_umma"
        elif "volume_cl" in n:
            f = "volume_cl16"
        elif "softargmin" in n:
            f = "upsample_softargmin"
        else:
            f = "other (" + n.split("(")[0].replace("void ", "")[:40] + ")"
        d = fam.setdefault(f, dict(launches=0, us=0.0, rd=0.0, wr=0.0))
        d["launches"] += 1
        d["us"] += k.get("gpu__time_duration.sum", 0.0)
        d["rd"] += k.get("dram__bytes_read.sum", 0.0)
        d["wr"] += k.get("dram__bytes_write.sum", 0.0)
    tot = sum(d["us"] for d in fam.values())
    print(f"# ncu launch list of one timed bench.py step ({precision}, B=8, 384x1248): per kernel family\n")
    print("(times under ncu are serialised and cold-cache: the SHARE of the step is what compares with bench.py's `kernels`)\n")
    print("| family | launches | time us | share | dram read GB | dram write GB |")
    print("|---|---|---|---|---|---|")
    for f, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| {f} | {d['launches']} | {d['us']:.0f} | {100 * d['us'] / tot:.1f} % | {d['rd'] / 1e9:.3f} | {d['wr'] / 1e9:.3f} |")
    out = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    js = json.load(open(out)) if os.path.exists(out) else {}
    for f, calls in (("conv3d_umma", 30), ("conv2d_umma", None)):
        if f in fam:
            d = fam[f]
            js[f"{precision}:{f}"] = {"dram_bytes_per_step": d["rd"] + d["wr"], "calls_per_step": calls or d["launches"],
                                      "launches_per_step": d["launches"], "time_us_under_ncu": d["us"],
                                      "source": f"ncu launch list of one timed step, {os.path.basename(path)} -> tools/ncu_traffic.py"}
    json.dump(js, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
