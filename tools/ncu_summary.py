#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into one line per launch:
duration, DRAM read/write bytes, DRAM %, SM %, tensor-pipe %, registers, grid.  Usage:
  python tools/ncu_summary.py gpurun_out/ops_full.ncu-rep [--md]"""
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    return rows[2:], idx, units


def fnum(r, idx, units, name, to=None):
    if name not in idx:
        return float("nan")
    v = float(r[idx[name]].replace(",", "") or 0)
    u = units[idx[name]].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale


def main():
    rep = sys.argv[1]
    rows, idx, units = load(rep)
    print("| kernel | time us | dram read MB | dram write MB | DRAM GB/s | dram % | sm % | tensor % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:44]
        t = fnum(r, idx, units, "gpu__time_duration.sum")
        rd = fnum(r, idx, units, "dram__bytes_read.sum")
        wr = fnum(r, idx, units, "dram__bytes_write.sum")
        g = lambda n: r[idx[n]] if n in idx else "-"
        print(f"| {name} | {t*1e6:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/t/1e9:.0f} | "
              f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
              f"{float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
              f"{float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
              f"{g('launch__registers_per_thread')} | {g('launch__grid_size')} x {g('launch__block_size')} |")


if __name__ == "__main__":
    main()
