#!/usr/bin/env python
"""Per-source-line summary of one kernel launch in an .ncu-rep (needs -lineinfo and --import-source on at capture time).

usage: tools/ncu_lines.py REPORT.ncu-rep LAUNCH_INDEX [TOP_N]
Prints, per CUDA source line: stall samples (share), warp instructions executed (share), shared-memory wavefronts, and the
dominant stall reasons; then the kernel totals.  Reads `ncu --page source --print-source cuda,sass --csv`.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, launch = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # sections: "File Path",<file> / blank / header / one aggregated row per CUDA source line followed by its SASS rows
    per = defaultdict(lambda: defaultdict(float))
    fname, hdr, col, stall_cols = "?", None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            col = {}
            for i, h in enumerate(hdr):
                col.setdefault(h, i)
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not r[0].isdigit():
            continue
        key = (f"{fname.replace('sg_', '').replace('.cu', '')}:{r[0]}", r[1].strip()[:100])
        for name in ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive",
                     "L2 Theoretical Sectors Global"] + stall_cols:
            if name in col:
                try:
                    per[key][name] += float(r[col[name]])
                except ValueError:
                    pass
    tot = defaultdict(float)
    for k, v in per.items():
        for n, x in v.items():
            tot[n] += x
    ts, ti, tw = tot["# Samples"] or 1, tot["Instructions Executed"] or 1, tot["L1 Wavefronts Shared"] or 1
    print(f"totals: samples {ts:.0f}  warp-instructions {ti:.0f}  shared wavefronts {tw:.0f} (excessive {tot['L1 Wavefronts Shared Excessive']:.0f})")
    print("stall mix: " + "  ".join(f"{s[6:]} {100 * tot[s] / ts:.1f}%" for s in sorted(stall_cols, key=lambda s: -tot[s])[:8]))
    print(f"{'line':>14} {'samp%':>6} {'inst%':>6} {'wave%':>6}  top stalls | source")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        st = sorted(stall_cols, key=lambda s: -v[s])[:2]
        sts = " ".join(f"{s[6:]}:{100 * v[s] / max(v['# Samples'], 1):.0f}" for s in st)
        print(f"{k[0]:>14} {100 * v['# Samples'] / ts:6.2f} {100 * v['Instructions Executed'] / ti:6.2f} {100 * v['L1 Wavefronts Shared'] / tw:6.2f}  {sts:28s} | {k[1]}")


if __name__ == "__main__":
    main()
