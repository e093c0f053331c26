"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one line per launch (kernel, grid, us)."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, seq = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                if d.get("Metric Unit") == "ns":
                    v /= 1e3
                elif d.get("Metric Unit") == "ms":
                    v *= 1e3
                seq.append((d["Kernel Name"], d.get("Grid Size", ""), v))
    return seq


if __name__ == "__main__":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
    for name, grid, us in load(sys.argv[1])[:n]:
        print(f"{name[:48]:48s} {grid:>14s} {us:9.1f}")
