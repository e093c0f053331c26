#!/usr/bin/env python
"""Key metrics of every launch in an .ncu-rep as JSON (the tracked profiles/*_ncu_summary.json files are made with this).

usage: tools/ncu_summary.py REPORT.ncu-rep "how the report was captured" [label ...] > profiles/NAME.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
]


def main():
    rep, how, labels = sys.argv[1], sys.argv[2], sys.argv[3:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in col]
    res = {"source": how, "units": {k: units[col[k]] for k in keys}, "launches": []}
    for n, r in enumerate(rows[2:]):
        d = {"kernel": r[col["Kernel Name"]][:72]}
        if n < len(labels):
            d["label"] = labels[n]
        for k in keys:
            d[k] = r[col[k]]
        res["launches"].append(d)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
