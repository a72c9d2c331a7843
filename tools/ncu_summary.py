#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_k2_v1.ncu-rep profiles/r01_k2_tile_match_ncu_full.json [note]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"Kernel Name": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                d[h] = {"value": v, "unit": u}
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                d.setdefault("stalls_per_issue_active", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = v
        res.append(d)
    json.dump({"source": rep, "note": note, "launches": res}, open(out, "w"), indent=1)
    print(out, len(res), "launch(es)")


if __name__ == "__main__":
    main()
