#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/ncu_summary.py full  gpurun_out/prof.ncu-rep profiles/r01_name      # --set full capture
    python scripts/ncu_summary.py list  gpurun_out/launches.csv profiles/r01_name      # gpu__time_duration launch list
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def full(report, out):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]
    launches = []
    for row in rows[2:]:
        entry = OrderedDict(kernel=row[header.index("Kernel Name")])
        for metric in METRICS:
            if metric in header:
                i = header.index(metric)
                entry[metric] = {"value": row[i], "unit": units[i]}

        def scaled(name, table):
            item = entry.get(name)
            return float(item["value"].replace(",", "")) * table.get(item["unit"], 1.0) if item else None

        read, write = scaled("dram__bytes_read.sum", TO_BYTES), scaled("dram__bytes_write.sum", TO_BYTES)
        duration = scaled("gpu__time_duration.sum", TO_US)
        if read is not None and write is not None:
            entry["dram_bytes_per_launch"] = read + write
            if duration:
                entry["dram_GBps_under_ncu"] = (read + write) / (duration * 1e-6) / 1e9
        launches.append(entry)
    with open(out + ".full.json", "w") as handle:
        json.dump({"source": report, "command": "ncu --set full --clock-control none --import-source on", "launches": launches},
                  handle, indent=1)
    print(out + ".full.json", len(launches), "launches")
    for entry in launches:
        print(" ", entry["kernel"][:70], entry.get("gpu__time_duration.sum"), entry.get("dram_bytes_per_launch"),
              entry.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", {}).get("value"),
              "regs", entry.get("launch__registers_per_thread", {}).get("value"))


def launch_list(path, out):
    with open(path) as handle:
        lines = [line for line in handle if line.startswith('"')]
    rows = list(csv.reader(lines))
    header = rows[0]
    k, v, u = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    totals, counts = OrderedDict(), OrderedDict()
    for row in rows[1:]:
        name = row[k].split("(")[0]
        value = float(row[v].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row[u], 1.0)
        totals[name] = totals.get(name, 0.0) + value
        counts[name] = counts.get(name, 0) + 1
    grand = sum(totals.values())
    with open(out + ".launches.md", "w") as handle:
        handle.write(f"# ncu launch list ({path}); gpu__time_duration.sum, --clock-control none (cold-cache, serialised)\n\n")
        handle.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name in sorted(totals, key=totals.get, reverse=True):
            handle.write(f"| `{name}` | {counts[name]} | {totals[name]:.1f} | {totals[name] / counts[name]:.1f} | "
                         f"{100 * totals[name] / grand:.2f} % |\n")
    print(open(out + ".launches.md").read())


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
