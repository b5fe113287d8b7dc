#!/usr/bin/env python
"""Summarise an `ncu --set full` report: per kernel (mean over its captured launches) duration, DRAM bytes, occupancy,
issue statistics -> a markdown table on stdout and, with --traffic-json, the entry bench.py reads for `roofline.traffic`.

  python tools/ncu_summary.py gpurun_out/r1_final.ncu-rep --traffic-json profiles/traffic.json > profiles/r1_ncu_full_summary.md
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import re
import subprocess
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def short_name(full: str) -> str:
    m = re.match(r"(?:void )?([\w:]+)(<[^(]*>)?", full)
    name = m.group(1).split("::")[-1]
    targs = (m.group(2) or "").replace(" ", "")
    return name + targs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--traffic-json")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    per = defaultdict(lambda: defaultdict(list))
    for r in body:
        k = short_name(r[col["Kernel Name"]])
        for m in METRICS:
            if m in col and r[col[m]] not in ("", "n/a"):
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                per[k][m].append(v * UNIT_SCALE.get(units[col[m]], 1.0))
    traffic = {}
    print(f"# ncu --set full summary of `{a.report}` (cold-cache, serialised launches; means over the captured launches)\n")
    for k, ms in per.items():
        n = len(ms["gpu__time_duration.sum"])
        print(f"## {k}  ({n} launches)\n")
        print("| metric | mean | unit |\n|---|---|---|")
        for m in METRICS:
            if m in ms:
                u = units[col[m]]
                u = {"Kbyte": "byte", "Mbyte": "byte", "Gbyte": "byte", "nsecond": "usecond", "msecond": "usecond"}.get(u, u)
                print(f"| {m} | {sum(ms[m]) / len(ms[m]):.6g} | {u} |")
        print()
        if "dram__bytes_read.sum" in ms:
            traffic[k] = {"dram_bytes_read": sum(ms["dram__bytes_read.sum"]) / n, "dram_bytes_write": sum(ms["dram__bytes_write.sum"]) / n,
                          "gpu_time_us": sum(ms["gpu__time_duration.sum"]) / n, "launches": n, "report": a.report}
    if a.traffic_json:
        try:
            with open(a.traffic_json) as f:
                old = json.load(f)
        except (OSError, ValueError):
            old = {}
        old.update(traffic)
        with open(a.traffic_json, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
