#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none ... --csv --log-file X.csv <cmd>`):
per kernel launches, mean / total device time and share -> markdown on stdout.

  python tools/launch_list_summary.py gpurun_out/r2k_launches.csv "python bench.py --steps 20 ..." > profiles/r2_ncu_launch_list.md
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    ci = {n: i for i, n in enumerate(head)}
    per = OrderedDict()
    for r in rd:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = r[ci["Kernel Name"]].replace("void ", "")
        name = name.split("(")[0] if not name.startswith("at::") else name[:80]
        per.setdefault(name, []).append(us)
    total = sum(sum(v) for v in per.values())
    print(f"# ncu launch list of `{what}`")
    print("# (`--metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised launches -- the SHARE of each kernel is what carries over to the timed run)\n")
    print("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {sum(v):.1f} | {100 * sum(v) / total:.1f}% |")


if __name__ == "__main__":
    main()
