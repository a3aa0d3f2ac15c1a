#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics <stall ratios ...> --csv` log: issue activity and the stall reasons per issued
instruction (warps waiting per issue slot), averaged over the launches of each kernel."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ii = hdr.index("ID")
acc = defaultdict(lambda: defaultdict(list))
for r in rows[1:]:
    try:
        acc[r[ki]][r[mi]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
short = {"no_instruction": "no_inst", "long_scoreboard": "long_sb", "math_pipe_throttle": "math", "short_scoreboard": "short_sb",
         "lg_throttle": "lg", "barrier": "bar", "wait": "wait"}
print("| kernel | launches | avg us | issue % | warps % | dram % | " + " | ".join(short.values()) + " |")
print("|---|---|---|---|---|---|" + "---|" * len(short))
for k, m in sorted(acc.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    def avg(name):
        v = m.get(name, [])
        return sum(v) / len(v) if v else float("nan")
    t = avg("gpu__time_duration.sum")
    t = t / 1e3 if t > 5e3 else t
    cells = [f"{avg(f'smsp__average_warps_issue_stalled_{n}_per_issue_active.ratio'):.2f}" for n in short]
    name = k.replace("void ffq::", "").replace("ffq::", "")[:70]
    print(f"| `{name}` | {len(m.get('gpu__time_duration.sum', []))} | {t:.1f} | {avg('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {avg('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | " + " | ".join(cells) + " |")
