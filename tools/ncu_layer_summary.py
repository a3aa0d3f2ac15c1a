#!/usr/bin/env python
"""Condense `ncu --page raw --csv --print-units base` of tools/prof_layer.py into a per-kernel table (markdown) and the
DRAM traffic per launch bench.py reports as `roofline.traffic` (json).  Only the launches of the LAST step are used.

    python tools/ncu_layer_summary.py raw.csv out.md out.json [steps=3] [launches of the last step]

The steady-state step of one decoder layer is 22 launches (4 per-tensor pairs, 7 per-channel weight launches, 7 linears);
earlier steps add fix-up launches, so the last step is taken by count when it is given."""
import csv
import json
import re
import sys
from collections import OrderedDict

raw, out_md, out_json = sys.argv[1:4]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
last_n = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rows = list(csv.reader(open(raw, newline="")))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
header = rows[start]
data = [r for r in rows[start + 2:] if len(r) == len(header)]
col = {n: i for i, n in enumerate(header)}


def num(r, name):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return None


def short(name):
    m = re.search(r"(ffq::)?(w4::)?([A-Za-z0-9_]+)\s*(<|\()", name)
    return m.group(3) if m else name[:40]


last = data[-last_n:] if last_n else data[len(data) - len(data) // steps:]
METRICS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "DRAM read MB", 1e-6),
    ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
    ("lts__t_bytes.sum", "L2 MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("sm__inst_executed_pipe_tensor_op_imma.avg.pct_of_peak_sustained_active", "imma pipe % (active)", 1),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_elapsed", "imma cycles % (elapsed)", 1),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
    ("smsp__cycles_active.avg", "SM active cycles", 1),
    ("sm__cycles_elapsed.avg.per_second", "SM clock GHz", 1e-9),
    ("launch__registers_per_thread", "regs", 1),
]
agg = OrderedDict()
for r in last:
    k = short(r[col["Kernel Name"]])
    a = agg.setdefault(k, {"n": 0})
    a["n"] += 1
    for name, _, scale in METRICS:
        v = num(r, name)
        if v is not None:
            a.setdefault(name, []).append(v * scale)
with open(out_md, "w") as f:
    f.write("# ncu --set full of one decoder layer's calibration step (tools/prof_layer.py, last of %d steps)\n\n" % steps)
    f.write("Per-kernel means over the launches of the last step; clocks not locked (`--clock-control none`), every launch\n"
            "replayed by the profiler, caches cold: durations are NOT benchmark numbers, the shares and the byte counts are\n"
            "what this capture is for.\n\n")
    have = [(n, t) for n, t, _ in METRICS if any(n in a for a in agg.values())]
    f.write("| kernel | launches | " + " | ".join(t for _, t in have) + " |\n|---|---|" + "---|" * len(have) + "\n")
    for k, a in agg.items():
        cells = []
        for n, _ in have:
            v = a.get(n)
            cells.append("-" if not v else f"{sum(v) / len(v):.3g}")
        f.write(f"| `{k}` | {a['n']} | " + " | ".join(cells) + " |\n")
    f.write("\nEvery launch of the last step, in order:\n\n| # | kernel | us | DRAM read MB | DRAM write MB | grid | block |\n|---|---|---|---|---|---|---|\n")
    for i, r in enumerate(last):
        g = r[col["Grid Size"]] if "Grid Size" in col else ""
        b = r[col["Block Size"]] if "Block Size" in col else ""
        f.write(f"| {i} | `{short(r[col['Kernel Name']])}` | {(num(r, 'gpu__time_duration.sum') or 0) * 1e-3:.1f} | "
                f"{(num(r, 'dram__bytes_read.sum') or 0) * 1e-6:.1f} | {(num(r, 'dram__bytes_write.sum') or 0) * 1e-6:.1f} | {g} | {b} |\n")
traffic = {}
for k, a in agg.items():
    rd, wr = a.get("dram__bytes_read.sum"), a.get("dram__bytes_write.sum")
    if rd and wr:
        traffic[k] = {"launches_in_capture": a["n"], "dram_bytes_per_launch": int((sum(rd) + sum(wr)) / a["n"] * 1e6),
                      "dram_read_bytes_per_launch": int(sum(rd) / a["n"] * 1e6),
                      "dram_write_bytes_per_launch": int(sum(wr) / a["n"] * 1e6)}
traffic["_source"] = ("ncu --set full capture of one decoder layer's calibration step (tools/prof_layer.py, summary in "
                      + out_md + "): dram__bytes_read.sum + dram__bytes_write.sum averaged over the last step's launches of each kernel")
json.dump(traffic, open(out_json, "w"), indent=1)
print(open(out_md).read()[:3000])
