#!/usr/bin/env python
"""Selected metrics of ONE profiled launch (`ncu -i rep --page raw --csv`, default units) as a markdown table.

    python tools/ncu_one_kernel_md.py raw.csv "title" > profiles/xyz.md"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "derived__lts__lts2xbar_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpc__cycles_elapsed.avg.per_second"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
header, units = rows[start], rows[start + 1]
data = [r for r in rows[start + 2:] if len(r) == len(header)]
col = {n: i for i, n in enumerate(header)}
r = data[-1]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu --set full, one launch'}\n")
print(f"Kernel: `{r[col['Kernel Name']]}`, grid {r[col['Grid Size']]} x block {r[col['Block Size']]}; launch {len(data)} of {len(data)} in the report.\n")
print("| metric | value | unit |\n|---|---|---|")
for m in WANT:
    if m in col and r[col[m]] not in ("", "n/a"):
        print(f"| `{m}` | {r[col[m]]} | {units[col[m]]} |")
