"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    tot[r[ki]] += v; cnt[r[ki]] += 1
allt = sum(tot.values())
ours = sum(v for k, v in tot.items() if "ffq::" in k or k.startswith("calq") or k.startswith("w8a8") or "ffq" in k)
print(f"Total {allt / 1e3:.2f} ms over {sum(cnt.values())} launches; kernels of this repo: {100 * ours / allt:.1f} % of device time.\n")
print("| time (us) | share | launches | avg (us) | kernel |\n|---:|---:|---:|---:|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"| {v:.1f} | {100 * v / allt:.1f}% | {cnt[k]} | {v / cnt[k]:.1f} | `{k[:110]}` |")
