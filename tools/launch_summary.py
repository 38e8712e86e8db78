"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of the total per kernel name.
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/xyz.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn]
    tot[name] += float(r[mv].replace(",", "")) / 1000.0      # ns -> us
    cnt[name] += 1
total = sum(tot.values())
print("| share | total us | launches | kernel |\n|---:|---:|---:|---|")
for name, t in tot.most_common():
    print(f"| {100 * t / total:.1f}% | {t:.0f} | {cnt[name]} | `{name[:110]}` |")
print(f"\nTotal {total / 1000:.1f} ms over {sum(cnt.values())} launches.")
