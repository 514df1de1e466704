"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
summed device time, share and launch count.  Usage: python tools/launch_summary.py X.csv "header note" > X.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    agg[name][0] += ms
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
n = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off")
print("#   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-clocks --profiler-window   (strict fp32, 1x B200)")
print(f"# one timed step = one 3840x2160 forward: {n} launches, {tot:.2f} ms summed device time (cold-cache, serialised)")
for note in sys.argv[2:]:
    print(f"# {note}")
print("# ms_total  share  launches  kernel")
for name, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:9.3f}  {100 * ms / tot:5.2f}%  {c:4d}  {name[:100]}")
