"""Summarise `ncu --page source --csv` output per opcode: executed warp instructions, share of the stall
samples, shared-memory wavefronts and the excess over the ideal count.
Usage: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_smem_summary.py f.csv [rows]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared Excessive",
                                "L1 Wavefronts Shared", "L1 Conflicts Shared N-Way")}


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


tot_w = sum(f(r[ci["L1 Wavefronts Shared"]]) for r in body)
tot_e = sum(f(r[ci["L1 Wavefronts Shared Excessive"]]) for r in body)
tot_s = sum(f(r[ci["# Samples"]]) for r in body)
tot_i = sum(f(r[ci["Instructions Executed"]]) for r in body)
print(f"# {rows[0][1][:100]}")
print(f"# warp instructions {tot_i / 1e6:.1f} M, stall samples {tot_s:.0f}, shared wavefronts {tot_w / 1e6:.1f} M "
      f"of which excessive {tot_e / 1e6:.1f} M ({100 * tot_e / max(tot_w, 1):.1f} %)")
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in body:
    parts = r[ci["Source"]].split()
    op = parts[0] if parts else "?"
    if op.startswith("@") and len(parts) > 1:
        op = parts[1]
    op = ".".join(op.split(".")[:2])
    a = agg[op]
    a[0] += f(r[ci["Instructions Executed"]])
    a[1] += f(r[ci["L1 Wavefronts Shared"]])
    a[2] += f(r[ci["L1 Wavefronts Shared Excessive"]])
    a[3] += 1
    a[4] += f(r[ci["# Samples"]])
print("# opcode                 sass  executed(M)  samples  wavefronts(M)  excessive(M)")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][4])[:top]:
    print(f"{op:24s} {a[3]:4d}  {a[0] / 1e6:10.2f}  {100 * a[4] / max(tot_s, 1):6.1f}%  {a[1] / 1e6:12.2f}  {a[2] / 1e6:11.2f}")
print("# lines with the most excessive wavefronts: source | executed | wavefronts | excessive | n-way")
for r in sorted(body, key=lambda r: -f(r[ci["L1 Wavefronts Shared Excessive"]]))[:8]:
    if f(r[ci["L1 Wavefronts Shared Excessive"]]) > 0:
        print("  ", r[ci["Source"]].strip()[:64], "|", r[ci["Instructions Executed"]], "|", r[ci["L1 Wavefronts Shared"]], "|",
              r[ci["L1 Wavefronts Shared Excessive"]], "|", r[ci["L1 Conflicts Shared N-Way"]])
