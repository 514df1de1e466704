"""Summarise `ncu --page source --csv` output: per kernel, the SASS lines with the most stall
samples and a per-opcode roll-up.  Usage: ncu -i X.ncu-rep --page source --csv > f.csv;
python tools/ncu_source_summary.py f.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(starts):
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    sec = rows[s:end]
    h = next(j for j, r in enumerate(sec[:12]) if "Source" in r and "# Samples" in r)
    hdr = sec[h]
    c_src, c_smp, c_exe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")]
    body = [r for r in sec[h + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[c_smp] or 0) for r in body)
    print("=" * 100)
    print(sec[0][1][:110], "| sass lines", len(body), "| samples", tot)
    by_op = collections.Counter()
    exe_op = collections.Counter()
    for r in body:
        op = r[c_src].split()[0] if r[c_src].split() else "?"
        if op.startswith("@"):
            op = r[c_src].split()[1]
        op = op.split(".")[0]
        by_op[op] += int(r[c_smp] or 0)
        exe_op[op] += int(r[c_exe] or 0)
    print("opcode: samples% / executed (M warp-instr)")
    for op, n in by_op.most_common(14):
        print(f"   {op:10s} {100.0 * n / max(tot, 1):5.1f}%   {exe_op[op] / 1e6:9.1f}")
    stall_tot = collections.Counter()
    for r in body:
        for i in stall_cols:
            try:
                stall_tot[hdr[i]] += int(r[i] or 0)
            except ValueError:
                pass
    print("stalls:", ", ".join(f"{k[6:]}={100.0 * v / max(tot, 1):.1f}%" for k, v in stall_tot.most_common(8)))
    print("top lines:")
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][c_smp] or 0))[:top_n]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"   #{idx:5d} {100.0 * int(r[c_smp] or 0) / max(tot, 1):5.2f}%  {r[c_src][:70]:70s} {st}")
