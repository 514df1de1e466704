"""Text summary of an .ncu-rep (run where ncu is installed): per kernel, the metrics DESIGN.md and
bench.py's roofline quote.  Usage: python tools/ncu_export.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print(f"# {rep}: ncu --set full --clock-control none (one replayed launch per row; cold caches)")
for r in rows[2:]:
    print("=" * 100)
    print(r[ki])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i]:>18s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
            try:
                stalls.append((int(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError:
                pass
    tot = sum(s[0] for s in stalls) or 1
    print("  warp-state samples: " + ", ".join(f"{n} {100 * c / tot:.1f}%" for c, n in sorted(stalls, reverse=True)[:8]))
