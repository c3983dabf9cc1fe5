"""Summarise an ncu report (read here, no GPU): per kernel duration, DRAM bytes, throughput %, stalls.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]
for r in data:
    print("=" * 100)
    print(r[idx["Kernel Name"]])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
    st = []
    for h, i in idx.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("  top stall reasons (pc samples): " + ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:6]))
