"""Prints a markdown table of the headline metrics of one kernel from `ncu -i X.ncu-rep --page raw --csv`
output (read from the file given as argv[1]); used to fill profiles/r01_ncu_summary.md."""
import csv, sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
]
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, units, vals = rows[0], rows[1], rows[2:]
for v in vals:
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {v[i]} | {units[i]} |")
    print()
