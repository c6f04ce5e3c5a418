#!/usr/bin/env python
"""ncu raw CSV (ncu -i x.ncu-rep --page raw --csv) -> the metrics JSON kept under profiles/.
  python tools/ncu_summary.py gpurun_out/prof_raw.csv profiles/r1_step_kernel_ncu_metrics.json [row]
Keeps the metrics DESIGN.md and profiles/*.md quote; `row` picks the launch (default 0)."""
import csv, json, sys

KEEP = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
r = rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
out = {"kernel": r[hdr.index("Kernel Name")]}
for i, h in enumerate(hdr):
    if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        out[h] = [r[i], units[i]]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{len(out) - 1} metrics of {out['kernel'][:60]} -> {sys.argv[2]}")
