"""Compact per-kernel view of profiles/*_raw.txt (output of ncu_summary.py raw): one line per headline counter."""
import re
import sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst",
        "sm__warps_active.avg.pct", "launch__registers_per_thread ", "launch__occupancy_limit", "smsp__inst_executed.sum ", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled", "smsp__average_warp", "sass__inst_executed_local", "launch__grid_size", "launch__block_size", "shared_mem_per_block"]
for line in open(sys.argv[1]):
    if line.startswith("---") or any(k in line for k in KEYS):
        if "realtime" in line or ".max." in line or ".min." in line or ".sum.pct" in line or "per_second" in line:
            continue
        print(re.sub(r"\s{2,}", "  ", line.rstrip())[:170])
