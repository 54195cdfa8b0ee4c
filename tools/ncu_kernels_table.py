"""One line of key ncu metrics per captured kernel:  python tools/ncu_kernels_table.py gpurun_out/prof_X.ncu-rep > profiles/X_kernels.txt"""
import re, subprocess, sys
raw = subprocess.run([sys.executable, "tools/ncu_summary.py", "raw", sys.argv[1]], capture_output=True, text=True).stdout
blocks = re.split(r'^--- ', raw, flags=re.M)[1:]
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum ", "dram_rd"), ("dram__bytes_write.sum ", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "simt"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread ", "regs"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%")]
print("# ncu --set full --clock-control none, one launch per kernel of tools/prof_step.py (1280x960, spp 8: the first k_trace_queue launch is a full 2^23-sample chunk = 16.8 M rays)")
print("# source: %s (not committed); values per launch" % sys.argv[1])
for b in blocks:
    name = b.split('(')[0].strip().replace('void ', '')
    grid = re.search(r'grid \((\d+)', b).group(1)
    row = []
    for k, lab in keys:
        m = re.search(r'^' + re.escape(k) + r'\s+(\S+)\s+(\S+)\s*$', b, flags=re.M) or re.search(r'^' + re.escape(k.strip()) + r'\s+(\S*)\s+([0-9.,]+)\s*$', b, flags=re.M)
        if m:
            u, v = m.group(1), m.group(2)
            row.append("%s=%s%s" % (lab, v[:8], "" if u in ("%", "register/thread", "") else " " + u))
    print("%-28s grid %-6s %s" % (name[:28], grid, "  ".join(row)))
