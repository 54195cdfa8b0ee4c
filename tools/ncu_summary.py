"""Turn gpurun_out ncu artefacts into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/X_launches.txt
    python tools/ncu_summary.py raw gpurun_out/prof_X.ncu-rep       > profiles/X_raw.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "smsp__warp_issue_stalled", "sass__inst_executed_local", "launch__waves_per_multiprocessor",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}[r[mu]]
        a = agg[r[kn].split("(")[0][:80]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES) -- %s" % path)
    print("%-82s %6s %12s %7s" % ("kernel", "n", "total_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
        print("%-82s %6d %12.3f %7.3f" % (k, v[0], v[1], v[1] / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none -- %s" % path)
    for r in rows[2:]:
        print("--- %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:100], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, hname in enumerate(hdr):
            if any(k in hname for k in KEYS) and r[i] != "":
                print("%-90s %-16s %s" % (hname, units[i], r[i]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
