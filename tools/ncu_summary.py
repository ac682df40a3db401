#!/usr/bin/env python
"""Selected metrics of an `ncu --set full` report as a small CSV (one row per captured launch):

  python tools/ncu_summary.py gpurun_out/full_c2.ncu-rep > profiles/r02_ncu_full_c2_summary.csv
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    keep = [m for m in METRICS if m in col]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + ["%s [%s]" % (m, units[col[m]]) for m in keep])
    for r in rows[2:]:
        w.writerow([r[col["Kernel Name"]]] + [r[col[m]] for m in keep])


if __name__ == "__main__":
    main()
