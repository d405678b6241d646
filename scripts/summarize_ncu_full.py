#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report, in the format of profiles/r0*_ncu_full_*.txt.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/summarize_ncu_full.py /tmp/raw.csv profiles/r02_ncu_full_sga.txt "header line ..."
"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    header = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none"
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["# " + header]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.append("kernel: %-70s id %s" % (r[col["Kernel Name"]][:70], r[col["ID"]]))
        for k in KEEP:
            if k in col:
                out.append("   %-78s %s %s" % (k, r[col[k]], units[col[k]]))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:24]))


if __name__ == "__main__":
    main()
