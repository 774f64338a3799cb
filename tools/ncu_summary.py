#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` (+ optionally `--page source --csv`) into the
short text summaries committed under profiles/.

  python tools/ncu_summary.py RAW.csv [SOURCE.csv] > profiles/NAME.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:82s} {d[k]:>16s} {u.get(k, '')}")
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr, data = rows[1], rows[2:]
        isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[isamp]) for r in data)
        totex = sum(int(r[iex]) for r in data)
        print(f"\nsource page: {len(data)} SASS instructions, {totex} warp-instructions executed, {tot} stall samples")
        agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in st}
        print("stall reasons (share of samples):",
              ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        ops = {}
        for r in data:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0].rstrip(";")
            ops[op] = ops.get(op, 0) + int(r[iex])
        print("executed opcode mix:", ", ".join(f"{k} {100 * v / totex:.1f}%" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]))
        print("hottest instructions (share of samples, top stall):")
        for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
            top = max(((hdr[i][6:], int(r[i] or 0)) for i in st), key=lambda x: x[1])
            print(f"  {100 * int(r[isamp]) / tot:5.2f}%  {r[isrc][:78]:78s} {top[0]}")


if __name__ == "__main__":
    main()
