#!/usr/bin/env python
"""Turns what tools/r2_profile.sh left in gpurun_out/ into the committed summaries under profiles/:
the default bench line, the condensed ncu launch list (durations kept), one text summary per fully
captured kernel (raw + source pages) and r2_counters.json (per-launch counters bench.py reads for
`roofline.traffic` and `roofline.issue`).

  python tools/make_profiles.py [round_tag]      # default r2
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"

# units (pixel x depth label) one launch of each profiled workload processes
UNITS = {"cfg4": 1920 * 1080 * 256, "cfg3": 1920 * 1080 * 256}


def ncu_csv(rep, page, dst):
    with open(dst, "w") as f:
        subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)


def short_name(full):
    return re.sub(r"<.*|\(.*", "", full).replace("void ", "").replace("sr::", "")


# 1. bench line
line = [l for l in open(os.path.join(OUT, f"{tag}_bench_default.json")) if l.startswith("{")][-1]
open(os.path.join(PROF, f"{tag}_bench_default.json"), "w").write(line)

# 2. launch list of one one-view step (SR_LANES=1: one kernel at a time)
rows = [r for r in csv.reader(l for l in open(os.path.join(OUT, f"{tag}_launches_raw.csv")) if l.startswith('"'))]
ix = {h: i for i, h in enumerate(rows[0])}
keep = ["ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"]
tot, cnt = collections.Counter(), collections.Counter()
with open(os.path.join(PROF, f"{tag}_launches.csv"), "w", newline="") as f:
    wr = csv.writer(f, quoting=csv.QUOTE_ALL)
    wr.writerow(keep)
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        wr.writerow([r[ix["ID"]], name] + [r[ix[k]] for k in keep[2:]])
        tot[name] += int(r[ix["Metric Value"]])
        cnt[name] += 1
T = sum(tot.values())
share_txt = []
for k, v in tot.most_common():
    share_txt.append(f"{k:56s} n={cnt[k]:3d} avg {v / cnt[k] / 1e6:8.3f} ms  share {100 * v / T:5.1f}%")
print("\n".join(share_txt))
open(os.path.join(PROF, f"{tag}_launch_shares.txt"), "w").write("\n".join(share_txt) + "\n")

# 3. full captures -> text summaries + counters
counters = {}
for wl, rep in (("cfg4", f"prof_{tag}_final.ncu-rep"), ("cfg3", f"prof_{tag}_cfg3.ncu-rep")):
    if not os.path.exists(os.path.join(OUT, rep)):
        continue
    raw, src = os.path.join(OUT, f"raw_{tag}_{wl}.csv"), os.path.join(OUT, f"src_{tag}_{wl}.csv")
    ncu_csv(rep, "raw", raw)
    ncu_csv(rep, "source", src)
    rr = list(csv.reader(open(raw)))
    lines = open(src).read().split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
    seen = set()
    counters[wl] = {}
    for row in rr[2:]:
        d = dict(zip(rr[0], row))
        short = short_name(d["Kernel Name"])
        if short in seen:
            continue
        seen.add(short)
        one_raw, one_src = os.path.join(OUT, f"raw_{tag}_{short}.csv"), os.path.join(OUT, f"src_{tag}_{short}.csv")
        csv.writer(open(one_raw, "w")).writerows([rr[0], rr[1], row])
        sec = [i for i in range(len(starts) - 1) if short in lines[starts[i]]][0]  # first section of this kernel
        open(one_src, "w").write("\n".join(lines[starts[sec]:starts[sec + 1]]))
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), one_raw, one_src],
                             capture_output=True, text=True, check=True).stdout
        open(os.path.join(PROF, f"{tag}_{wl}_{short}.txt"), "w").write(txt)
        un = dict(zip(rr[0], rr[1]))
        SCALE = {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6,                      # durations -> ms
                 "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}  # sizes -> bytes

        def f(k):
            if d.get(k) in (None, ""):
                return None
            return float(d[k].replace(",", "")) * SCALE.get(un.get(k, ""), 1.0)
        units = UNITS[wl]
        if short == "build_refr_kernel":
            units = UNITS[wl]  # one launch = one neighbour: per (pixel, label, neighbour)
        counters[wl][short] = {
            "inst_executed": f("smsp__inst_executed.sum"), "units": units,
            "duration_ms": f("gpu__time_duration.sum"),
            "dram_bytes": int(f("dram__bytes_read.sum") + f("dram__bytes_write.sum")),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_fp64_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "registers": f("launch__registers_per_thread"), "warps_active_per_sm": f("sm__warps_active.avg.per_cycle_active"),
        }
        print(wl, short, counters[wl][short]["duration_ms"], "ms", counters[wl][short]["dram_bytes"], "B")
    if "match_mvs_screen2_kernel" in counters[wl]:
        # bench.py's "match stage" = support weights + match launch: the weights' share of that stage (launch list)
        wk = [k for k in tot if "weights_" in k]
        if wk:
            counters[wl]["match_mvs_screen2_kernel"]["other_ms_in_stage"] = tot[wk[0]] / cnt[wk[0]] / 1e6
json.dump(counters, open(os.path.join(PROF, f"{tag}_counters.json"), "w"), indent=1)
