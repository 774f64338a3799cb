#!/usr/bin/env python
"""Turns what tools/gpu_profile_final.sh left in gpurun_out/ into the committed summaries under
profiles/: the default bench line, the condensed ncu launch list (durations kept), one text summary
per fully captured kernel (raw + source pages) and traffic.json.

  python tools/make_profiles.py [round_tag]      # default r1
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
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def ncu_csv(page, dst):
    with open(dst, "w") as f:
        subprocess.run(["ncu", "-i", os.path.join(OUT, f"prof_{tag}_final.ncu-rep"), "--page", page, "--csv"],
                       stdout=f, stderr=subprocess.DEVNULL, check=True)


# 1. bench line
with open(os.path.join(OUT, "bench_default.log")) as f:
    line = [l for l in f if l.startswith("{")][-1]
open(os.path.join(PROF, f"{tag}_bench_default.json"), "w").write(line)

# 2. launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(OUT, f"launches_{tag}.csv")) if l.startswith('"'))]
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
for k, v in tot.most_common():
    print(f"{k:48s} n={cnt[k]:3d} avg {v / cnt[k] / 1e6:7.3f} ms  share {100 * v / T:5.1f}%")

# 3. full captures
raw, src = os.path.join(OUT, "raw_final.csv"), os.path.join(OUT, "src_final_all.csv")
ncu_csv("raw", raw)
ncu_csv("source", src)
rr = list(csv.reader(open(raw)))
lines = open(src).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
seen, traffic = set(), {}
for k, row in enumerate(rr[2:]):
    d = dict(zip(rr[0], row))
    short = re.sub(r"<.*|\(.*", "", d["Kernel Name"]).replace("void ", "").replace("sr::", "")
    if short in seen:
        continue
    seen.add(short)
    one_raw, one_src = os.path.join(OUT, f"raw_{short}.csv"), os.path.join(OUT, f"src_{short}.csv")
    csv.writer(open(one_raw, "w")).writerows([rr[0], rr[1], row])
    sec = [i for i in range(len(starts) - 1) if short in lines[starts[i]]][0]  # first section of this kernel
    open(one_src, "w").write("\n".join(lines[starts[sec]:starts[sec + 1]]))
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), one_raw, one_src],
                         capture_output=True, text=True, check=True).stdout
    open(os.path.join(PROF, f"{tag}_final_{short}.txt"), "w").write(txt)
    traffic[short] = int((float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e9)
    print(short, d["gpu__time_duration.sum"], "ms", traffic[short], "B")
tj = os.path.join(PROF, "traffic.json")
t = json.load(open(tj))
t["cfg4"].update(traffic)
json.dump(t, open(tj, "w"), indent=1)
