"""Turns the raw files tools/profile_round.sh brings back (gpurun_out/<tag>_launches.csv, <tag>_full_raw.csv) into the
tracked summaries under profiles/: <tag>_launches_short60.csv, <tag>_launch_summary.csv, <tag>_ncu_full.csv.
usage: python tools/profiles_from_ncu.py r2"""
import collections
import csv
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")

lines = [l for l in open(os.path.join(src, tag + "_launches.csv")) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
with open(os.path.join(dst, tag + "_launch_summary.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, first 60 sweeps of the cfg-2 sequence through loam_process_sweep with "
            "registered + surround outputs on (tools/short_run.py 60; cold-cache, serialised: compare shares)\n")
    f.write(f"# total {tot:.0f} us = {tot / 60:.1f} us per sweep, {n} launches = {n / 60:.1f} per sweep\n")
    f.write("kernel,launches,total_us,avg_us,share_pct\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{100 * v[1] / tot:.1f}\n")
shutil.copy(os.path.join(src, tag + "_launches.csv"), os.path.join(dst, tag + "_launches_short60.csv"))

r = list(csv.reader(open(os.path.join(src, tag + "_full_raw.csv"))))
h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
ix = [h.index(w) for w in want if w in h]
seen = {}
with open(os.path.join(dst, tag + "_ncu_full.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on (tools/profile_round.sh), cfg-2 sequence with registered + surround on "
            "(tools/short_run.py 30), up to two launches per kernel; cold-cache per-launch values\n")
    w = csv.writer(f)
    w.writerow([h[i] for i in ix])
    w.writerow([r[1][i] for i in ix])
    for row in r[2:]:
        name = row[h.index("Kernel Name")]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] <= 2:
            w.writerow([row[i] for i in ix])
print(f"{tag}: {n} launches, {tot / 60:.1f} us of kernels per sweep, {len(seen)} kernels in the full capture")
