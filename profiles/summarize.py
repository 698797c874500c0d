"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
usage: python profiles/summarize.py <tag>   (expects gpurun_out/<tag>_launches.csv and <tag>_*.ncu-rep)"""
import collections
import csv
import glob
import re
import subprocess
import sys

tag = sys.argv[1]
lines = open(f"gpurun_out/{tag}_launches.csv").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({len(rows)} launches, {tot:.1f} us total)\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-48s n=%4d %10.1f us %5.1f%%\n" % (k[:48], v[0], v[1], 100 * v[1] / tot))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
with open(f"profiles/{tag}_ncu_top_kernels.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on ; selected metrics per captured launch\n")
    for rep in sorted(glob.glob(f"gpurun_out/{tag}_*.ncu-rep")):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(out.splitlines()))
        hdr, units = r[0], r[1]
        idx = {h: i for i, h in enumerate(hdr)}
        f.write(f"\n== {rep}\n")
        for row in r[2:]:
            f.write("kernel: " + row[idx["Kernel Name"]][:110] + "\n")
            for w in want:
                if w in idx:
                    f.write(f"   {w:<66s} {row[idx[w]]:>16s} {units[idx[w]]}\n")
print(open(f"profiles/{tag}_launches_summary.txt").read())
