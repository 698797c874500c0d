"""Summarise the round-2 ncu outputs (gpurun_out/r3/*) into tracked files under profiles/.
  round2_launches_summary.txt   per-kernel device time + DRAM bytes of ONE relaxation (21 ensemble evaluations)
  round2_ncu_traffic.json       per kernel class: DRAM bytes per launch, per evaluation; pipe utilisations (--set full)
  round2_ncu_top_kernels.txt    selected --set full metrics per captured launch
usage: python profiles/round2_summarize.py"""
import collections, csv, glob, json, re, subprocess
D = "gpurun_out/r3"
CLASS = [("gemm_tc", "gemm_tcgen05_3xtf32"), ("message_fwd_memo", "message_fwd_memo"), ("message_bwd_memo", "message_bwd_memo"),
         ("message_fwd", "message_fwd"), ("message_bwd", "message_bwd"), ("edge_geometry", "edge_geometry"), ("row_order", "edge_geometry"),
         ("nbr_kernel", "nbr"), ("scan_kernel", "nbr"), ("readout", "readout"), ("energy_reduce", "readout"), ("ensemble", "ensemble_stats"),
         ("fire", "fire"), ("relax_finalize", "fire"), ("to_float", "fire")]
def cls(name):
    for k, c in CLASS:
        if k in name:
            return c
    return "elementwise"
lines = open(f"{D}/launches.csv").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    per[name][r["Metric Name"]] += float(r["Metric Value"])
    if r["Metric Name"] == "gpu__time_duration.sum":
        cnt[name] += 1
tot = sum(v["gpu__time_duration.sum"] for v in per.values())
evals = 21
with open("profiles/round2_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
    f.write(f"# profiles/round2_probe.py: ONE relaxation (21 ensemble evaluations) of the burnt-in bench regime, {sum(cnt.values())} launches, {tot/1e6:.1f} ms total\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write("%-52s %6s %12s %6s %14s\n" % ("kernel", "n", "time us", "share", "DRAM MB"))
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        f.write("%-52s %6d %12.1f %5.1f%% %14.1f\n" % (k[:52], cnt[k], v["gpu__time_duration.sum"] / 1e3, 100 * v["gpu__time_duration.sum"] / tot,
                                                       (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / 1e6))
classes = collections.defaultdict(lambda: {"time_us": 0.0, "dram_bytes": 0.0, "launches": 0, "kernels": set()})
for k, v in per.items():
    c = classes[cls(k)]
    c["time_us"] += v["gpu__time_duration.sum"] / 1e3; c["dram_bytes"] += v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
    c["launches"] += cnt[k]; c["kernels"].add(k)
out = {"source": "profiles/round2_launches_summary.txt (ncu launch list with DRAM counters, profiles/round2_probe.py, one relaxation = 21 evaluations)",
       "dram_bytes_per_evaluation": sum(c["dram_bytes"] for c in classes.values()) / evals,
       "classes": {k: {"dram_bytes_per_launch": c["dram_bytes"] / max(c["launches"], 1), "launches_captured": c["launches"],
                       "dram_bytes_per_evaluation": c["dram_bytes"] / evals, "time_share": c["time_us"] * 1e3 / tot, "kernels": sorted(c["kernels"])}
                   for k, c in classes.items()}}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__inst_executed_pipe_fma.sum"]
with open("profiles/round2_ncu_top_kernels.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on ; selected metrics per captured launch (profiles/round2_probe.py)\n")
    for rep in sorted(glob.glob(f"{D}/prof_*.ncu-rep")):
        res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(res.splitlines()))
        if len(r) < 3:
            continue
        hdr, units = r[0], r[1]
        idx = {h: i for i, h in enumerate(hdr)}
        f.write(f"\n== {rep}\n")
        for row in r[2:]:
            name = row[idx["Kernel Name"]]
            f.write("kernel: " + name[:110] + "\n")
            c = out["classes"].get(cls(name))
            for w in want:
                if w in idx:
                    f.write(f"   {w:<66s} {row[idx[w]]:>16s} {units[idx[w]]}\n")
            if c is not None:
                for key, met in (("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                 ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")):
                    if met in idx:
                        try:
                            c.setdefault(key + "_samples", []).append(float(row[idx[met]].replace(",", "")))
                        except ValueError:
                            pass
for c in out["classes"].values():
    for key in ("fma_pipe_pct", "tensor_pipe_pct"):
        s = c.pop(key + "_samples", None)
        if s:
            c[key] = sum(s) / len(s)
json.dump(out, open("profiles/round2_ncu_traffic.json", "w"), indent=1)
print(open("profiles/round2_launches_summary.txt").read())
print(json.dumps({k: (round(v["dram_bytes_per_evaluation"] / 1e6, 1), round(v["time_share"], 3)) for k, v in out["classes"].items()}), out["dram_bytes_per_evaluation"] / 1e9, "GB/eval")
