"""Summarise profiles/round2_classical_ncu.sh: executed FP64 flops per relaxed proposal of the classical relax kernel.

usage: python profiles/round2_classical_fp64.py gpurun_out/r4 > profiles/round2_classical_fp64.json"""
import csv
import json
import sys
from pathlib import Path

src = Path(sys.argv[1])
out = {}
for w in ("gan_tersoff", "si_sw"):
    rows = []
    with open(src / f"classical_{w}.csv") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        k = r["ID"]
        per.setdefault(k, {"grid": r["Grid Size"], "name": r["Kernel Name"]})
        per[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    tot_flop = tot_prop = tot_ns = tot_inst = 0.0
    for k, d in per.items():
        chains = int(d["grid"].strip("()").split(",")[0])
        flop = 2 * d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] \
            + d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
        tot_flop += flop
        tot_prop += chains
        tot_ns += d["gpu__time_duration.sum"]
        tot_inst += d["smsp__inst_executed.sum"]
    out[w] = {"launches_captured": len(per), "proposals_captured": int(tot_prop),
              "fp64_flop_per_proposal": tot_flop / tot_prop, "warp_instructions_per_proposal": tot_inst / tot_prop,
              "fp64_tflops_under_ncu": tot_flop / (tot_ns * 1e-9) / 1e12,
              "source": "ncu smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum on bench.py's own launches "
                        "(profiles/round2_classical_ncu.sh); flop = 2*dfma + dmul + dadd"}
print(json.dumps(out, indent=1))
