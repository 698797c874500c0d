#!/bin/bash
# FP64 instruction counts of the classical relax kernel on the bench's own launches (one B200, under gpurun).
# Summarised by profiles/round2_classical_fp64.py into profiles/round2_classical_fp64.json, which bench.py reads.
set -x
mkdir -p gpurun_out/r4
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
for w in gan_tersoff si_sw; do
  ncu --metrics $M --clock-control none -k regex:classical_kernel -s 30 -c 24 --csv --log-file gpurun_out/r4/classical_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r4/classical_$w.log 2>&1
done
