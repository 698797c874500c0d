#!/bin/bash
# Round-2 ncu passes (run under gpurun on ONE B200; outputs under gpurun_out/r3, summarised by profiles/round2_summarize.py)
set -x
mkdir -p gpurun_out/r3
P="python profiles/round2_probe.py"
# (a) every launch of the second relaxation with its device time and DRAM bytes (the first relaxation = 1580 launches is warm-up)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1550 -c 1537 --csv \
    --log-file gpurun_out/r3/launches.csv $P 2 20 > gpurun_out/r3/ncu_a.log 2>&1
# (b) --set full on the top kernels (3 launches each; 2-step relaxation keeps the run short)
for k in message_bwd_v2 message_fwd_v2 gemm_tc_tma_kernel message_fwd_memo_group message_bwd_memo_state_group update_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 3 -f -o gpurun_out/r3/prof_$k $P 1 2 > gpurun_out/r3/ncu_b_$k.log 2>&1
done
ls -la gpurun_out/r3
