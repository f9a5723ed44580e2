#!/bin/bash
# usage: tools/l1_probe.sh <model> <warps...>  -> L1 hit / miss sectors of the global (table) loads and the warp-state breakdown of one
# step launch per BT_WARPS setting (12 warps per CTA = 196 KB shared-memory carve-out = 60 KB L1 instead of 22 KB at 14)
MODEL=$1; shift
OUT=$PWD/gpurun_out; mkdir -p $OUT
for w in "$@"; do
  BT_WARPS=$w ncu --section WarpStateStats --section LaunchStats \
      --metrics l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,gpu__time_duration.sum,smsp__inst_executed.sum \
      --clock-control none -k regex:bt_k_step -s 8 -c 1 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --model $MODEL > $OUT/l1probe_${MODEL}_w$w.txt 2>&1
  echo "== warps $w"; grep -E "hit_rate|lookup_miss|lookup_hit|long_scoreboard|latency_per_inst|time_duration|inst_executed.sum|Shared Memory Configuration|Stall Long|Warp Cycles Per Issued" $OUT/l1probe_${MODEL}_w$w.txt
done
