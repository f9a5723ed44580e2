#!/bin/bash
# usage: tools/ab_metrics.sh <model> <dirA> <dirB>  -> timing + ncu sections of one step launch for two builds of the repo
MODEL=$1; shift
OUT=$PWD/gpurun_out; mkdir -p $OUT
for d in "$@"; do
  tag=$(echo $d | tr -c 'a-zA-Z0-9' '_')
  ( cd $d; X=$( grep -q no-extra bench.py && echo --no-extra )
    python bench.py --steps 30 --warmup 5 --no-cpu $X --model $MODEL > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python -c "
import json; l=json.load(open('/tmp/ab.json')); print('$d', '$MODEL', l['config']['warps_per_cta'], 'warps', round(l['value']), 'env-steps/s', round(l['ms_per_step'],3), 'ms')"
    ncu --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section InstructionStats --section LaunchStats --section Occupancy --section SpeedOfLight \
        --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct \
        --clock-control none -k regex:bt_k_step -s 30 -c 1 python bench.py --steps 40 --warmup 3 --no-cpu $X --model $MODEL > $OUT/ab_${MODEL}_${tag}.txt 2>&1 )
done
