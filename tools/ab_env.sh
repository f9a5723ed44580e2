#!/bin/bash
# usage: tools/ab_env.sh <model> <rounds> "<VAR=val ... >" "<VAR=val ...>" ...  -> alternating bench runs, one per environment setting
# (e.g. "BT_B200_LIB=$PWD/brax_tracking_b200/build/libbt_base.so" "BT_STAGE=0" "")
MODEL=$1; N=$2; shift; shift
for i in $(seq $N); do
  for E in "$@"; do
    env $E python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --model $MODEL > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python -c "
import json; l=json.load(open('/tmp/ab.json')); print('$MODEL', '[$E]'.replace('$PWD/',''), round(l['value']), 'env-steps/s', round(l['ms_per_step'],4), 'ms', l['config']['warps_per_cta'], 'warps', l['config']['smem_bytes_per_cta'], 'B')"
  done
done
