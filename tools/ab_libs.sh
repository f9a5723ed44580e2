#!/bin/bash
# usage: tools/ab_libs.sh <model> <rounds> <libA> <libB> [envs] -> alternating runs of two builds of libbt_b200.so on the same box
MODEL=$1; N=$2; A=$3; B=$4; ENVS=${5:-8192}
for i in $(seq $N); do
  for L in $A $B; do
    BT_B200_LIB=$PWD/$L python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --model $MODEL --envs $ENVS > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python -c "
import json; l=json.load(open('/tmp/ab.json')); print('$MODEL', '$L'.split('/')[-1], round(l['value']), 'env-steps/s', round(l['ms_per_step'],4), 'ms')"
  done
done
