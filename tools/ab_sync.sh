#!/bin/bash
# usage: tools/ab_sync.sh <model> <rounds> <lib> <sync1> <sync2> ...  -> alternating runs of one build with different BT_SYNC placements
MODEL=$1; N=$2; L=$3; shift; shift; shift
for i in $(seq $N); do
  for S in "$@"; do
    BT_SYNC=$S BT_B200_LIB=$PWD/$L python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --model $MODEL > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python -c "
import json; l=json.load(open('/tmp/ab.json')); print('$MODEL', 'sync=$S', round(l['value']), 'env-steps/s', round(l['ms_per_step'],4), 'ms')"
  done
done
