#!/bin/bash
# usage: tools/ab_dirs.sh <model> <rounds> <dirA> <dirB> ...  -> alternating bench runs of whole copies of the repo (builds whose
# packed tables differ cannot share one model.py)
MODEL=$1; N=$2; shift; shift
for i in $(seq $N); do
  for d in "$@"; do
    ( cd $d; python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --model $MODEL > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
      python -c "
import json; l=json.load(open('/tmp/ab.json')); print('$MODEL', '$d', round(l['value']), 'env-steps/s', round(l['ms_per_step'],4), 'ms', l['config']['warps_per_cta'], 'warps')" )
  done
done
