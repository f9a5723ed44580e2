#!/bin/bash
# usage: tools/warps_sweep.sh <model> <w1> <w2> ...   -> env-steps/s per BT_WARPS setting (30 timed steps each)
MODEL=$1; shift
for w in "$@"; do
  BT_WARPS=$w python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --model $MODEL > /tmp/ws.json 2>/tmp/ws.err || tail -3 /tmp/ws.err
  python -c "
import json; l=json.load(open('/tmp/ws.json')); print('$MODEL warps', l['config']['warps_per_cta'], round(l['value']), 'env-steps/s', round(l['ms_per_step'],3), 'ms')"
done
