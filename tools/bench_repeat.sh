#!/bin/bash
# usage: tools/bench_repeat.sh <model> <repeats> [envs]  -> one line per run (same box): env-steps/s, ms per launch
MODEL=$1; N=${2:-3}; ENVS=${3:-8192}
for i in $(seq $N); do
  python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --model $MODEL --envs $ENVS > /tmp/br.json 2>/tmp/br.err || tail -3 /tmp/br.err
  python -c "
import json; l=json.load(open('/tmp/br.json')); print('$MODEL', round(l['value']), 'env-steps/s', round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value']))"
done
