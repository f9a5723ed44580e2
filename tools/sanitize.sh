#!/bin/bash
# compute-sanitizer over tools/race_probe.py (one mjx.forward + one wrapped step, 4 envs) for the three kernel variants.
# usage: tools/sanitize.sh <tag>    -> gpurun_out/sanitize_<tag>_<tool>_<model>.log (+ one-line summaries on stdout)
TAG=${1:-x}
mkdir -p gpurun_out
for model in rodent fly_free rodent_pair; do
  for tool in memcheck racecheck initcheck synccheck; do
    log=gpurun_out/sanitize_${TAG}_${tool}_${model}.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/race_probe.py $model > $log 2>&1
    echo "$tool $model rc=$? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) :: $(grep -c 'step ok' $log) step-ok"
  done
done
