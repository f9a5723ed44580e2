#!/bin/bash
# usage: tools/gpu_profile.sh <tag> <model> [envs]   -> gpurun_out/prof_<tag>_<model>.ncu-rep (one bt_k_step launch, --set full)
#        and (model = rodent) gpurun_out/launches_<tag>.csv, the launch list of a whole bench.py run
TAG=$1; MODEL=${2:-rodent}; ENVS=${3:-8192}
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-cpu --no-extra --model $MODEL --envs $ENVS"
python bench.py $ARGS > gpurun_out/plain_${TAG}_${MODEL}.log 2>&1 || { tail -5 gpurun_out/plain_${TAG}_${MODEL}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:bt_k_step -s 4 -c 1 -o gpurun_out/prof_${TAG}_${MODEL} python bench.py $ARGS > gpurun_out/ncu_${TAG}_${MODEL}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_${MODEL}.log
if [ "$MODEL" = "rodent" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launches_${TAG}.log 2>&1
fi
ls -la gpurun_out/prof_${TAG}_${MODEL}.ncu-rep
