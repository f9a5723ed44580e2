#!/bin/bash
# usage: tools/ab_full.sh <model> <skip> <dirA> <dirB> -> gpurun_out/abfull_<model>_<tag>.ncu-rep (ncu --set full of launch #skip)
MODEL=$1; SKIP=$2; shift; shift
OUT=$PWD/gpurun_out; mkdir -p $OUT
for d in "$@"; do
  tag=$(echo $d | tr -c 'a-zA-Z0-9' '_')
  ( cd $d; X=$( grep -q no-extra bench.py && echo --no-extra )
    ncu --set full --clock-control none --import-source on -k regex:bt_k_step -s $SKIP -c 1 -f -o $OUT/abfull_${MODEL}_${tag} python bench.py --steps 40 --warmup 3 --no-cpu $X --model $MODEL > $OUT/abfull_${MODEL}_${tag}.log 2>&1; tail -1 $OUT/abfull_${MODEL}_${tag}.log )
done
