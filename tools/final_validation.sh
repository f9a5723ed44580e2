#!/bin/bash
# usage: tools/final_validation.sh <tag>  -> GPU tests, smoke, bench lines of the four models (+ the reference arm), ncu launch list and
# --set full capture of the rodent step kernel, all under gpurun_out/ (one GPU)
TAG=$1
tools/run_gpu.sh $TAG rodent fly_free fly_tethered rodent_pair
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; cut -c1-300 gpurun_out/bench_${TAG}_reference.json
tools/gpu_profile.sh $TAG rodent
tools/gpu_profile.sh $TAG rodent_pair 4096
