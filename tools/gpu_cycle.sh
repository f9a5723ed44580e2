#!/bin/bash
# Standard GPU iteration: parity tests, bench line, ncu launch list + full capture of the step kernel.
# usage: tools/gpu_cycle.sh <tag> [skip_tests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_${TAG}.log; cat gpurun_out/pytest_${TAG}.log
fi
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bt_k_step -s 4 -c 1 -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
