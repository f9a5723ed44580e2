#!/bin/bash
# usage: tools/run_gpu.sh <tag> [models...]   -> GPU tests + one bench line per model under gpurun_out/
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log; tail -25 gpurun_out/pytest_${TAG}.log
for model in ${@:-rodent}; do
  envs=8192; [ "$model" = "rodent_pair" ] && envs=4096
  python bench.py --steps 30 --warmup 5 --model $model --envs $envs $( [ "$model" != "rodent" ] && echo "--no-cpu --no-extra" ) > gpurun_out/bench_${TAG}_${model}.json 2> gpurun_out/bench_${TAG}_${model}.err
  python - <<PY
import json
l=json.load(open("gpurun_out/bench_${TAG}_${model}.json"))
print("${model}", round(l["value"]), "env-steps/s", round(l["ms_per_step"],3), "ms e2e", round(l["e2e"]["value"]), "frac", round(l["roofline"]["frac"],5), l["config"]["warps_per_cta"], "warps")
PY
done
