mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_r2a.log; cat gpurun_out/pytest_r2a.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
tools/sanitize.sh r2a
