#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
summ='import json,sys
d=json.loads(sys.stdin.readline()); print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["kernel_class_ms_per_step"])'
for b in 100 128 148; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/bench_batch$b.json 2>> gpurun_out/bench.err; python -c "$summ" batch$b < gpurun_out/bench_batch$b.json
done
