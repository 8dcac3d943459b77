#!/bin/bash
# Tuning session: bench the product build and the launch-shape variants (tools/build_variant.sh).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
summ='import json,sys
d=json.loads(sys.stdin.readline()); print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["kernel_class_ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench.err; python -c "$summ" main < gpurun_out/bench_main.json
for f in go_kzg_b200/lib/variants/*.so; do
  n=$(basename $f .so)
  B200_KZG_LIB=$PWD/$f timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$n.json 2>> gpurun_out/bench.err; python -c "$summ" $n < gpurun_out/bench_$n.json
done
for b in 96 160 192 256; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/bench_batch$b.json 2>> gpurun_out/bench.err; python -c "$summ" batch$b < gpurun_out/bench_batch$b.json
done
