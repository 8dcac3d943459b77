#!/bin/bash
# Tuning session: bench the product build and the variants under go_kzg_b200/lib/variants (tools/build_variant.sh).
mkdir -p gpurun_out
summ='import json,sys
d=json.loads(sys.stdin.readline()); print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["kernel_class_ms_per_step"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench.err; python -c "$summ" main < gpurun_out/bench_main.json
for f in go_kzg_b200/lib/variants/*.so; do
  n=$(basename $f .so)
  B200_KZG_LIB=$PWD/$f timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$n.json 2>> gpurun_out/bench.err; python -c "$summ" $n < gpurun_out/bench_$n.json
done
if [ -n "$TEST_VARIANT" ]; then
  B200_KZG_LIB=$PWD/go_kzg_b200/lib/variants/libb200kzg_$TEST_VARIANT.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TEST_VARIANT.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TEST_VARIANT.log
fi
