#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fk20 or smoke or selftest" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_main.json').readline()); print('value', d['value'], 'e2e', d['e2e']['value'], d['kernel_class_ms_per_step'])"
