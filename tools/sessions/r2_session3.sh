#!/bin/bash
# Round 2, GPU session 3: parity suite (across-block lane mapping, sharded settings), bench with components, component table.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
