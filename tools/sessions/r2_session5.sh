#!/bin/bash
# Round 2, GPU session 5: parity suite at HEAD (sharded merge, batched DA), sanitizers over the round-2 kernels,
# bench with the full component block, component table incl. MSM at 2^18 / 2^20, smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
bash tools/sessions/r2_sanitizer.sh
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"; tail -2 gpurun_out/components.err
