#!/bin/bash
# Round 2, GPU session 2: parity suite after the stream fix + quad-cooperative MSM tails, component numbers.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msm -c 14 --csv --log-file gpurun_out/msm_launches.csv \
    python tools/msm_probe.py > gpurun_out/msm_probe.log 2>&1; echo "ncu msm rc=$?"
