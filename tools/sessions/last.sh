#!/bin/bash
# last session of the round: bench sanity after the final edits + launch list of the final code
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sanity.json 2> gpurun_out/bench_sanity.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_sanity.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
