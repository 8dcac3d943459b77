#!/bin/bash
# One GPU session: parity tests, bench, pipe probe, ncu launch list, ncu full capture of the stage kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tools/pipe_probe > gpurun_out/pipe_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
if [ "$1" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_fft_stage -s 14 -c 2 -f -o gpurun_out/prof_stage \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/prof_stage.log 2>&1
fi
cat gpurun_out/pipe_probe.txt
