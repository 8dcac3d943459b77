#!/bin/bash
# Round 2, GPU session 1: full parity suite (old + new shapes), bench at three table widths, component numbers,
# ncu full captures (stage kernel at the benchmarked batch, MSM kernels), launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in 8 10 12; do
  B200_FB_WINDOW=$w timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; echo "bench w=$w rc=$?"
  cut -c1-400 gpurun_out/bench_w$w.json
done
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_fft_stage -s 14 -c 2 -f -o gpurun_out/prof_stage128 \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline > gpurun_out/prof_stage128.log 2>&1; echo "ncu stage rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm -c 14 -f -o gpurun_out/prof_msm \
    python tools/msm_probe.py > gpurun_out/prof_msm.log 2>&1; echo "ncu msm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
