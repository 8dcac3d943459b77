#!/bin/bash
# Round 2, GPU session 4: parity suite at HEAD, smem-table experiment of the stage kernel, ncu captures of the current
# kernels (MSM full set, stage kernel with/without the experiment, launch list of one bench step).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"; cut -c1-200 gpurun_out/bench_default.json
B200_KZG_LIB=$PWD/go_kzg_b200/lib/libb200kzg_smemtab.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-components > gpurun_out/bench_smemtab.json 2> gpurun_out/bench_smemtab.err; echo "bench smemtab rc=$?"; cut -c1-200 gpurun_out/bench_smemtab.json
B200_KZG_LIB=$PWD/go_kzg_b200/lib/libb200kzg_smemtab.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_fft_stage -s 14 -c 1 -f -o gpurun_out/prof_stage128_smemtab \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-components > gpurun_out/prof_stage128_smemtab.log 2>&1; echo "ncu smemtab rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm -s 7 -c 7 -f -o gpurun_out/prof_msm \
    python tools/msm_probe.py > gpurun_out/prof_msm.log 2>&1; echo "ncu msm rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'k_fk20_part2_fold2|k_g1_mul_programs|k_fr_ntt_pass|k_g1_mul_fixed_base' -c 5 -f -o gpurun_out/prof_others \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-components > gpurun_out/prof_others.log 2>&1; echo "ncu others rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-components > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
