#!/bin/bash
# memcheck on the small parity tests of the new kernels; ncu captures of the non-headline kernels; stage kernel at batch 128
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q \
   -k "compress_on_device or zero_poly_vs_oracle or recover_batch or fft_fr_golden or das_ext_golden or empty_lincomb or fft_g1_vs_oracle" \
   > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck.log | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fr_ntt_pass|k_das_block|k_das_level|k_g1_mul_fixed_base|k_g1_fold" -c 12 -f -o gpurun_out/prof_components \
    python tools/bench_components.py > gpurun_out/prof_components.log 2>&1; echo "ncu components rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_g1_fft_stage -s 30 -c 2 -f -o gpurun_out/prof_stage128 \
    python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline > gpurun_out/prof_stage128.log 2>&1; echo "ncu stage rc=$?"
