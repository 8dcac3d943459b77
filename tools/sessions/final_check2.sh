#!/bin/bash
bash tools/sessions/final_check.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fk20_part2_fold2|k_g1_mul_fixed_base|k_g1_mul_programs" -c 4 -f -o gpurun_out/prof_g1_misc \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/prof_g1_misc.log 2>&1; echo "ncu misc rc=$?"
