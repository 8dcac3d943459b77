#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (bucket MSM incl. quad operations and shared-memory reductions,
# device decompression, verification helpers, G1 DAS extension, sharded settings), small shapes.
mkdir -p gpurun_out
SEL="lincomb_bucket_msm and (32 or 33 or 257 or 600 or 1024) or from_compressed_on_device or check_proof or das_fft_extension_over_g1 or settings_sharded or toeplitz_part2 or evaluate_poly or lane_mappings or latency_modes or concurrent"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
