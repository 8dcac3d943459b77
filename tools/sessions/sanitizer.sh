#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "fft_fr_golden or das_ext_golden or fft_fr_vs_oracle or kzg_proof" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck.log | tail -3
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "compress_on_device or zero_poly_vs_oracle or kzg_proof or blob_to or sharded" > gpurun_out/initcheck.log 2>&1; echo "initcheck rc=$?" | tee -a gpurun_out/initcheck.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/initcheck.log | tail -3
