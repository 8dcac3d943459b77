#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -k "fft_fr or das or zero_poly or recover or fk20 or smoke or commit" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/components.json'))[:6]:
    print('%-55s %10.3f ms %12.1f units/s %8.1f GB/s'%(r['kernel'],r['device_ms'],r['units_per_s'],r['achieved_GBs']))
PY
