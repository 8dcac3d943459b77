#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -k "zero_poly or recover or das" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/components.json')):
    print('%-55s %10.3f ms %12.1f units/s launches %d'%(r['kernel'],r['device_ms'],r['units_per_s'],r['launches']), r['classes_ms'])
PY
