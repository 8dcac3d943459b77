#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_main.json').readline()); print('value', d['value'], 'e2e', d['e2e']['value'], d['kernel_class_ms_per_step'])"
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err
python - <<'PY'
import json
for r in json.load(open('gpurun_out/components.json'))[-4:]:
    print('%-55s %10.3f ms %12.1f units/s'%(r['kernel'],r['device_ms'],r['units_per_s']))
PY
