#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py > gpurun_out/sharded_${N}gpu.log 2>&1; echo "sharded rc=$?"; grep sharded gpurun_out/sharded_${N}gpu.log | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_config5.py > gpurun_out/config5_${N}gpu.json 2> gpurun_out/config5_${N}gpu.err; echo "config5 x$N rc=$?"; tail -1 gpurun_out/config5_${N}gpu.json
