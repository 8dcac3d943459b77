#!/bin/bash
# 2-GPU session: component benchmark (GPU 0), sharded FK20-multi / commitment check and the 2-GPU bench line.
mkdir -p gpurun_out
timeout 900 python tools/bench_components.py > gpurun_out/components.json 2> gpurun_out/components.err; echo "components rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py > gpurun_out/sharded.log 2>&1; echo "sharded rc=$?"; tail -2 gpurun_out/sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_2gpu.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json | cut -c1-300
