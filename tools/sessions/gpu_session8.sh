#!/bin/bash
# N-GPU session: scaling line of bench.py and config 5 sharded over all GPUs.   usage: gpu_session8.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-200
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_config5.py > gpurun_out/config5_${N}gpu.json 2> gpurun_out/config5_${N}gpu.err; echo "config5 x$N rc=$?"; tail -1 gpurun_out/config5_${N}gpu.json
