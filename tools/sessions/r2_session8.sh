#!/bin/bash
# Round 2, N-GPU session: parity of the sharded paths, then the bench line at N GPUs (with the component block:
# config 4 batched per GPU, config 5 offset-sharded with per-rank settings).   usage: r2_session8.sh N
N=${1:-8}
mkdir -p gpurun_out
SHARD_SCALE=12 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py > gpurun_out/sharded_${N}gpu.log 2>&1; echo "sharded rc=$?"; tail -2 gpurun_out/sharded_${N}gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-300; tail -3 gpurun_out/bench_${N}gpu.err
