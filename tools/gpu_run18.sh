#!/bin/bash
# 2 GPUs: bench with the shared-upload e2e
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_g2.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_g2.log | cut -c1-2500
