#!/bin/bash
# 8 GPUs: register peer stores vs bulk-copy peer stores, then the default bench line
mkdir -p gpurun_out
for b in 0 1; do echo "== PEER_BULK=$b"; WK_GEMM_PEER_BULK=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$b bench.py --gpus 8 --steps 4 --warmup 3 --quick --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['gather_variants_tflops'], d['gather_check'])"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_g8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_g8.log | cut -c1-900
