#!/bin/bash
# 2 GPUs: fused all-gather with bulk-copy peer stores: correctness + cost vs register stores
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -4
for b in 1 0; do echo "== PEER_BULK=$b"; WK_GEMM_PEER_BULK=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 4 --warmup 3 --quick --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['gather_variants_tflops'], d['gather_check'])"; done
