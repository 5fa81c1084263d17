#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -k replicated 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 4 --warmup 3 --quick > gpurun_out/bench_g2.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_g2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['how'])" || tail -30 gpurun_out/bench_g2.log
