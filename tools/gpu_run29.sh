#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_g8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_g8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gather_variants_tflops'], d['gather_check'])" || tail -30 gpurun_out/bench_g8.log
