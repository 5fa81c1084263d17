#!/bin/bash
# usage: bash tools/gpu_run_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
  tail -5 gpurun_out/pytest_multi.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_g$N.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_g$N.log
grep -E '^\{|exit|Error|error' gpurun_out/bench_g$N.log | tail -5
