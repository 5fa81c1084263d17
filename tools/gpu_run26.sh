#!/bin/bash
# 8 GPUs: host topology + e2e with / without NUMA-local pinned buffers
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^CPU\(s\)"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; free -g | head -2) > gpurun_out/topo_g8.txt 2>&1
for n in 1 0; do echo "== WK_NUMA=$n"; WK_DEBUG=1 WK_NUMA=$n timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus 8 --steps 4 --warmup 3 --quick > gpurun_out/bench_g8_numa$n.log 2>&1; grep "host memory placement" gpurun_out/bench_g8_numa$n.log | head -8; tail -1 gpurun_out/bench_g8_numa$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('host_memory'))"; done
