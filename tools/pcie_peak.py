"""Host<->device copy ceiling of this box, the denominator of the multi-GPU `e2e` number (bench.py): every rank copies
pinned host memory to its GPU and back, first one direction at a time, then both at once (two streams), all ranks at the
same time.  Reports per-GPU and aggregate GB/s.

    python tools/pcie_peak.py                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_peak.py

torch is plumbing here (pinned buffers, streams, events, the rendezvous); no kernel of the product is involved."""
import json
import os
import sys

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))

GIB = 1 << 30
nbytes = int(float(os.environ.get("WK_PCIE_GIB", "1")) * GIB)
reps = 8
host_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).fill_(3)
host_dn = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
dev_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
dev_b = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()


def run(up, dn):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_up.wait_event(e0)
    s_dn.wait_event(e0)
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s_up):
                dev_a.copy_(host_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s_dn):
                host_dn.copy_(dev_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)  # everybody is done when the slowest is
    return float(t.item())


run(True, True)  # warm-up
out = {"n_gpus": world, "bytes_per_copy": nbytes, "reps": reps}
for name, up, dn in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    ms = run(up, dn)
    per_dir = nbytes * reps / (ms * 1e-3) / 1e9
    out[name] = {"ms": ms, "per_gpu_gbs_each_direction": per_dir, "aggregate_gbs": per_dir * world * (2 if up and dn else 1)}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
