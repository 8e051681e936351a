#!/usr/bin/env python
"""torchrun probe: cost of the (time, bus) result all-gather by itself and of the
bench step (kernel + gather), per step, for several step counts."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
loc = torch.randn(8760, 100, device=dev)
out = torch.empty(8760 * world, 100, device=dev)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {}
for _ in range(3):
    dist.all_gather_into_tensor(out, loc)
for n in (1, 5, 20):
    res[f"gather_only_ms_n{n}"] = timed(lambda: dist.all_gather_into_tensor(out, loc), n)
# a ~1 ms memory-bound kernel in front of the gather, like the bench step
for n in (5, 20):
    res[f"fill+gather_ms_n{n}"] = timed(lambda: (big.fill_(1), dist.all_gather_into_tensor(out, loc)), n)
    res[f"fill_only_ms_n{n}"] = timed(lambda: big.fill_(1), n)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
