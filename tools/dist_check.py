#!/usr/bin/env python
"""Multi-GPU check (run under torchrun): every rank holds one time shard of a
small cutout; the gathered (bus, time) results of pv / wind / heat_demand and the
all-reduced per-cell means must equal the unsharded computation."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import atlite_b200 as ab  # noqa: E402
from atlite_b200 import synthetic as syn  # noqa: E402
from atlite_b200.dist import TimeShard, shard_bounds  # noqa: E402

warnings.simplefilter("ignore")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
nx, ny, nbus = 64, 40, 9
nt = 24 * int(os.environ.get("DIST_CHECK_DAYS", "7"))  # fewer day blocks than ranks -> EMPTY shards (must work)
full = syn.make_dataset(nx, ny, nt, x0=0.0, y0=30.0)
m = syn.make_shapes(nx, ny, nbus)
lo, hi = shard_bounds(nt, world, rank, align=24)
mine = syn.make_dataset(nx, ny, hi - lo, x0=0.0, y0=30.0, t_offset=lo)
for k in full.keys():
    assert np.array_equal(full.raw(k)[lo:hi], mine.raw(k)), k
sharded = ab.Cutout(data=mine, time_shard=TimeShard())
whole = ab.Cutout(data=full)
ok = True
for name, call in (
    ("pv", lambda c: c.pv("CSi", "latitude_optimal", matrix=m, aggregate_time=None)),
    ("wind", lambda c: c.wind("Vestas_V112_3MW", matrix=m, aggregate_time=None)),
    ("heat", lambda c: c.heat_demand(matrix=m, aggregate_time=None, hour_shift=0.0)),
    ("pv_mean_cells", lambda c: c.pv("CSi", "latitude_optimal", aggregate_time="mean")),
    ("wind_device", lambda c: c.to_device().wind("Vestas_V112_3MW", matrix=m, aggregate_time="sum")),
):
    a, b = call(sharded), call(whole)
    err = float(np.max(np.abs(np.asarray(a.values) - np.asarray(b.values)) / (np.abs(np.asarray(b.values)) + 1e-3)))
    same = np.asarray(a.values).shape == np.asarray(b.values).shape and err < 5e-5
    ok &= same
    if rank == 0:
        print(f"{name}: shape {np.asarray(a.values).shape} max rel diff {err:.2e} {'OK' if same else 'MISMATCH'}")
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", f"world={world}")
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
