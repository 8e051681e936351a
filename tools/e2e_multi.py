#!/usr/bin/env python
"""End-to-end throughput of ONE process driving 1 .. N GPUs through the public API
(`Cutout(data=<pinned host arrays>, devices=...)`): the reference's single call in a single
process (convert.py:59-75), no torchrun.  The host cutout (1440 x 720 grid -> 3000 shapes,
`--steps-per-gpu` hourly steps per GPU) is pinned NUMA-aware (`Cutout.pin_host`), every
device streams its own time shard from its own host thread, and the (time, bus) result
lands in one host array.

    python tools/e2e_multi.py [--steps-per-gpu 365] [--kind pv|wind] [--reps 3]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import atlite_b200 as ab  # noqa: E402
from atlite_b200 import _lib, synthetic as syn  # noqa: E402
from atlite_b200.dist import shard_bounds  # noqa: E402

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--steps-per-gpu", type=int, default=240)
ap.add_argument("--kind", default="pv")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

NX, NY, NBUS = 1440, 720, 3000
ndev = torch.cuda.device_count()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _host_limits  # noqa: E402  (cgroup-aware host memory)

_, host_avail = _host_limits()
per_step = NX * NY * 4 * (5 if args.kind == "pv" else 2)
# the pageable source and its pinned copy both live in host memory: stay below 40 % of what the
# container may use
args.steps_per_gpu = int(min(args.steps_per_gpu, 0.4 * host_avail / (2 * per_step * ndev)))
assert args.steps_per_gpu >= 24, "not enough host memory for this measurement"
nt = args.steps_per_gpu * ndev
x, y = syn.make_coords(NX, NY, -180.0, -90.0)
tm = syn.make_time(nt + 24 * 150)[24 * 150:]
shapes = syn.make_shapes(NX, NY, NBUS)
names = ("influx_toa", "influx_direct", "influx_diffuse", "albedo", "temperature") if args.kind == "pv" else ("wnd100m", "roughness")
host = {n: np.empty((nt, NY, NX), dtype=np.float32) for n in names}
for r in range(ndev):  # generate every shard on its GPU, bring it to (pageable) host memory
    lo, hi = shard_bounds(nt, ndev, r)
    dev = torch.device("cuda", r)
    f = (syn.make_pv_fields_device(tm[lo:hi], x, y, dev, t_offset=lo + 24 * 150) if args.kind == "pv"
         else syn.make_wind_fields_device(hi - lo, NY, NX, dev, t_offset=lo))
    for n in names:
        host[n][lo:hi] = f[n].cpu().numpy()
    del f
    torch.cuda.empty_cache()
coords = dict(time=tm, x=x, y=y, lon=x, lat=y)
bpc = 4 * len(names)


def call(c):
    if args.kind == "pv":
        return c.pv("CSi", "latitude_optimal", matrix=shapes, aggregate_time=None)
    return c.wind("Vestas_V112_3MW", matrix=shapes, aggregate_time=None)


def measure(cut, steps):
    call(cut)  # warm-up: plans, operators, staging buffers
    t0 = time.perf_counter()
    for _ in range(args.reps):
        res = call(cut)
    dt = (time.perf_counter() - t0) / args.reps
    return float(NX) * NY * steps / dt, dt, res


out = {"kind": args.kind, "grid": f"{NX}x{NY}", "shapes": NBUS, "steps_per_gpu": args.steps_per_gpu, "n_gpus_visible": ndev,
       "local_cpus_per_gpu": [len(_lib.device_local_cpus(d)) for d in range(ndev)], "runs": []}
ref = None
for n in sorted({1, 2, 4, ndev} & set(range(1, ndev + 1))):
    steps = args.steps_per_gpu * n  # weak in data (every GPU streams steps_per_gpu), one process
    sub = ab.Dataset({k: v[:steps] for k, v in host.items()}, coords=dict(coords, time=tm[:steps]))
    t0 = time.perf_counter()
    cut = ab.Cutout(data=sub, devices=list(range(n))).pin_host()
    pin_s = time.perf_counter() - t0
    rate, dt, res = measure(cut, steps)
    if n == 1:
        ref = rate
        first = np.asarray(res.values).copy()
    else:  # the first GPU's shard must reproduce the 1-GPU result
        np.testing.assert_allclose(np.asarray(res.values)[:, : first.shape[1]], first, rtol=2e-5, atol=1e-6)
    out["runs"].append({"n_gpus": n, "steps": steps, "cell_ts_per_s": rate, "seconds_per_call": dt,
                        "host_GBs": rate * bpc / 1e9, "x_vs_1gpu": rate / ref, "pin_seconds": pin_s})
    del cut, sub
    _lib.release_host_staging()
print(json.dumps(out))
