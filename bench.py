#!/usr/bin/env python
"""bench.py -- grid-cell-timesteps/s on the fused PV convert+aggregate path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the largest PV configuration that fits one
GPU): synthetic ERA5 200 x 200 x 8760, cutout.pv(panel="CSi",
orientation="latitude_optimal") aggregated to 100 NUTS-like shapes.  One
"step" = one full pass of the hot path over that cutout (3.504e8
cell-timesteps, 7.0 GB of float32 input, >> L2).  At N > 1 every rank holds
its own year (weak scaling: rank r = year 2013+r of a multi-year cutout,
time-sharded as in atlite_b200.dist) and each step ends with the NCCL
all-gather that re-assembles the (time, bus) result.

`value`  : device-resident inputs, CUDA-event timed (kernel + result gather).
`e2e`    : the public API (Cutout.pv) on HOST (pinned) arrays -- H2D streaming,
           kernels, D2H of the result all inside the timed region.
`roofline`: algorithmic bytes (20 B / cell-timestep) / the
           fused kernel's CUDA-event time, against MEASURED_PEAKS.json.
`cpu_baseline`: the NumPy oracle (reference restatement) on this box's cores,
           bounded sample, rank 0 at N = 1 only.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, NT, NBUS = 200, 200, 8760, 100
X0, Y0 = 0.0, 30.0
PANEL, ORIENT = "CSi", "latitude_optimal"
BYTES_PER_CELL_TS = 20.0  # 5 float32 fields
WORKLOAD = f"synthetic ERA5 {NX}x{NY}x{NT}, cutout.pv(panel=CSi, orientation=latitude_optimal) -> {NBUS} shapes"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, every
    ~2 ms; nvidia-smi -lms 200 is too coarse for millisecond steps)."""

    def __init__(self, index=0):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.max_mhz = None

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self._stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), int(rs)))
                self._stop.wait(0.002)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def __enter__(self):
        self._t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                "sw_thermal_slowdown": 0x20}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for _, rs in self.rows for n, b in bits.items() if rs & b})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sm), **({"error": self.error} if hasattr(self, "error") else {})}


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the NumPy oracle on host cores
# ----------------------------------------------------------------------------


def oracle_pass(nt_sample, threads):
    """One bounded pass of the oracle (pv CSi/latitude_optimal -> 100 shapes) over
    200 x 200 x nt_sample, 24-step chunks on a thread pool (mirrors the
    reference's dask threaded scheduler over time chunks).  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import atlite_oracle as O

    import atlite_b200 as ab
    from atlite_b200 import synthetic as syn

    cache = oracle_pass.__dict__.setdefault("cache", {})
    if nt_sample not in cache:
        ds = syn.make_dataset(NX, NY, nt_sample, X0, Y0, kinds=("pv",), t_offset=24 * 150)
        d = {k: np.asarray(ds.raw(k)) for k in ds.keys()}
        d.update(time=ds.coords["time"], lon=ds.coords["lon"], lat=ds.coords["lat"])
        cache[nt_sample] = (d, syn.make_shapes(NX, NY, NBUS))
    d, m = cache[nt_sample]
    panel, orient = ab.get_solarpanelconfig(PANEL), O.get_orientation(ORIENT)

    def chunk(i):
        sub = {k: (v[i:i + 24] if (k == "time" or getattr(v, "ndim", 0) == 3) else v) for k, v in d.items()}
        return O.aggregate_matrix(O.convert_pv(sub, panel, orient), m)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(chunk, range(0, nt_sample, 24)))
    res = np.concatenate(parts, axis=0)
    dt = time.perf_counter() - t0
    assert res.shape == (nt_sample, NBUS)
    return dt


def cpu_baseline(target_s=12.0):
    threads = os.cpu_count() or 1
    warnings.simplefilter("ignore")
    nt0 = 24 * min(threads, 8)
    dt0 = oracle_pass(nt0, threads)
    rate0 = NX * NY * nt0 / dt0
    nt = int(np.clip(round(rate0 * target_s / (NX * NY) / 24), 1, 365)) * 24
    dt = oracle_pass(nt, threads)
    return {"value": NX * NY * nt / dt, "unit": "grid-cell-timesteps/s", "cores": threads,
            "kind": "port",
            "sample": f"NumPy float64 oracle (restatement of the reference; xarray/dask absent), "
                      f"{NX}x{NY}x{nt} steps of the same workload, 24-step chunks on {threads} threads, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warnings.simplefilter("ignore")
    threads = os.cpu_count() or 1
    nt = 24 * int(np.clip(threads, 4, 40))
    for _ in range(args.warmup):
        oracle_pass(nt, threads)
    times = [oracle_pass(nt, threads) for _ in range(args.steps)]
    total = sum(times)
    value = NX * NY * nt * args.steps / total
    line = {
        "impl": "reference", "metric": "grid-cell-timesteps/s on PV convert+aggregate",
        "value": value, "unit": "grid-cell-timesteps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"each step = {NX}x{NY}x{nt} time steps of it"},
        "cpu_baseline": {"value": value, "unit": "grid-cell-timesteps/s", "cores": threads, "kind": "port",
                         "sample": f"NumPy float64 oracle port of the reference CPU path (the reference "
                                   f"itself needs xarray/dask, absent here), {nt} of {NT} steps per step"},
        "e2e": {"value": value, "unit": "grid-cell-timesteps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------


def run_ours(args):
    import torch
    import torch.distributed as dist

    import atlite_b200 as ab
    from atlite_b200 import _lib, engine, synthetic as syn
    from atlite_b200.convert import _PvSpec
    from atlite_b200.dist import TimeShard

    warnings.simplefilter("ignore")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard = TimeShard() if world > 1 else None

    # ---- workload: this rank's year, generated on the device, mirrored to pinned host memory
    x, y = syn.make_coords(NX, NY, X0, Y0)
    time_axis = syn.make_time(NT * (rank + 1))[NT * rank:]
    fields_dev = syn.make_pv_fields_device(time_axis, x, y, dev, seed=rank)
    shapes = syn.make_shapes(NX, NY, NBUS)
    ds_dev = ab.Dataset(fields_dev, coords=dict(time=time_axis, x=x, y=y, lon=x, lat=y))
    cut_dev = ab.Cutout(data=ds_dev, time_shard=shard)

    spec = _PvSpec(ds_dev, ab.get_solarpanelconfig(PANEL), ab.get_orientation(ORIENT))
    plan = engine.get_plan(shapes, NY, NX)

    pending = []  # in-flight result gathers (N > 1): the next pass overlaps the NVLink transfer

    def step_device():
        out = spec.op.reduce(plan, spec.fields)  # memset + fused kernel on the current stream
        if shard is not None:
            out, work = shard.gather_time(out, counts=[NT] * world, async_op=True)
            pending.append(work)
        return out

    def drain():
        for w in pending:
            w.wait()  # the current stream waits for the gathers: the closing event covers them
        pending.clear()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    drain()
    sync_all()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        sync_all()
        ev0.record()
        for _ in range(args.steps):
            out = step_device()
        drain()
        ev1.record()
        sync_all()
        ms_total = ev0.elapsed_time(ev1)
        # ---- kernel-only timing (no gather), same stream, for the roofline
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        for a, b in kev:
            a.record()
            spec.op.reduce(plan, spec.fields)
            b.record()
        torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    kern_ranks = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-rank kernel times: the step is gated by the slowest GPU at every gather
        kr = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(kr, torch.tensor([kern_ms], dtype=torch.float64, device=dev))
        kern_ranks = [round(float(k.item()), 4) for k in kr]
    ms_total = float(t.item())
    cell_ts_rank = float(NX) * NY * NT
    value = cell_ts_rank * world * args.steps / (ms_total * 1e-3)
    peak, peak_src = hbm_peak()
    achieved = cell_ts_rank * BYTES_PER_CELL_TS / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get("pv_fused_200x200x8760_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass

    # ---- e2e through the public API on pinned host arrays
    host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in fields_dev.items()}
    for k, v in fields_dev.items():
        host[k].copy_(v)
    torch.cuda.synchronize()
    ds_host = ab.Dataset({k: v.numpy() for k, v in host.items()},
                         coords=dict(time=time_axis, x=x, y=y, lon=x, lat=y))
    cut_host = ab.Cutout(data=ds_host, time_shard=shard)

    def step_e2e():
        return cut_host.pv(PANEL, ORIENT, matrix=shapes, aggregate_time=None)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        res = step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = cell_ts_rank * world * e2e_steps / e2e_s
    # the two paths must agree (device-resident vs host-streamed)
    dev_res = out.float().cpu().numpy()
    api_res = np.asarray(res.values).T
    agree = float(np.max(np.abs(dev_res - api_res) / (np.abs(api_res) + 1e-3)))
    # the same call on plain (pageable) NumPy arrays: staged through the library's pinned ring
    pageable_value = None
    if world == 1 and not args.no_extra:
        ds_page = ab.Dataset({k: np.array(v.numpy()) for k, v in host.items()},
                             coords=dict(time=time_axis, x=x, y=y, lon=x, lat=y))
        cut_page = ab.Cutout(data=ds_page)
        cut_page.pv(PANEL, ORIENT, matrix=shapes, aggregate_time=None)
        t0 = time.perf_counter()
        for _ in range(2):
            cut_page.pv(PANEL, ORIENT, matrix=shapes, aggregate_time=None)
        pageable_value = cell_ts_rank * 2 / (time.perf_counter() - t0)
        del ds_page, cut_page

    line = {
        "metric": "grid-cell-timesteps/s on PV convert+aggregate",
        "value": value, "unit": "grid-cell-timesteps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu": "one year per rank, time-sharded; NCCL all-gather of (time,bus) inside the step (asynchronous: overlaps the next pass, all gathers complete inside the timed region)" if world > 1 else "single GPU",
                   "l2_policy": "inputs (7.0 GB per pass) larger than L2; no flush needed",
                   "kernel": "k_fused_reduce<PvPhys<true>> (+1 memset)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                     "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": cell_ts_rank * BYTES_PER_CELL_TS,
                     **({"kernel_ms_per_rank": kern_ranks} if kern_ranks else {})},
        "e2e": {"value": e2e_value, "unit": "grid-cell-timesteps/s",
                "h2d_bytes_per_step": int(cell_ts_rank * BYTES_PER_CELL_TS),
                "d2h_bytes_per_step": int(NT * NBUS * 4), "steps": e2e_steps,
                **({"pageable_numpy_value": pageable_value} if pageable_value else {}),
                "api": "atlite_b200.Cutout(data=<pinned host arrays>).pv('CSi','latitude_optimal',matrix=...,aggregate_time=None)",
                "max_rel_diff_vs_device_path": agree},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }

    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline()
        if not args.no_extra:
            line["extra"] = extra_measurements(torch, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_measurements(torch, dev):
    """Kernel-only roofline points at north-star spatial scale (1440 x 720 -> 3000
    shapes) on device-resident slabs; reported next to the headline, not as it."""
    import atlite_b200 as ab
    from atlite_b200 import engine, synthetic as syn
    from atlite_b200.convert import _HeatSpec, _PvSpec, _WindSpec

    peak, _ = hbm_peak()
    out = {}
    nx, ny, nt, nbus = 1440, 720, 438, 3000  # 1/20 year: 9.1 GB of PV input per pass
    x, y = syn.make_coords(nx, ny, -180.0, -90.0)
    tm = syn.make_time(nt + 24 * 170)[24 * 170:]
    shapes = syn.make_shapes(nx, ny, nbus)
    plan = engine.get_plan(shapes, ny, nx)
    out["plan_1440x720_3000"] = {k: plan.info[k] for k in ("nnz", "n_active_tiles", "n_slots", "slots_per_active_tile", "fused")}

    def timeit(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev]))

    f = syn.make_pv_fields_device(tm, x, y, dev, seed=7)
    ds = ab.Dataset(f, coords=dict(time=tm, x=x, y=y, lon=x, lat=y))
    spec = _PvSpec(ds, ab.get_solarpanelconfig(PANEL), ab.get_orientation(ORIENT))
    ms = timeit(lambda: spec.op.reduce(plan, spec.fields))
    cts = float(nx) * ny * nt
    out["pv_1440x720_slab"] = {"steps": nt, "kernel_ms": ms, "cell_ts_per_s": cts / ms * 1e3,
                               "achieved_GBs": cts * 20 / ms / 1e6, "frac_of_hbm_peak": cts * 20 / ms / 1e6 / peak}
    # wind: reuse two of the slabs as wnd100m / roughness stand-ins (values in range)
    wnd = (f["temperature"] - 255.0) * 0.5
    rough = f["albedo"] * 0.5 + 1e-3
    dsw = ab.Dataset({"wnd100m": wnd, "roughness": rough}, coords=dict(time=tm, x=x, y=y, lon=x, lat=y))
    ws = _WindSpec(dsw, ab.get_windturbineconfig("Vestas_V112_3MW"))
    ms = timeit(lambda: ws.op.reduce(plan, ws.wnd, ws.aux))
    out["wind_1440x720_slab"] = {"steps": nt, "kernel_ms": ms, "cell_ts_per_s": cts / ms * 1e3,
                                 "achieved_GBs": cts * 8 / ms / 1e6, "frac_of_hbm_peak": cts * 8 / ms / 1e6 / peak}
    nth = (nt // 24) * 24
    dsh = ab.Dataset({"temperature": f["temperature"][:nth]}, coords=dict(time=tm[:nth], x=x, y=y, lon=x, lat=y))
    hs = _HeatSpec(dsh, 15.0, 1.0, 0.0, 0.0)
    ms = timeit(lambda: hs.op.reduce(plan, hs.temp, hs.day_start))
    ctsh = float(nx) * ny * nth
    out["heat_1440x720_slab"] = {"steps": nth, "kernel_ms": ms, "cell_ts_per_s": ctsh / ms * 1e3,
                                 "achieved_GBs": ctsh * 4 / ms / 1e6, "frac_of_hbm_peak": ctsh * 4 / ms / 1e6 / peak}
    del f, ds, dsw, dsh, wnd, rough, spec, ws, hs
    out["indicatormatrix_1440x720_3000"] = indicator_measurement(x, y, nbus)
    # BASELINE configs[4]: Europe-scale 1000 x 800, per-cell pv + wind capacity factors (no shapes
    # reduction: the no-matrix branch, convert.py:200-211) -> k_cells time-sum kernels
    nx, ny, nt = 1000, 800, 240
    x, y = syn.make_coords(nx, ny, -12.0, 33.0, 0.05, 0.05)
    tm = syn.make_time(nt + 24 * 170)[24 * 170:]
    f = syn.make_pv_fields_device(tm, x, y, dev, seed=9)
    co = dict(time=tm, x=x, y=y, lon=x, lat=y)
    spec = _PvSpec(ab.Dataset(f, coords=co), ab.get_solarpanelconfig(PANEL), ab.get_orientation(ORIENT))
    ws = _WindSpec(ab.Dataset({"wnd100m": (f["temperature"] - 255.0) * 0.5, "roughness": f["albedo"] * 0.5 + 1e-3}, coords=co),
                   ab.get_windturbineconfig("Vestas_V112_3MW"))
    cts = float(nx) * ny * nt
    ms_pv = timeit(lambda: spec.cells(timesum=True))
    ms_w = timeit(lambda: ws.cells(timesum=True))
    out["percell_cf_1000x800_slab"] = {
        "steps": nt, "pv_kernel_ms": ms_pv, "wind_kernel_ms": ms_w,
        "cell_ts_per_s_combined": cts / (ms_pv + ms_w) * 1e3,
        "pv_frac_of_hbm_peak": cts * 20 / ms_pv / 1e6 / peak, "wind_frac_of_hbm_peak": cts * 8 / ms_w / 1e6 / peak,
        "combined_achieved_GBs": cts * 28 / (ms_pv + ms_w) / 1e6,
        "combined_frac_of_hbm_peak": cts * 28 / (ms_pv + ms_w) / 1e6 / peak}
    return out


def indicator_measurement(x, y, n_shapes):
    """The step in front of the path: shapes -> indicator matrix, 3000
    Voronoi regions on the 1440 x 720 grid.  Wall time of the public call (host
    packing + H2D + kernels + CSR back on the host), with the oracle's clipping
    loop timed on a few shapes beside it."""
    import time

    from atlite_b200 import gis, synthetic as syn

    rings = syn.make_voronoi_shapes(x, y, n_shapes)
    gis.compute_indicatormatrix(x, y, rings[:8])  # warm-up (context, allocator)
    t0 = time.perf_counter()
    m = gis.compute_indicatormatrix(x, y, rings)
    dt = time.perf_counter() - t0
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle"))
    import indicator_oracle as IO  # the checker, timed as the CPU stand-in for shapely's loop

    k = 6
    t0 = time.perf_counter()
    mo = IO.indicatormatrix(x, y, [gis.geometry_rings(r) for r in rings[:k]])
    dto = time.perf_counter() - t0
    err = float(abs(m[:k] - mo).max())
    return {"shapes": n_shapes, "nnz": int(m.nnz), "edges": int(sum(len(r) for r in rings)), "gpu_call_ms": dt * 1e3,
            "cpu_oracle_ms_per_shape": dto / k * 1e3, "cpu_oracle_extrapolated_ms": dto / k * n_shapes * 1e3,
            "max_abs_diff_on_sample": err, "column_sum_max_dev": float(abs(np.asarray(m.sum(0)).ravel() - 1).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the 1440x720 roofline points")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
