"""``Cutout``: the container the conversion methods are bound to.

The reference's ``Cutout`` (cutout.py:61-689) also creates/downloads/prepares
NetCDF cutouts and does GIS work; none of that is on the hot path.  This class
keeps what ``convert_and_aggregate`` consumes -- ``.data`` (variables as
``(time, y, x)`` arrays with ``x, y, time, lon, lat`` coordinates), ``.grid``
(cells in y-major / x-minor order, cutout.py:355-376) -- binds the same
conversion methods (cutout.py:653-689), and adds device residency:
``to_device()`` uploads the fields once so repeated pv / wind / heat_demand
calls read them straight from HBM.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from . import convert as _convert
from .labelled import HAVE_XARRAY, Dataset, make_dataarray

if HAVE_XARRAY:  # pragma: no cover
    import xarray as xr


class _Registered:
    """Owns a cudaHostRegister'ed NumPy array (unregisters it when collected)."""

    def __init__(self, arr):
        self.arr = arr

    def __del__(self):
        try:
            import torch

            torch.cuda.cudart().cudaHostUnregister(self.arr.ctypes.data)
        except Exception:  # noqa: BLE001
            pass


def _numa_sharded_copy(a, devices, shard_bounds, local_cpus):
    """Copy ``a`` (time, y, x) into a fresh page-aligned array whose time shard r is written
    -- first touched -- by a thread pinned to the CPUs local to devices[r]."""
    import os
    import threading

    nbytes = a.nbytes
    raw = np.empty(nbytes + 4096, dtype=np.uint8)  # untouched pages: placement happens at first write
    off = (-raw.ctypes.data) % 4096
    dst = raw[off:off + nbytes].view(a.dtype).reshape(a.shape)
    nt = a.shape[0]

    def work(r, dev):
        lo, hi = shard_bounds(nt, len(devices), r)
        old = None
        try:
            cpus = local_cpus(dev)
            if cpus:
                old = os.sched_getaffinity(0)
                both = set(cpus) & set(old)
                if both:
                    os.sched_setaffinity(0, both)  # pid 0 = the calling thread
            dst[lo:hi] = a[lo:hi]
        finally:
            if old is not None:
                os.sched_setaffinity(0, old)

    ths = [threading.Thread(target=work, args=(r, d)) for r, d in enumerate(devices)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return dst


class Cutout:
    """``devices``: which GPUs of this process convert the cutout.  ``None`` = the current
    CUDA device; ``"all"`` or a list of device indices = the time axis of a HOST-resident
    cutout (NumPy / pinned / lazily loaded xarray data) is cut into one contiguous shard per
    device, every device streams and converts its shard from its own host thread, and the
    shard results land in one host array -- the reference's single call in a single process
    (convert.py:59-75), no ``torchrun`` needed.  (``time_shard=`` is the other multi-GPU mode:
    one PROCESS per GPU under ``torchrun``, results gathered over NCCL, see dist.py.)"""

    def __init__(self, path=None, data=None, time_shard=None, devices=None, **kwargs):
        if data is None:
            if path is None:
                raise ValueError("Cutout needs `data=` (or a NetCDF `path` when xarray is installed)")
            if HAVE_XARRAY:  # the reference's own way in (cutout.py:142-154): lazy, {"time": 100} chunks
                chunks = kwargs.pop("chunks", {"time": 100})
                data = xr.open_dataset(str(path), chunks=chunks)
            else:  # without xarray: this package's readers (chunk container, NetCDF-4 via h5py, NetCDF-3)
                from . import ingest

                data = ingest.open_cutout(path)
        elif isinstance(data, dict):
            data = Dataset(
                {k: v for k, v in data.items() if k not in ("time", "x", "y", "lon", "lat")},
                coords={k: data[k] for k in ("time", "x", "y", "lon", "lat") if k in data},
            )
        for c in ("x", "y", "time"):
            if c not in data.coords:
                raise ValueError(f"cutout data lacks coordinate {c!r}")
        if isinstance(data, Dataset):
            if "lon" not in data.coords:
                data.coords["lon"] = data.coords["x"]
            if "lat" not in data.coords:
                data.coords["lat"] = data.coords["y"]
        self.path = path
        self.data = data
        self._devices = devices
        # set by atlite_b200.dist.TimeShard: this process holds one contiguous
        # time shard of the cutout and results are gathered across ranks
        self.time_shard = time_shard

    @property
    def devices(self):
        """Resolved list of CUDA device indices, or None (= the current device)."""
        d = self._devices
        if d is None:
            return None
        if isinstance(d, str):
            if d != "all":
                raise ValueError(f"devices must be None, 'all' or a list of device indices, not {d!r}")
            import torch

            d = list(range(torch.cuda.device_count()))
        d = [int(k) for k in (d if hasattr(d, "__iter__") else [d])]
        if not d:
            raise RuntimeError("no CUDA device visible (atlite_b200 has no CPU fallback)")
        return d

    # ---- geometry (cutout.py:300-376)
    @property
    def coords(self):
        return self.data.coords

    @property
    def shape(self):
        return len(self.coords["y"]), len(self.coords["x"])

    @property
    def dx(self):
        x = np.asarray(self.coords["x"])
        return float(np.round(x[1] - x[0], 8)) if len(x) > 1 else 0.0

    @property
    def dy(self):
        y = np.asarray(self.coords["y"])
        return float(np.round(y[1] - y[0], 8)) if len(y) > 1 else 0.0

    @property
    def crs(self):
        return getattr(self.data, "attrs", {}).get("crs", "EPSG:4326")

    @property
    def extent(self):
        x, y = np.asarray(self.coords["x"]), np.asarray(self.coords["y"])
        return np.array([x.min() - self.dx / 2, x.max() + self.dx / 2,
                         y.min() - self.dy / 2, y.max() + self.dy / 2])

    @property
    def grid(self):
        """Cell centres in y-major, x-minor order: flat index s = iy*nx + ix."""
        x, y = np.asarray(self.coords["x"]), np.asarray(self.coords["y"])
        xs, ys = np.meshgrid(x, y)
        return pd.DataFrame({"x": xs.ravel(), "y": ys.ravel()})

    def uniform_layout(self):
        ny, nx = self.shape
        return make_dataarray(
            np.ones((ny, nx)), ("y", "x"),
            {"y": np.asarray(self.coords["y"]), "x": np.asarray(self.coords["x"])},
        )

    def indicatormatrix(self, shapes, shapes_crs=4326):
        """Cell x shape overlap matrix (cutout.py:492-515 -> gis.py:104-145): entry
        [i, j] is the fraction of grid cell j (``cutout.grid`` order) inside shape i.
        Computed on the GPU from the polygon edges (csrc/indicator.cu); shapes may be
        shapely / geopandas objects, GeoJSON dicts or coordinate arrays (see gis.py).
        Shapes must be in the cutout's CRS (no reprojection here)."""
        from . import gis

        if str(shapes_crs).upper().replace("EPSG:", "") != str(self.crs).upper().replace("EPSG:", ""):
            raise NotImplementedError(
                f"shapes_crs={shapes_crs!r}: reprojection is outside this package; "
                f"pass shapes in the cutout's CRS ({self.crs})"
            )
        return gis.compute_indicatormatrix(self.coords["x"], self.coords["y"], shapes)

    # ---- host residency
    def pin_host(self, variables=None, devices=None):
        """Return a Cutout whose (time, y, x) variables live in page-locked host memory
        (float32 NumPy arrays).  Host-streamed conversions then DMA the time slabs directly at
        PCIe rate instead of staging pageable memory through the library's pinned ring.

        With ``devices`` (default: this cutout's ``devices``) the copy is NUMA-aware: the steps
        each GPU will stream (``dist.shard_bounds``) are first touched -- hence physically
        placed -- by a thread running on the CPUs next to that GPU, so that on a two-socket box
        no shard crosses the inter-socket link on its way to its GPU."""
        import torch

        from . import _lib
        from .dist import shard_bounds

        ds = self.data
        devs = self.devices if devices is None else Cutout(data=ds, devices=devices).devices
        names = list(ds.data_vars) if variables is None else list(variables)
        out = Dataset(coords={k: np.asarray(getattr(v, "values", v)) for k, v in dict(ds.coords).items()
                              if k in ("time", "x", "y", "lon", "lat")},
                      attrs=dict(getattr(ds, "attrs", {})))
        keep = []
        for n in names:
            a = _convert._to_host(_convert._raw(ds, n))
            if not (n.startswith("solar_") and a.dtype == np.float64):
                a = np.asarray(a, dtype=np.float32)
            if devs and len(devs) > 1 and a.ndim == 3:
                dst = _numa_sharded_copy(a, devs, shard_bounds, _lib.device_local_cpus)
                rc = torch.cuda.cudart().cudaHostRegister(dst.ctypes.data, dst.nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister failed ({rc}) for variable {n!r}")
                keep.append(_Registered(dst))
                out[n] = (("time", "y", "x"), dst)
                continue
            t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
            t.numpy()[...] = a
            keep.append(t)
            out[n] = (("time", "y", "x")[-a.ndim:], t.numpy())
        res = Cutout(data=out, time_shard=self.time_shard, devices=self._devices if devices is None else devices)
        res._pinned = keep  # these objects own the page-locked memory
        return res

    # ---- device residency
    def to_device(self, device=None, variables=None, pad=True):
        """Return a Cutout whose (time, y, x) variables live in GPU memory as
        float32 torch tensors (solar position variables keep float64).  With
        ``pad`` (default) rows are zero-padded to a multiple of 4 elements when
        the width is not one already, so that the 128-bit kernels apply (the
        padding is invisible: coordinates and results keep the logical width)."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        ds = self.data
        names = list(ds.data_vars) if variables is None else list(variables)
        out = Dataset(coords={k: np.asarray(getattr(v, "values", v)) for k, v in dict(ds.coords).items()
                              if k in ("time", "x", "y", "lon", "lat")},
                      attrs=dict(getattr(ds, "attrs", {})))
        for n in names:
            arr = _convert._raw(ds, n)
            if not _convert.engine._is_torch(arr):
                a = np.asarray(arr)
                if not (n.startswith("solar_") and a.dtype == np.float64):
                    a = np.ascontiguousarray(a, dtype=np.float32)
                arr = torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)
            nx = len(out.coords["x"])
            if pad and arr.shape[-1] == nx and nx % 4:
                arr = torch.nn.functional.pad(arr, (0, 4 - nx % 4))
            dims = ("time", "y", "x")[-arr.ndim:]
            out[n] = (dims, arr.contiguous())
        return Cutout(data=out, time_shard=self.time_shard)

    def __repr__(self):
        ny, nx = self.shape
        return f"<atlite_b200.Cutout {nx} x {ny} x {len(self.coords['time'])}>"

    # ---- conversion and aggregation (cutout.py:653-689)
    convert_and_aggregate = _convert.convert_and_aggregate
    heat_demand = _convert.heat_demand
    cooling_demand = _convert.cooling_demand
    temperature = _convert.temperature
    soil_temperature = _convert.soil_temperature
    dewpoint_temperature = _convert.dewpoint_temperature
    coefficient_of_performance = _convert.coefficient_of_performance
    solar_thermal = _convert.solar_thermal
    irradiation = _convert.irradiation
    wind = _convert.wind
    pv = _convert.pv
    csp = _convert.csp
    runoff = _convert.runoff
