"""Cutout ingest without xarray: lazily loaded cutouts whose chunks are decoded by the
library's parallel chunk decoder (``atl_decode_chunks``, csrc/decode.cu).

The reference writes cutouts as NetCDF-4 (HDF5) with zlib level 9 + byte shuffle
(data.py:139,245-248) and opens them lazily with ``chunks={"time": 100}``
(cutout.py:142-154); every read then inflates chunk by chunk on one thread inside libhdf5.
Here a variable is described by a CHUNK INDEX -- where each compressed chunk sits in the
file -- and the time parts ``convert_and_aggregate`` asks for are decoded on all host cores
(pread + inflate + un-shuffle + scatter), optionally straight into page-locked memory, from
where the host-streaming entry points DMA them to the GPU.

Where the index comes from:
  * ``open_netcdf4(path)``   -- an HDF5 / NetCDF-4 file, index read with h5py's low-level
                                chunk query (``dataset.id.get_chunk_info``).  h5py is not part
                                of the build image, so this path is exercised only where it is
                                installed; there is no own HDF5 metadata parser.
  * ``open_chunked(path)``   -- the container ``write_chunked`` produces: the same chunk
                                encoding (zlib + HDF5 byte shuffle), chunks back to back in one
                                file and a kerchunk-style JSON index next to it.
  * ``open_netcdf3(path)``   -- NetCDF-3 classic files through scipy's memory map (no
                                compression in that format).
All three return an ``atlite_b200.LazyDataset`` ready for ``Cutout(data=...)``.
"""

from __future__ import annotations

import ctypes as C
import json
import os
import zlib

import numpy as np
import pandas as pd

from . import _lib
from .labelled import LazyDataset


class ChunkedVariable:
    """One (time, y, x) variable stored as compressed chunks in a file."""

    def __init__(self, path, shape, chunk, dtype, offsets, sizes, origins, shuffle=True, deflate=True,
                 scale_factor=None, add_offset=None, fill_value=None, threads=0):
        self.path = os.fspath(path)
        self.shape = tuple(int(v) for v in shape)
        self.chunk = tuple(int(v) for v in chunk)
        self.dtype = np.dtype(dtype)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        self.origins = np.ascontiguousarray(origins, dtype=np.int64).reshape(-1, 3)
        if not (len(self.offsets) == len(self.sizes) == len(self.origins)):
            raise ValueError("chunk index arrays differ in length")
        if len(self.shape) != 3 or len(self.chunk) != 3:
            raise ValueError("ChunkedVariable describes (time, y, x) variables")
        self.shuffle, self.deflate = bool(shuffle), bool(deflate)
        self.scale_factor, self.add_offset, self.fill_value = scale_factor, add_offset, fill_value
        self.threads = int(threads)
        order = np.argsort(self.origins[:, 0], kind="stable")  # by first time step: parts select a contiguous run
        self.offsets, self.sizes, self.origins = self.offsets[order], self.sizes[order], self.origins[order]

    def load(self, lo, hi, pinned=False):
        """Steps [lo, hi) decoded into a float32 (nt, ny, nx) array (native byte order; packed
        integer variables are unpacked with scale_factor / add_offset, the fill value -> NaN)."""
        lo, hi = int(lo), int(hi)
        nt, (_, ny, nx) = hi - lo, self.shape
        raw_dt = self.dtype.newbyteorder("=")
        if pinned:
            import torch

            t = torch.empty((nt, ny, nx), dtype=getattr(torch, raw_dt.name), pin_memory=True)
            out = t.numpy()
        else:
            out = np.empty((nt, ny, nx), dtype=raw_dt)
        if nt > 0:
            sel = np.flatnonzero((self.origins[:, 0] < hi) & (self.origins[:, 0] + self.chunk[0] > lo))
            spec = _lib.ChunkSpec()
            spec.ny, spec.nx, spec.elem_bytes = ny, nx, self.dtype.itemsize
            spec.shuffle, spec.deflate = int(self.shuffle), int(self.deflate)
            for i in range(3):
                spec.chunk[i] = self.chunk[i]
            offs, sizes = np.ascontiguousarray(self.offsets[sel]), np.ascontiguousarray(self.sizes[sel])
            orig = np.ascontiguousarray(self.origins[sel])
            covered = self._covers(orig, lo, hi)
            if not covered:
                out[...] = 0  # chunks never written (HDF5 leaves them out): fill value below
            _lib.check(_lib.load().atl_decode_chunks(self.path.encode(), C.byref(spec), len(sel), _lib.ptr(offs),
                                                     _lib.ptr(sizes), _lib.ptr(orig), lo, nt,
                                                     out.ctypes.data_as(C.c_void_p), self.threads))
        if self.dtype.byteorder == ">" or (self.dtype.byteorder == "=" and not np.little_endian):
            out.byteswap(inplace=True)
        if out.dtype == np.float32 and self.scale_factor is None and self.add_offset is None and self.fill_value is None:
            res = out
        else:
            res = out.astype(np.float32)
            if self.fill_value is not None:
                res[out == self.fill_value] = np.nan
            if self.scale_factor is not None:
                res *= np.float32(self.scale_factor)
            if self.add_offset is not None:
                res += np.float32(self.add_offset)
        return res  # a page-locked `out` stays alive through the array's base (the tensor's storage)

    def _covers(self, orig, lo, hi):
        nty = -(-(min(hi, self.shape[0]) - (lo // self.chunk[0]) * self.chunk[0]) // self.chunk[0])
        want = nty * -(-self.shape[1] // self.chunk[1]) * -(-self.shape[2] // self.chunk[2])
        return len(orig) >= want



def _lazy(variables, coords, static, attrs, time_chunk, pinned=False):
    loaders = {n: (lambda lo, hi, v=v: v.load(lo, hi, pinned=pinned)) for n, v in variables.items()}
    return LazyDataset(loaders, coords, static=static, attrs=attrs, time_chunk=time_chunk)


# ----------------------------------------------------------------------------------------
# the chunk container (tests, benchmarks, fast re-load of a prepared cutout)
# ----------------------------------------------------------------------------------------


def shuffle_bytes(a):
    """HDF5 shuffle filter: all first bytes of the elements, then all second bytes, ..."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1, a.dtype.itemsize).T.copy().tobytes()


def write_chunked(path, data_vars, coords, chunk=(100, None, None), complevel=9, shuffle=True, static=None,
                  attrs=None):
    """Write (time, y, x) variables as zlib (+ byte shuffle) chunks, the encoding the reference
    uses for cutouts (data.py:245-248), plus ``<path>.index.json`` with the chunk index."""
    path = os.fspath(path)
    index = {"format": "atlite_b200-chunks-1", "variables": {}, "attrs": dict(attrs or {}), "static": {},
             "coords": {k: (pd.DatetimeIndex(v).as_unit("ns").asi8.tolist() if k == "time" else np.asarray(v).tolist())
                        for k, v in coords.items()}}
    with open(path, "wb") as fh:
        for name, arr in data_vars.items():
            arr = np.asarray(arr)
            nt, ny, nx = arr.shape
            c = tuple(int(min(s, n) if s else n) for s, n in zip(chunk, arr.shape))
            offs, sizes, orig = [], [], []
            for t in range(0, nt, c[0]):
                for y in range(0, ny, c[1]):
                    for x in range(0, nx, c[2]):
                        blk = np.zeros(c, dtype=arr.dtype)  # edge chunks are stored whole
                        part = arr[t:t + c[0], y:y + c[1], x:x + c[2]]
                        blk[:part.shape[0], :part.shape[1], :part.shape[2]] = part
                        raw = shuffle_bytes(blk) if shuffle else blk.tobytes()
                        comp = zlib.compress(raw, complevel) if complevel else raw
                        offs.append(fh.tell())
                        sizes.append(len(comp))
                        orig.append([t, y, x])
                        fh.write(comp)
            index["variables"][name] = dict(shape=list(arr.shape), chunk=list(c), dtype=arr.dtype.str,
                                            shuffle=bool(shuffle), deflate=bool(complevel), offsets=offs, sizes=sizes,
                                            origins=orig)
        for name, arr in (static or {}).items():
            index["static"][name] = np.asarray(arr).tolist()
    with open(path + ".index.json", "w") as fh:
        json.dump(index, fh)
    return path


def open_chunked(path, threads=0, pinned=False):
    path = os.fspath(path)
    with open(path + ".index.json") as fh:
        idx = json.load(fh)
    if idx.get("format") != "atlite_b200-chunks-1":
        raise ValueError(f"{path}.index.json is not a chunk index of this package")
    coords = {k: (pd.DatetimeIndex(np.asarray(v, dtype="int64").astype("datetime64[ns]")) if k == "time"
                  else np.asarray(v, dtype=np.float64)) for k, v in idx["coords"].items()}
    variables = {n: ChunkedVariable(path, v["shape"], v["chunk"], v["dtype"], v["offsets"], v["sizes"], v["origins"],
                                    v["shuffle"], v["deflate"], threads=threads) for n, v in idx["variables"].items()}
    static = {n: np.asarray(a, dtype=np.float32) for n, a in idx.get("static", {}).items()}
    tc = min((v.chunk[0] for v in variables.values()), default=100)
    return _lazy(variables, coords, static, idx.get("attrs", {}), tc, pinned)


# ----------------------------------------------------------------------------------------
# NetCDF-4 (HDF5) through h5py's chunk query; NetCDF-3 classic through scipy
# ----------------------------------------------------------------------------------------


def _cf_time(values, units, calendar="standard"):
    """CF 'units since epoch' -> DatetimeIndex (the encodings xarray writes for hourly data)."""
    unit, _, epoch = units.partition(" since ")
    scale = {"seconds": "s", "second": "s", "minutes": "m", "minute": "m", "hours": "h", "hour": "h",
             "days": "D", "day": "D"}[unit.strip().lower()]
    if calendar not in ("standard", "gregorian", "proleptic_gregorian"):
        raise NotImplementedError(f"calendar {calendar!r}")
    return pd.Timestamp(epoch.strip()) + pd.to_timedelta(np.asarray(values, dtype="float64"), unit=scale)


def open_netcdf4(path, threads=0, pinned=False):
    """A NetCDF-4 / HDF5 cutout as the reference writes it.  Needs h5py for the METADATA only
    (shapes, attributes, where every chunk sits); the chunk payloads are read and decoded by
    this package's native decoder."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is not in the build image
        raise ImportError(
            "reading the chunk index of a NetCDF-4 / HDF5 cutout needs h5py (metadata only); "
            "without it, open the cutout with xarray or convert it once with ingest.write_chunked") from e
    path = os.fspath(path)
    variables, static, coords, attrs = {}, {}, {}, {}
    with h5py.File(path, "r") as f:  # pragma: no cover
        attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in f.attrs.items()}

        def decode(a):
            return a.decode() if isinstance(a, bytes) else a

        for name, d in f.items():
            if not isinstance(d, h5py.Dataset):
                continue
            if name in ("x", "y", "lon", "lat"):
                coords[name] = np.asarray(d[...], dtype=np.float64)
            elif name == "time":
                coords[name] = _cf_time(d[...], decode(d.attrs["units"]), decode(d.attrs.get("calendar", "standard")))
            elif d.ndim == 3:
                if d.chunks is None:
                    raise NotImplementedError(f"variable {name!r} is stored contiguously; use xarray for this file")
                other = [k for k in (d.compression, d.scaleoffset, d.fletcher32) if k not in (None, False, "gzip")]
                if other:
                    raise NotImplementedError(f"variable {name!r} uses filters {other}; only shuffle + gzip are decoded")
                n = d.id.get_num_chunks()
                info = [d.id.get_chunk_info(i) for i in range(n)]
                variables[name] = ChunkedVariable(
                    path, d.shape, d.chunks, d.dtype, [c.byte_offset for c in info], [c.size for c in info],
                    [list(c.chunk_offset) for c in info], shuffle=bool(d.shuffle), deflate=d.compression == "gzip",
                    scale_factor=d.attrs.get("scale_factor"), add_offset=d.attrs.get("add_offset"),
                    fill_value=d.attrs.get("_FillValue"), threads=threads)
            elif d.ndim == 2:
                static[name] = np.asarray(d[...], dtype=np.float32)
    coords.setdefault("lon", coords.get("x"))
    coords.setdefault("lat", coords.get("y"))
    tc = min((v.chunk[0] for v in variables.values()), default=100)
    return _lazy(variables, coords, static, attrs, max(tc, 1), pinned)


def open_netcdf3(path):
    """NetCDF-3 classic (no compression in this format): variables are memory-mapped by
    scipy and converted to native float32 part by part."""
    from scipy.io import netcdf_file

    f = netcdf_file(os.fspath(path), "r", mmap=True)
    coords, loaders, static = {}, {}, {}
    for name, v in f.variables.items():
        if name in ("x", "y", "lon", "lat"):
            coords[name] = np.array(v[:], dtype=np.float64)
        elif name == "time":
            cal = getattr(v, "calendar", b"standard")
            coords[name] = _cf_time(np.array(v[:]), v.units.decode(), cal.decode() if isinstance(cal, bytes) else cal)
        elif len(v.shape) == 3:
            sf, ao = getattr(v, "scale_factor", None), getattr(v, "add_offset", None)

            def load(lo, hi, v=v, sf=sf, ao=ao):
                a = np.array(v[lo:hi], dtype=np.float32)
                if sf is not None:
                    a *= np.float32(sf)
                if ao is not None:
                    a += np.float32(ao)
                return a

            loaders[name] = load
        elif len(v.shape) == 2:
            static[name] = np.array(v[:], dtype=np.float32)
    coords.setdefault("lon", coords.get("x"))
    coords.setdefault("lat", coords.get("y"))
    ds = LazyDataset(loaders, coords, static=static, attrs={}, time_chunk=100)
    ds._file = f  # keeps the memory map alive
    return ds


def open_cutout(path, **kw):
    """Pick the reader by what the file is: chunk container (side-car index), HDF5 signature,
    NetCDF-3 classic magic."""
    path = os.fspath(path)
    if os.path.exists(path + ".index.json"):
        return open_chunked(path, **kw)
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        return open_netcdf4(path, **kw)
    if magic[:3] == b"CDF":
        return open_netcdf3(path)
    raise ValueError(f"{path}: neither a chunk container, an HDF5 / NetCDF-4 file nor a NetCDF-3 file")
