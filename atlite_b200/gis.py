"""Shapes -> indicator matrix on the cutout grid, computed on the GPU.

Host-side mirror of the reference's ``compute_indicatormatrix`` (gis.py:104-145)
for the case the hot path needs: ``orig`` = the cells of a regular cutout grid
(cutout.py:355-376).  I[i, j] is the fraction of cell j covered by shape i, j in
``cutout.grid`` order (iy * nx + ix).  The areas come from libatlite_b200's
``atl_indicator_compute`` (csrc/indicator.cu); there is no CPU fallback.

Shapes may be given without shapely:
  * anything with ``__geo_interface__`` (shapely geometries, geopandas rows),
  * GeoJSON geometry dicts of type Polygon / MultiPolygon (or Feature /
    GeometryCollection of those),
  * an ``(N, 2)`` array-like: one exterior ring,
  * a list of such arrays: exterior followed by holes.
Collections: list/tuple, dict (values), pandas Series, geopandas GeoSeries /
GeoDataFrame (its geometry column).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib


def _is_ring(obj):
    try:
        a = np.asarray(obj, dtype=np.float64)
    except (TypeError, ValueError):
        return False
    return a.ndim == 2 and a.shape[1] >= 2 and a.shape[0] >= 1


def _polygon_rings(coords):
    """GeoJSON Polygon coordinates -> [(ring, is_hole)], exterior first."""
    out = []
    for k, ring in enumerate(coords):
        a = np.asarray(ring, dtype=np.float64)
        if a.size == 0:
            continue
        if a.ndim != 2 or a.shape[1] < 2:
            raise ValueError("a polygon ring must be an (N, >=2) coordinate array")
        out.append((np.ascontiguousarray(a[:, :2]), k > 0))
    return out


def geometry_rings(geom):
    """One shape -> list of (ring (N,2) float64, is_hole)."""
    if geom is None:
        return []
    if hasattr(geom, "__geo_interface__"):
        geom = geom.__geo_interface__
    if isinstance(geom, dict):
        t = geom.get("type")
        if t == "Feature":
            return geometry_rings(geom.get("geometry"))
        if t == "Polygon":
            return _polygon_rings(geom["coordinates"])
        if t == "MultiPolygon":
            return [r for poly in geom["coordinates"] for r in _polygon_rings(poly)]
        if t == "GeometryCollection":
            return [r for g in geom.get("geometries", []) for r in geometry_rings(g)]
        raise ValueError(f"unsupported geometry type {t!r} (need Polygon / MultiPolygon)")
    if _is_ring(geom):
        return _polygon_rings([geom])
    if isinstance(geom, (list, tuple)) and all(_is_ring(r) for r in geom):
        return _polygon_rings(geom)
    raise TypeError(f"cannot interpret {type(geom).__name__} as a polygon")


def _iter_shapes(shapes):
    if hasattr(shapes, "columns") and hasattr(shapes, "geometry"):  # GeoDataFrame
        shapes = shapes.geometry
    if isinstance(shapes, dict):
        return list(shapes.values())
    if hasattr(shapes, "values") and not isinstance(shapes, np.ndarray):  # pandas / geopandas Series
        return list(shapes.values)
    return list(shapes)


def pack_shapes(shapes):
    """Collection of shapes -> (shape_ring_ptr, ring_ptr, ring_is_hole, xy) in the
    layout ``atl_indicator_compute`` takes (include/atlite_b200.h)."""
    shape_ring_ptr, ring_ptr, holes, chunks = [0], [0], [], []
    for g in _iter_shapes(shapes):
        for ring, hole in geometry_rings(g):
            chunks.append(ring)
            ring_ptr.append(ring_ptr[-1] + len(ring))
            holes.append(1 if hole else 0)
        shape_ring_ptr.append(len(holes))
    xy = np.ascontiguousarray(np.concatenate(chunks) if chunks else np.zeros((0, 2)), dtype=np.float64)
    return (np.asarray(shape_ring_ptr, dtype=np.int64), np.asarray(ring_ptr, dtype=np.int64),
            np.asarray(holes, dtype=np.int8), xy)


def regular_axis(c, name):
    """(first centre, step) of an ascending, evenly spaced coordinate."""
    c = np.asarray(c, dtype=np.float64)
    if c.ndim != 1 or len(c) < 1:
        raise ValueError(f"coordinate {name!r} must be 1-D and non-empty")
    if len(c) == 1:
        raise ValueError(f"coordinate {name!r} has a single value: the cell size is undefined")
    d = np.diff(c)
    step = (c[-1] - c[0]) / (len(c) - 1)
    if not step > 0 or not np.allclose(d, step, rtol=1e-6, atol=1e-9 * abs(step)):
        raise ValueError(f"coordinate {name!r} must be ascending and evenly spaced")
    return float(c[0]), float(step)


def compute_indicatormatrix(x, y, shapes, device=None):
    """CSR (n_shapes, len(y) * len(x)) of covered cell fractions for the regular
    grid with cell centres ``x`` (columns) and ``y`` (rows)."""
    from .engine import current_device

    lib = _lib.load()
    x0, dx = regular_axis(x, "x")
    y0, dy = regular_axis(y, "y")
    nx, ny = len(x), len(y)
    srp, rp, holes, xy = pack_shapes(shapes)
    n_shapes = len(srp) - 1
    h = C.c_void_p()
    device = current_device() if device is None else device
    _lib.check(lib.atl_indicator_compute(device, ny, nx, x0, dx, y0, dy, n_shapes, _lib.ptr(srp), _lib.ptr(rp),
                                         _lib.ptr(holes), _lib.ptr(xy), C.byref(h)))
    try:
        nnz = C.c_int64()
        _lib.check(lib.atl_indicator_nnz(h, C.byref(nnz)))
        indptr = np.empty(n_shapes + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float64)
        _lib.check(lib.atl_indicator_export(h, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data)))
    finally:
        lib.atl_indicator_destroy(h)
    return sp.csr_matrix((data, indices, indptr), shape=(n_shapes, ny * nx))
