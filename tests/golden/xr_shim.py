"""Minimal stand-ins for ``xarray`` / ``dask`` / ``geopandas`` so that the
REFERENCE'S OWN hot-path modules (atlite/convert.py, aggregate.py, wind.py,
pv/*.py, resource.py, utils.py) can be executed, unmodified, in a container
where those third-party packages are not installed.

TEST INFRASTRUCTURE ONLY -- used by ``tests/golden/make_golden.py`` to produce
the golden vectors that pin ``oracle/atlite_oracle.py``.  Nothing in the
product imports this.

What is emulated is only third-party *container* behaviour the path relies on
(documented xarray semantics): named-dimension broadcasting (dims ordered by
first appearance), ``where`` / ``fillna`` / ``clip`` (NaN-preserving),
NaN-skipping ``sum`` / ``mean``, ``stack`` / ``transpose`` / ``expand_dims``,
``resample(time="1D").mean()`` as left-closed calendar-day bins, the ``.dt``
accessor, ``apply_ufunc`` for element-wise functions, ``Coordinates``.  All
arithmetic is the reference's own source executing on NumPy arrays (dask.array
ufuncs map to the NumPy ufuncs dask itself dispatches to for NumPy inputs).
"""

from __future__ import annotations

import importlib.util
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference/atlite"


def _unwrap(x):
    return x.values if isinstance(x, DataArray) else x


def _unify(args):
    """Broadcast DataArrays by dim name (order of first appearance); raw ndarrays
    align positionally (trailing) with the result, like xarray does."""
    dims = []
    for a in args:
        if isinstance(a, DataArray):
            for d in a.dims:
                if d not in dims:
                    dims.append(d)
    vals, coords = [], {}
    for a in args:
        if isinstance(a, DataArray):
            perm = [a.dims.index(d) for d in dims if d in a.dims]
            v = a.values.transpose(perm) if perm else a.values
            shape = [a.values.shape[a.dims.index(d)] if d in a.dims else 1 for d in dims]
            vals.append(v.reshape(shape))
            for k, c in a.coords.items():
                coords.setdefault(k, c)
        else:
            vals.append(a)
    return dims, vals, coords


class DataArray:
    __array_priority__ = 50

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            coords = coords if coords is not None else data.coords
            dims = dims if dims is not None else data.dims
            data = data.values
        self.values = np.asarray(data)
        if isinstance(coords, Coordinates):
            cdict = dict(coords._c)
            if dims is None:
                dims = tuple(coords.dims)
        elif isinstance(coords, (list, tuple)):
            cdict = {}
            dims = dims or tuple(c.name for c in coords)
            for d, c in zip(dims, coords):
                cdict[d] = c
        else:
            cdict = dict(coords or {})
            if dims is None:
                dims = tuple(cdict)[: self.values.ndim] if cdict else ()
        dims = (dims,) if isinstance(dims, str) else tuple(dims)
        assert len(dims) == self.values.ndim, (dims, self.values.shape)
        self.dims = dims
        self._coords = {}
        for k, v in cdict.items():
            if isinstance(v, DataArray):
                v = v.values
            if k in dims or k in ("lon", "lat"):
                self._coords[k] = v if isinstance(v, pd.Index) else np.asarray(v)
        self.name = name
        self.attrs = dict(attrs or {})

    # ---- structure
    @property
    def coords(self):
        return Coordinates(self._coords, [d for d in self.dims if d in self._coords], owner=self)

    data = property(lambda s: s.values)
    shape = property(lambda s: s.values.shape)
    ndim = property(lambda s: s.values.ndim)
    dtype = property(lambda s: s.values.dtype)
    sizes = property(lambda s: dict(zip(s.dims, s.values.shape)))

    @property
    def indexes(self):
        return {d: pd.Index(self._coords[d], name=d) for d in self.dims if d in self._coords}

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def __bool__(self):
        return bool(self.values)

    def __float__(self):
        return float(self.values)

    def _new(self, values, dims=None, coords=None, keep_name=True):
        dims = self.dims if dims is None else tuple(dims)
        c = self._coords if coords is None else coords
        c = {k: v for k, v in c.items() if k in dims}
        return DataArray(values, c, dims, self.name if keep_name else None)

    def rename(self, name):
        out = self._new(self.values)
        out.name = name
        out.attrs = dict(self.attrs)
        return out

    def load(self, **kw):
        return self

    def chunk(self, *a, **k):
        return self

    def assign_attrs(self, *args, **kw):
        out = self._new(self.values)
        out.attrs = dict(self.attrs)
        for a in args:
            out.attrs.update(a)
        out.attrs.update(kw)
        return out

    def copy(self):
        out = self._new(self.values.copy())
        out.attrs = dict(self.attrs)
        return out

    # ---- numpy protocol: ufuncs broadcast by name
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        dims, vals, coords = _unify(inputs)
        with np.errstate(all="ignore"):
            res = ufunc(*vals, **kwargs)
        return DataArray(res, {k: v for k, v in coords.items() if k in dims}, dims, self.name)

    def _bin(self, other, fn, refl=False):
        if isinstance(other, pd.Timedelta):
            other = other.to_timedelta64()
        args = (other, self) if refl else (self, other)
        dims, vals, coords = _unify(args)
        with np.errstate(all="ignore"):
            res = fn(*vals)
        return DataArray(res, {k: v for k, v in coords.items() if k in dims}, dims, self.name)

    def __add__(s, o): return s._bin(o, np.add)
    def __radd__(s, o): return s._bin(o, np.add, True)
    def __sub__(s, o): return s._bin(o, np.subtract)
    def __rsub__(s, o): return s._bin(o, np.subtract, True)
    def __mul__(s, o): return s._bin(o, np.multiply)
    def __rmul__(s, o): return s._bin(o, np.multiply, True)
    def __truediv__(s, o): return s._bin(o, np.true_divide)
    def __itruediv__(s, o): return s._bin(o, np.true_divide)
    def __rtruediv__(s, o): return s._bin(o, np.true_divide, True)
    def __pow__(s, o): return s._bin(o, np.power)
    def __rpow__(s, o): return s._bin(o, np.power, True)
    def __mod__(s, o): return s._bin(o, np.mod)
    def __neg__(s): return s._new(-s.values)
    def __abs__(s): return s._new(np.abs(s.values))
    def __invert__(s): return s._new(~s.values)
    def __and__(s, o): return s._bin(o, np.logical_and)
    def __rand__(s, o): return s._bin(o, np.logical_and, True)
    def __or__(s, o): return s._bin(o, np.logical_or)
    def __ror__(s, o): return s._bin(o, np.logical_or, True)
    def __lt__(s, o): return s._bin(o, np.less)
    def __le__(s, o): return s._bin(o, np.less_equal)
    def __gt__(s, o): return s._bin(o, np.greater)
    def __ge__(s, o): return s._bin(o, np.greater_equal)
    def __eq__(s, o): return s._bin(o, np.equal)
    def __ne__(s, o): return s._bin(o, np.not_equal)
    __hash__ = None

    # ---- xarray methods used on the path
    def where(self, cond, other=np.nan):
        if callable(cond):
            cond = cond(self)
        dims, vals, coords = _unify((self, cond, other))
        res = np.where(vals[1], vals[0], vals[2])
        return DataArray(res, {k: v for k, v in coords.items() if k in dims}, dims, self.name)

    def fillna(self, value):
        return self._new(np.where(np.isnan(self.values), value, self.values))

    def clip(self, min=None, max=None):
        dims, vals, coords = _unify((self, min, max))
        v = vals[0]
        with np.errstate(all="ignore"):
            if min is not None:
                v = np.where(np.isnan(v), v, np.maximum(v, vals[1]))
            if max is not None:
                v = np.where(np.isnan(v), v, np.minimum(v, vals[2]))
        if v.dtype != self.values.dtype and min is not None and not isinstance(min, DataArray) \
                and max is None and np.isscalar(min):
            v = v.astype(self.values.dtype)  # weak python scalars keep the array dtype
        return DataArray(v, {k: c for k, c in coords.items() if k in dims}, dims, self.name)

    def notnull(self):
        return self._new(~np.isnan(self.values))

    def any(self):
        return bool(np.any(self.values))

    def all(self):
        return bool(np.all(self.values))

    def _reduce(self, fn, dim, keep_attrs):
        if dim is None:
            return DataArray(fn(self.values), {}, ())
        ax = self.dims.index(dim)
        out = self._new(fn(self.values, axis=ax), [d for d in self.dims if d != dim])
        if keep_attrs:
            out.attrs = dict(self.attrs)
        return out

    def sum(self, dim=None, keep_attrs=False):
        return self._reduce(np.nansum, dim, keep_attrs)

    def mean(self, dim=None, keep_attrs=False):
        return self._reduce(np.nanmean, dim, keep_attrs)

    def transpose(self, *dims):
        dims = tuple(dims) if dims else self.dims[::-1]
        return self._new(self.values.transpose([self.dims.index(d) for d in dims]), dims)

    def expand_dims(self, name):
        return DataArray(self.values[None], self._coords, (name,) + self.dims, self.name)

    def stack(self, **kw):
        (new, old), = kw.items()
        old = list(old)
        rest = [d for d in self.dims if d not in old]
        v = self.transpose(*rest, *old).values
        v = v.reshape(v.shape[: len(rest)] + (-1,))
        c = {k: self._coords[k] for k in rest if k in self._coords}
        c[new] = pd.MultiIndex.from_product([self._coords[d] for d in old], names=old)
        return DataArray(v, c, tuple(rest) + (new,), self.name)

    def reindex_like(self, other):
        out = self
        for d in self.dims:
            tgt = np.asarray(other.coords[d].values if isinstance(other.coords[d], DataArray) else other.coords[d])
            idx = pd.Index(self._coords[d]).get_indexer(tgt)
            v = np.take(out.values, np.clip(idx, 0, None), axis=out.dims.index(d)).astype(float)
            sl = [slice(None)] * v.ndim
            sl[out.dims.index(d)] = idx < 0
            v[tuple(sl)] = np.nan
            c = dict(out._coords)
            c[d] = tgt
            out = DataArray(v, c, out.dims, out.name)
        return out

    def assign_coords(self, coords=None, **kw):
        c = dict(self._coords)
        for k, v in {**(coords._c if isinstance(coords, Coordinates) else (coords or {})), **kw}.items():
            c[k] = v.values if isinstance(v, DataArray) else v
        out = DataArray(self.values, c, self.dims, self.name)
        out.attrs = dict(self.attrs)
        return out

    @property
    def dt(self):
        idx = pd.DatetimeIndex(self.values)
        return types.SimpleNamespace(
            hour=DataArray(np.asarray(idx.hour), self._coords, self.dims),
            minute=DataArray(np.asarray(idx.minute), self._coords, self.dims),
        )

    def resample(self, time=None):
        assert time == "1D"
        return _DailyResample(self)

    def interp(self, **indexers):
        """xarray's advanced (pointwise) linear interpolation: DataArray indexers that
        share dims -> scipy.interpolate.interpn(linear, bounds_error=False, fill_value=nan),
        which is what xarray delegates to for n-d linear interpolation."""
        from scipy.interpolate import interpn

        names = list(indexers)
        assert set(names) == set(self.dims), "shim interp: all dims must be interpolated"
        table = self.transpose(*names)
        targets = [indexers[n] for n in names]
        dims, vals, coords = _unify(targets)
        shape = np.broadcast_shapes(*[np.shape(v) for v in vals])
        xi = np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), shape).ravel() for v in vals], axis=-1)
        pts = tuple(np.asarray(table._coords[n], dtype=np.float64) for n in names)
        out = interpn(pts, np.asarray(table.values, dtype=np.float64), xi, method="linear",
                      bounds_error=False, fill_value=np.nan).reshape(shape)
        return DataArray(out, {k: v for k, v in coords.items() if k in dims}, dims, self.name)

    def isel(self, **kw):
        out = self
        for d, i in kw.items():
            ax = out.dims.index(d)
            v = out.values[(slice(None),) * ax + (i,)]
            c = dict(out._coords)
            if d in c:
                c[d] = np.asarray(c[d])[i]
            out = DataArray(v, c, out.dims, out.name)
        return out


class _DailyResample:
    """xarray ``resample(time="1D")``: left-closed calendar-day bins from the first
    to the last day, labelled by the day start; ``mean`` skips NaN."""

    def __init__(self, da):
        self.da = da

    def mean(self, dim="time"):
        da = self.da
        t = pd.DatetimeIndex(da._coords["time"])
        days = t.floor("D")
        labels = pd.date_range(days[0], days[-1], freq="D")
        ax = da.dims.index("time")
        outs = []
        for lab in labels:
            sel = np.compress(np.asarray(days == lab), da.values, axis=ax)
            with np.errstate(all="ignore"):
                import warnings

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    outs.append(np.nanmean(sel, axis=ax, keepdims=True) if sel.shape[ax] else
                                np.full(sel.shape[:ax] + (1,) + sel.shape[ax + 1:], np.nan, da.dtype))
        v = np.concatenate(outs, axis=ax).astype(da.dtype)
        c = dict(da._coords)
        c["time"] = labels
        return DataArray(v, c, da.dims, da.name)


class Coordinates:
    def __init__(self, coords=None, dims=None, owner=None):
        self._c = dict(coords or {})
        self.dims = tuple(dims) if dims is not None else tuple(self._c)
        self._owner = owner

    @property
    def sizes(self):
        return {d: len(self._c[d]) for d in self.dims}

    def __getitem__(self, k):
        v = self._c[k]
        dim = {"lon": "x", "lat": "y"}.get(k, k) if k not in self.dims else k
        return DataArray(np.asarray(v), {dim: self._c.get(dim, v)}, (dim,), k)

    def __contains__(self, k):
        return k in self._c

    def __iter__(self):
        return iter(self._c)

    def keys(self):
        return self._c.keys()

    def items(self):
        return self._c.items()

    def assign(self, **kw):
        c = dict(self._c)
        dims = list(self.dims)
        for k, v in kw.items():
            c[k] = v.values if isinstance(v, DataArray) else v
            if k not in dims:
                dims.append(k)
        return Coordinates(c, dims)

    @classmethod
    def from_pandas_multiindex(cls, index, name):
        return cls({name: index}, [name])


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._coords = {}
        for k, v in (coords or {}).items():
            self._coords[k] = v.values if isinstance(v, DataArray) else (v if isinstance(v, pd.Index) else np.asarray(v))
        self._vars = {}
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, DataArray):
            for ck, cv in v._coords.items():
                self._coords.setdefault(ck, cv)
            self._vars[k] = (v.dims, v.values, dict(v.attrs))
        elif isinstance(v, tuple):
            self._vars[k] = (tuple(v[0]), np.asarray(v[1]), {})
        else:
            a = np.asarray(v)
            dims = {3: ("time", "y", "x"), 0: ()}.get(a.ndim)
            if dims is None:
                raise ValueError("ambiguous dims")
            self._vars[k] = (dims, a, {})

    def __contains__(self, k):
        return k in self._vars or k in self._coords

    def __iter__(self):
        return iter(self._vars)

    @property
    def data_vars(self):
        return {k: self[k] for k in self._vars}

    @property
    def coords(self):
        return Coordinates(self._coords, [d for d in ("time", "y", "x") if d in self._coords])

    @property
    def indexes(self):
        return {d: pd.Index(self._coords[d], name=d) for d in ("time", "y", "x") if d in self._coords}

    @property
    def chunksizes(self):
        return {}

    def __getitem__(self, k):
        if isinstance(k, (set, list)):
            return Dataset({n: self[n] for n in k}, self._coords)
        if k in self._vars:
            dims, a, attrs = self._vars[k]
            c = {d: self._coords[d] for d in dims if d in self._coords}
            for extra in ("lon", "lat"):
                if extra in self._coords:
                    c[extra] = self._coords[extra]
            return DataArray(a, c, dims, k, attrs)
        if k in self._coords:
            dim = {"lon": "x", "lat": "y"}.get(k, k)
            return DataArray(np.asarray(self._coords[k]), {dim: self._coords[dim]}, (dim,), k)
        raise KeyError(k)

    def rename(self, mapping):
        out = Dataset(coords=self._coords)
        for k in self._vars:
            out._vars[mapping.get(k, k)] = self._vars[k]
        return out

    def drop_vars(self, names, errors="raise"):
        names = [names] if isinstance(names, str) else list(names)
        out = Dataset(coords={k: v for k, v in self._coords.items() if k not in names}, attrs=self.attrs)
        for k, v in self._vars.items():
            if k not in names:
                out._vars[k] = v
        return out

    def load(self, **kw):
        return self


def merge(objs):
    out = Dataset(attrs=getattr(objs[0], "attrs", {}))
    for o in objs:
        for ck, cv in o._coords.items():
            out._coords.setdefault(ck, cv)
        for k, v in o._vars.items():
            out._vars[k] = v
    return out


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=None, output_dtypes=None,
                dask=None, **kw):
    """Only the element-wise form used by convert.py:650-657."""
    assert all(not c for c in (input_core_dims or [[]])) and all(not c for c in (output_core_dims or [[]]))
    da = args[0]
    return DataArray(func(*[_unwrap(a) for a in args]), da._coords, da.dims, da.name)


def date_range(*a, **k):
    return pd.date_range(*a, **k)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules and load the reference's hot-path modules
    from their source files (atlite/__init__.py is NOT executed: it pulls in
    the GIS / download stack).  Returns the loaded ``atlite.convert`` module."""
    if "atlite.convert" in sys.modules and getattr(sys.modules["atlite.convert"], "_shimmed", False):
        return sys.modules["atlite.convert"]
    xr = _module("xarray", DataArray=DataArray, Dataset=Dataset, Coordinates=Coordinates,
                 apply_ufunc=apply_ufunc, date_range=date_range, merge=merge)
    xr.testing = types.SimpleNamespace()

    class _DaskArray:  # isinstance(da.data, Array) is False for NumPy-backed data
        pass

    ufuncs = {n: getattr(np, n) for n in
              ("sin", "cos", "arcsin", "arccos", "arctan", "arctan2", "radians", "sqrt", "fmin", "fmax",
               "absolute", "maximum", "minimum", "mod", "logical_and", "logical_or", "exp", "log")}
    d = _module("dask", compute=lambda *a, **k: a, delayed=lambda f=None, **k: f)
    d.array = _module("dask.array", **ufuncs)
    _module("dask.array.core", Array=_DaskArray)
    _module("dask.utils", SerializableLock=type("SerializableLock", (), {}))
    _module("cdsapi")  # datasets/era5.py imports it at module level; never called here

    class ProgressBar:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    _module("dask.diagnostics", ProgressBar=ProgressBar)
    _module("geopandas", GeoDataFrame=type("GeoDataFrame", (), {}), GeoSeries=type("GeoSeries", (), {}))

    import scipy.sparse as sp

    def spdiag(v):  # the reference's gis.spdiag (gis.py:78-84) lives in a module that needs rasterio
        v = np.asarray(v)
        n = len(v)
        inds = np.arange(n + 1, dtype=np.int32)
        return sp.csr_matrix((v, inds[:-1], inds), (n, n))

    pkg = _module("atlite")
    pkg.__path__ = [REF]
    _module("atlite.gis", spdiag=spdiag, maybe_swap_spatial_dims=lambda ds, *a, **k: ds)
    _module("atlite.datasets", modules={})
    _module("atlite.hydro")
    pvpkg = _module("atlite.pv")
    pvpkg.__path__ = [REF + "/pv"]

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, f"{REF}/{rel}")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("atlite.utils", "utils.py")
    load("atlite.resource", "resource.py")
    load("atlite.aggregate", "aggregate.py")
    load("atlite.wind", "wind.py")
    for m in ("solar_position", "orientation", "irradiation", "solar_panel_model"):
        load(f"atlite.pv.{m}", f"pv/{m}.py")
    load("atlite.csp", "csp.py")
    conv = load("atlite.convert", "convert.py")
    conv._shimmed = True
    return conv


def load_reference_era5():
    """atlite/datasets/era5.py from its source file, for the arithmetic after the download
    (get_data_wind / get_data_influx / sanitize_*).  The caller replaces ``retrieve_data``
    and ``_rename_and_clean_coords`` (CDS request + coordinate renaming) by a stand-in."""
    install()
    spec = importlib.util.spec_from_file_location("atlite.datasets.era5", f"{REF}/datasets/era5.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["atlite.datasets.era5"] = mod
    spec.loader.exec_module(mod)
    return mod
