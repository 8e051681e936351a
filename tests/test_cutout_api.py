"""Host-side members of ``Cutout`` next to the hot path (reference cutout.py:211-651): identity and
geometry properties, ``sel``, ``layout_from_capacity_list``, ``area`` / ``uniform_density_layout``,
``equals``, ``to_file``.  No GPU needed."""

import os

import numpy as np
import pandas as pd
import pytest

import atlite_b200 as ab
from atlite_b200 import labelled, synthetic as syn


@pytest.fixture(scope="module")
def cutout():
    return ab.Cutout(data=syn.make_dataset(40, 24, 48, x0=5.0, y0=45.0))


def test_identity_and_geometry_properties(cutout, tmp_path):
    c = cutout
    assert c.name is None and c.module == "era5" and c.dt == "h" and c.chunks is None
    assert c.shape == (24, 40) and c.dx == 0.25 and c.dy == 0.25
    np.testing.assert_allclose(c.extent, [4.875, 14.875, 44.875, 50.875])        # (x, X, y, Y)  cutout.py:266-274
    np.testing.assert_allclose(c.bounds, [4.875, 44.875, 14.875, 50.875])        # (x, y, X, Y)  cutout.py:277-281
    t, tr = c.transform, c.transform_r
    assert (t.a, t.b, t.c, t.d, t.e, t.f) == (0.25, 0.0, 4.875, 0.0, 0.25, 44.875)
    assert (tr.a, tr.c, tr.e, tr.f) == (0.25, 4.875, -0.25, 50.875)
    named = ab.Cutout(path=tmp_path / "europe-2013.nc", data=c.data)
    assert named.name == "europe-2013"
    d = labelled.Dataset({"temperature": np.zeros((2, 3, 4), np.float32)},
                         coords=dict(time=pd.date_range("2013", periods=2, freq="3h"), x=np.arange(4.0), y=np.arange(3.0)),
                         attrs={"module": "sarah", "chunksize_time": 100, "chunksize_x": 4})
    assert ab.Cutout(data=d).chunks == {"time": 100, "x": 4} and ab.Cutout(data=d).module == "sarah"


def test_sel_bounds_buffer_and_time(cutout):
    sub = cutout.sel(bounds=(6.0, 46.0, 8.0, 47.0), buffer=0.25, time=slice("2013-01-01 06:00", "2013-01-01 17:00"))
    x, y = np.asarray(sub.coords["x"]), np.asarray(sub.coords["y"])
    assert x[0] == 5.75 and x[-1] == 8.25 and y[0] == 45.75 and y[-1] == 47.25       # inclusive label slices
    assert len(sub.coords["time"]) == 12 and sub.shape == (7, 11)
    full = cutout.data.raw("temperature")
    assert np.shares_memory(sub.data.raw("temperature"), full)                       # views, no copies
    np.testing.assert_array_equal(sub.data.raw("temperature"), full[6:18, 3:10, 3:14])
    np.testing.assert_array_equal(np.asarray(sub.coords["lon"]), x)
    assert cutout.sel(x=slice(100, 200)).shape == (24, 0)                            # empty like xarray
    with pytest.raises(KeyError):
        cutout.sel(band=slice(0, 1))
    # a lazily loaded cutout stays lazy and reads only what the selection asks for
    a = np.asarray(full)
    lz = labelled.LazyDataset({"temperature": lambda lo, hi: a[lo:hi]}, dict(cutout.data.coords), time_chunk=7)
    s2 = ab.Cutout(data=lz).sel(bounds=(6.0, 46.0, 8.0, 47.0), time=slice("2013-01-01 06:00", None))
    assert getattr(s2.data, "lazy", False) and s2.shape == (5, 9)
    np.testing.assert_array_equal(s2.data.isel_time(2, 5).raw("temperature"), a[8:11, 4:9, 4:13])


def reference_layout(cutout, data, col="Capacity"):
    """cutout.py:637-651 with pandas only (the reference ends in ``.to_xarray().reindex_like().fillna(0)``)."""
    x_grid, y_grid = np.asarray(cutout.coords["x"]), np.asarray(cutout.coords["y"])
    ix = np.clip(np.searchsorted(x_grid, data.x.values, side="left"), 0, len(x_grid) - 1)
    iy = np.clip(np.searchsorted(y_grid, data.y.values, side="left"), 0, len(y_grid) - 1)
    ix = ix - (data.x.values - x_grid[ix - 1] < x_grid[ix] - data.x.values)
    iy = iy - (data.y.values - y_grid[iy - 1] < y_grid[iy] - data.y.values)
    g = data.assign(x=x_grid[ix], y=y_grid[iy]).groupby(["y", "x"])[col].sum()
    return g.unstack("x").reindex(index=y_grid, columns=x_grid).fillna(0).values


def test_layout_from_capacity_list_matches_the_reference_arithmetic(cutout):
    rng = np.random.default_rng(0)
    n = 500
    df = pd.DataFrame(dict(x=rng.uniform(4.0, 16.0, n), y=rng.uniform(44.0, 52.0, n), Capacity=rng.uniform(0, 50, n)))
    df.loc[::37, "Capacity"] = np.nan                       # groupby-sum counts NaN as 0
    df.loc[3, ["x", "y"]] = [5.0, 45.0]                     # exactly on the first coordinates (wraps, like the reference)
    df.loc[4, ["x", "y"]] = [14.75, 50.75]                  # exactly on the last
    df.loc[5, ["x", "y"]] = [7.125, 46.375]                 # exactly between two cells
    lay = cutout.layout_from_capacity_list(df)
    np.testing.assert_allclose(np.asarray(lay.values), reference_layout(cutout, df), rtol=0, atol=1e-9)
    assert lay.dims == ("y", "x") and np.isclose(np.asarray(lay.values).sum(), np.nansum(df.Capacity))
    other = cutout.layout_from_capacity_list(df.rename(columns={"Capacity": "p_nom"}), col="p_nom")
    np.testing.assert_array_equal(np.asarray(other.values), np.asarray(lay.values))


def test_area_density_layout_equals_and_to_file(cutout, tmp_path):
    np.testing.assert_allclose(np.asarray(cutout.area().values), 0.0625)
    np.testing.assert_allclose(np.asarray(cutout.uniform_density_layout(3.0).values), 0.1875)
    with pytest.raises(NotImplementedError):
        cutout.area(crs=3035)
    same = ab.Cutout(data=syn.make_dataset(40, 24, 48, x0=5.0, y0=45.0))
    assert cutout.equals(same) and not cutout.equals(cutout.sel(time=slice(None, "2013-01-01 05:00")))
    assert cutout.equals(42) is NotImplemented
    if labelled.HAVE_XARRAY:
        return
    sub = cutout.sel(bounds=(6.0, 46.0, 8.0, 47.0))
    fn = os.path.join(tmp_path, "sub.atlc")
    sub.to_file(fn)
    back = ab.Cutout(fn)
    assert back.name == "sub" and back.shape == sub.shape and getattr(back.data, "lazy", False)
    got = back.data.isel_time(0, 48)
    for k in sub.data.keys():
        if len(sub.data.dims_of(k)) == 3:
            np.testing.assert_array_equal(got.raw(k), sub.data.raw(k))
    assert ab.Cutout(data=got).equals(ab.Cutout(data=sub.data.isel_time(0, 48)))
