"""Self-tests of tests/golden/xr_shim.py -- the stand-in for the xarray / dask CONTAINER
API under which the reference's own hot-path source is executed to make the golden vectors.

The golden vectors are only as good as the stand-in's reading of xarray.  Every expectation
below is a literal value that follows from xarray's DOCUMENTED behaviour (user guide
"Computation": broadcasting by dimension name, missing values; API reference of
DataArray.where / fillna / clip / sum / mean / stack / transpose / resample / interp), not
from running the shim; the reference file:line that relies on the behaviour is cited."""

import os
import sys

import numpy as np
import pandas as pd
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import xr_shim as xs  # noqa: E402

nan = np.nan


def da(values, dims, **coords):
    return xs.DataArray(np.asarray(values, dtype=float), coords, dims)


def test_broadcasting_is_by_dimension_name_not_position():
    """xarray user guide, "Broadcasting by dimension name": operands are aligned by dim NAME and
    missing dims are inserted, whatever their position.  (pv/orientation.py:114-117 multiplies
    (y,) tables with (time, y, x) fields.)"""
    a = da([1, 2], ("x",), x=[10, 20])
    b = da([10, 20, 30], ("y",), y=[0, 1, 2])
    r = a + b
    assert r.dims == ("x", "y")
    np.testing.assert_array_equal(r.values, [[11, 21, 31], [12, 22, 32]])
    r = b + a
    assert r.dims == ("y", "x")
    np.testing.assert_array_equal(r.values, [[11, 12], [21, 22], [31, 32]])


def test_result_dims_are_ordered_by_first_appearance():
    """"the dimensions of the result are ordered by their first appearance in the operands":
    sin(slope[y]) * field[time, y, x] comes out as (y, time, x) -- the reason the reference's
    per-cell PV cube has dims (y, time, x) (pv/orientation.py:115)."""
    slope = da([1.0, 2.0], ("y",), y=[0, 1])
    field = da(np.arange(12).reshape(3, 2, 2), ("time", "y", "x"), time=[0, 1, 2], y=[0, 1], x=[0, 1])
    r = np.sin(slope) * field
    assert r.dims == ("y", "time", "x") and r.shape == (2, 3, 2)
    np.testing.assert_allclose(r.values[1, 2, 0], np.sin(2.0) * field.values[2, 1, 0])
    assert (field * slope).dims == ("time", "y", "x")


def test_numpy_operands_align_positionally_on_the_trailing_axes():
    """A raw ndarray has no dim names: NumPy broadcasting rules apply against the DataArray's own
    axis order (irradiation.py:198-200 passes `influx_max....data` for exactly this reason)."""
    f = da(np.zeros((2, 3)), ("y", "x"))
    r = f + np.array([1.0, 2.0, 3.0])
    assert r.dims == ("y", "x")
    np.testing.assert_array_equal(r.values, [[1, 2, 3], [1, 2, 3]])


def test_where_fillna_clip_follow_the_missing_value_rules():
    """DataArray.where(cond): "Locations at which to preserve this object's values", others
    become NaN (or `other`); fillna replaces NaN only; clip -> numpy.clip semantics: NaN stays
    NaN, a NaN BOUND gives NaN.  (irradiation.py:132, 198-200, 226, 252; solar_panel_model.py:23-36)"""
    x = da([-1.0, 0.0, 2.0, nan], ("t",))
    np.testing.assert_array_equal(x.where(x > 0).values, [nan, nan, 2.0, nan])
    np.testing.assert_array_equal(x.where(x > 0, 0).values, [0.0, 0.0, 2.0, 0.0])
    np.testing.assert_array_equal(x.fillna(7.0).values, [-1.0, 0.0, 2.0, 7.0])
    np.testing.assert_array_equal(x.clip(min=0).values, [0.0, 0.0, 2.0, nan])
    np.testing.assert_array_equal(x.clip(max=1).values, [-1.0, 0.0, 1.0, nan])
    np.testing.assert_array_equal(x.clip(min=0, max=np.array([5.0, 5.0, nan, 5.0])).values, [0.0, 0.0, nan, nan])
    # comparisons with NaN are False, so `~(a | b)` keeps cells whose inputs are NaN (irradiation.py:252)
    np.testing.assert_array_equal((x <= 0.01).values, [True, True, False, False])
    np.testing.assert_array_equal((~((x < 0) | (x <= 0.01))).values, [False, False, True, True])


def test_sum_and_mean_skip_nan_by_default():
    """DataArray.sum / mean: skipna defaults to True for float dtypes -- an all-NaN slice sums to 0
    and averages to NaN.  (convert.py:51-56 _aggregate_time; convert.py:259-262 capacity)"""
    x = da([[1.0, nan], [3.0, nan]], ("time", "s"))
    np.testing.assert_array_equal(x.sum("time").values, [4.0, 0.0])
    with np.errstate(invalid="ignore"):
        m = x.mean("time").values
    assert m[0] == 2.0 and np.isnan(m[1])
    assert x.sum("time").dims == ("s",)
    kept = xs.DataArray(x.values, {}, ("time", "s"), attrs={"units": "MW"}).sum("time", keep_attrs=True)
    assert kept.attrs == {"units": "MW"} and x.sum("time").attrs == {}


def test_stack_is_row_major_over_the_listed_dims_and_transpose_reorders():
    """stack(spatial=("y", "x")): the new dim goes LAST and runs over (y, x) in C order -- flat index
    iy * nx + ix, the column order of the indicator matrix (aggregate.py:22; cutout.py:369-370)."""
    f = da(np.arange(24).reshape(2, 3, 4), ("time", "y", "x"), time=[0, 1], y=[0, 1, 2], x=[0, 1, 2, 3])
    s = f.stack(spatial=("y", "x"))
    assert s.dims == ("time", "spatial") and s.shape == (2, 12)
    assert s.values[1, 2 * 4 + 3] == f.values[1, 2, 3]
    t = f.transpose("x", "time", "y")
    assert t.dims == ("x", "time", "y") and t.values[3, 1, 2] == f.values[1, 2, 3]


def test_resample_1d_mean_bins_are_calendar_days():
    """resample(time="1D").mean(): bins are calendar days, left-closed, labelled by their start;
    a partial first / last day averages the samples present (skipping NaN); a day without samples
    inside the range is NaN.  (convert.py:408-412 with the hour_shift applied to the time axis first)"""
    t = pd.date_range("2013-01-01 21:00", periods=8, freq="h").append(pd.DatetimeIndex(["2013-01-04 05:00"]))
    v = np.array([1.0, 2.0, 3.0, 10.0, nan, 30.0, 40.0, 50.0, 7.0])
    r = xs.DataArray(v, {"time": t}, ("time",)).resample(time="1D").mean()
    labels = pd.DatetimeIndex(r.coords["time"].values if hasattr(r.coords["time"], "values") else r.coords["time"])
    assert list(labels) == list(pd.date_range("2013-01-01", periods=4, freq="D"))
    with np.errstate(invalid="ignore"):
        got = np.asarray(r.values)
    np.testing.assert_allclose(got[:2], [2.0, (10 + 30 + 40 + 50) / 4.0])
    assert np.isnan(got[2]) and got[3] == 7.0


def test_interp_is_linear_on_a_regular_grid():
    """DataArray.interp(...) (scipy interpn, method="linear") on (altitude, azimuth): csp.py:18-58."""
    eff = xs.DataArray(np.array([[0.0, 1.0], [2.0, 3.0]]), {"altitude": [0.0, 1.0], "azimuth": [0.0, 2.0]},
                       ("altitude", "azimuth"))
    pts_alt = xs.DataArray(np.array([0.5, 1.0]), {}, ("p",))
    pts_az = xs.DataArray(np.array([1.0, 0.5]), {}, ("p",))
    r = eff.interp(altitude=pts_alt, azimuth=pts_az)
    np.testing.assert_allclose(np.asarray(r.values), [1.5, 2.25])


def test_dt_accessor_and_dataset_basics():
    t = pd.date_range("2013-03-09 05:30", periods=3, freq="h")
    d = xs.DataArray(t.values, {"time": t}, ("time",))
    np.testing.assert_array_equal(np.asarray(d.dt.hour), [5, 6, 7])
    np.testing.assert_array_equal(np.asarray(d.dt.minute), [30, 30, 30])
    ds = xs.Dataset({"a": (("time",), np.arange(3.0))}, coords={"time": t})
    assert "a" in ds and "b" not in ds and ds["a"].dims == ("time",)
    ds2 = ds.rename({"a": "b"})
    assert "b" in ds2 and "a" not in ds2
    with pytest.raises(KeyError):
        ds["nope"]


def test_apply_ufunc_maps_elementwise_functions():
    """apply_ufunc(np.interp, da, V, P) -- how convert_wind evaluates the power curve (convert.py:648-656)."""
    x = da([[0.0, 5.0], [10.0, nan]], ("y", "x"))
    r = xs.apply_ufunc(np.interp, x, np.array([0.0, 10.0]), np.array([0.0, 1.0]))
    assert r.dims == ("y", "x")
    np.testing.assert_array_equal(np.asarray(r.values)[0], [0.0, 0.5])
    assert np.asarray(r.values)[1, 0] == 1.0 and np.isnan(np.asarray(r.values)[1, 1])
