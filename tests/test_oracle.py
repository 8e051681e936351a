"""CPU tests of the oracle itself: the orchestration semantics the reference
pins in test/test_aggregate_time.py and the physical properties it checks in
test/test_preparation_and_conversion.py (night => exactly 0, no NaN, > 0,
capacity == layout sums, tracking ordering), on synthetic data."""


import numpy as np
import pytest
from conftest import oracle_ds

import atlite_oracle as O
from atlite_b200 import synthetic as syn
from atlite_b200.resource import get_solarpanelconfig, get_windturbineconfig


@pytest.fixture(scope="module")
def ds():
    return oracle_ds(syn.make_dataset(20, 12, 48, x0=-5.0, y0=40.0))


def identity_convert(d, **kw):
    return d["var"]


@pytest.fixture
def tiny():
    rng = np.random.RandomState(42)
    import pandas as pd

    return dict(
        var=rng.rand(24, 3, 4),
        time=pd.date_range("2020-01-01", periods=24, freq="h"),
        lon=np.array([5.0, 6.0, 7.0, 8.0]),
        lat=np.array([50.0, 51.0, 52.0]),
    )


def test_aggregate_time_semantics(tiny):
    # test/test_aggregate_time.py:44-66
    r = O.convert_and_aggregate(tiny, identity_convert, aggregate_time=None)
    assert r.shape == (24, 3, 4)
    np.testing.assert_allclose(
        O.convert_and_aggregate(tiny, identity_convert, aggregate_time="mean"), tiny["var"].mean(0))
    np.testing.assert_allclose(
        O.convert_and_aggregate(tiny, identity_convert, aggregate_time="sum"), tiny["var"].sum(0))
    np.testing.assert_allclose(
        O.convert_and_aggregate(tiny, identity_convert), tiny["var"].sum(0))  # legacy, no spatial
    with pytest.raises(ValueError):
        O.convert_and_aggregate(tiny, identity_convert, aggregate_time="invalid")


def test_layout_and_per_unit(tiny):
    # test/test_aggregate_time.py:83-128
    lay = np.ones((3, 4)) * 2.0
    ts = O.convert_and_aggregate(tiny, identity_convert, layout=lay, aggregate_time=None)
    assert ts.shape == (24, 1)
    np.testing.assert_allclose(ts[:, 0], 2.0 * tiny["var"].reshape(24, -1).sum(1))
    pu = O.convert_and_aggregate(tiny, identity_convert, layout=lay, per_unit=True, aggregate_time=None)
    np.testing.assert_allclose(pu[:, 0], tiny["var"].reshape(24, -1).mean(1))
    mean = O.convert_and_aggregate(tiny, identity_convert, layout=lay, per_unit=True, aggregate_time="mean")
    np.testing.assert_allclose(mean, pu.mean(0))
    with pytest.raises(ValueError):
        O.convert_and_aggregate(tiny, identity_convert, per_unit=True)


def test_pv_properties(ds):
    panel = get_solarpanelconfig("CdTe")
    flat = O.convert_pv(ds, panel, O.get_orientation({"slope": 0.0, "azimuth": 0.0}))
    opt = O.convert_pv(ds, panel, O.get_orientation("latitude_optimal"))
    assert not np.isnan(flat).any() and not np.isnan(opt).any()
    assert flat.sum() > 0 and opt.sum() > flat.sum()
    alt = O.solar_position(ds)["altitude"]
    assert (flat[alt < 0] == 0).all() and (opt[alt < np.radians(1.0)] == 0).all()
    hd = O.convert_pv(ds, panel, O.get_orientation("latitude_optimal"), trigon_model="other")
    assert round(hd.sum() / opt.sum()) == 1
    kan = O.convert_pv(ds, get_solarpanelconfig("KANENA"), O.get_orientation("latitude_optimal"))
    assert round(kan.sum() / opt.sum()) == 1


def test_pv_tracking_ordering(ds):
    # test/test_preparation_and_conversion.py:155-223: dual >= 1-axis >= fixed
    panel = get_solarpanelconfig("CSi")
    o = O.get_orientation({"slope": 0.0, "azimuth": 180.0})
    cf = {t: O.convert_pv(ds, panel, o, tracking=t).mean() for t in (None, "horizontal", "vertical", "dual")}
    ot = O.get_orientation({"slope": 30.0, "azimuth": 180.0})
    cf["tilted_horizontal"] = O.convert_pv(ds, panel, ot, tracking="tilted_horizontal").mean()
    cf["tilted_fixed"] = O.convert_pv(ds, panel, ot).mean()
    assert cf["dual"] >= cf["horizontal"] >= cf[None]
    assert cf["dual"] >= cf["tilted_horizontal"] >= cf["tilted_fixed"]
    assert cf["dual"] >= cf["vertical"]


def test_wind_properties(ds):
    t = get_windturbineconfig("Vestas_V112_3MW")
    cf = O.convert_wind(ds, t)
    assert cf.dtype == np.float64 and not np.isnan(cf).any()
    assert cf.min() >= 0 and cf.max() <= 1 and cf.sum() > 0
    # np.interp at a duplicated knot returns the LAST value (cut-out => 0)
    assert np.interp(25.0, t["V"], t["POW"]) == 0.0
    sm = O.windturbine_smooth(t)
    assert len(sm["V"]) == 72
    pw = O.convert_wind({**ds, "wnd_shear_exp": np.full_like(ds["wnd100m"], 0.14)}, t, "power")
    assert pw.sum() > 0


def test_heat_demand_bins(ds):
    hd, labels = O.convert_heat_demand(ds, threshold=15.0, a=1.0, constant=0.0, hour_shift=0.0)
    assert hd.shape == (2, 12, 20) and len(labels) == 2 and (hd >= 0).all()
    hd2, labels2 = O.convert_heat_demand(ds, threshold=15.0, a=2.0, constant=1.0, hour_shift=5.0)
    assert hd2.shape[0] == 3 and len(labels2) == 3  # partial edge days
    T = ds["temperature"]
    np.testing.assert_allclose(hd2[0], 1.0 + 2.0 * np.maximum(288.15 - T[:19].mean(0), 0), rtol=1e-5)


def test_capacity_matches_layout_sums(ds):
    # test/test_preparation_and_conversion.py:98-114
    lay = syn.make_layout(20, 12)
    m = syn.make_shapes(20, 12, 4)
    res, cap = O.convert_and_aggregate(
        ds, O.convert_wind, matrix=m, layout=lay, return_capacity=True,
        turbine=get_windturbineconfig("Vestas_V112_3MW"))
    np.testing.assert_allclose(cap, np.asarray(m @ lay.reshape(-1)))
    assert res.shape == (48, 4)
