"""ERA5 prepare-time derivations: oracle known answers on the CPU,
CUDA kernels (csrc/era5.cu, through the C ABI) against the oracle on the GPU."""

import os

import numpy as np
import pandas as pd
import pytest

import atlite_oracle as O
import era5_oracle as EO
import atlite_b200 as ab
from atlite_b200 import synthetic as syn


def test_oracle_known_answers():
    w = EO.get_data_wind(u100=[3.0, 0.0, -2.0, 0.0], v100=[4.0, 1.0, 0.0, -5.0], u10=[0.3, 0.0, -1.0, 0.0],
                         v10=[0.4, 0.5, 0.0, -2.5], fsr=[0.1, -1.0, 0.0, np.nan])
    np.testing.assert_allclose(w["wnd100m"], [5, 1, 2, 5])
    # shear exponent: ln(w10/w100)/ln(0.1): 0.5/5 -> 1.0, 0.5/1 -> log10(2), ...
    np.testing.assert_allclose(w["wnd_shear_exp"], [1.0, np.log(0.5) / np.log(0.1), np.log(0.5) / np.log(0.1),
                                                    np.log(0.5) / np.log(0.1)])
    # azimuth: 0 = wind towards north, pi/2 towards east, pi south, 3 pi/2 west (era5.py:128)
    np.testing.assert_allclose(w["wnd_azimuth"], [np.arctan2(3, 4), 0.0, 1.5 * np.pi, np.pi])
    np.testing.assert_allclose(w["roughness"], [0.1, 2e-4, 0.0, 2e-4])
    i = EO.get_data_influx(ssrd=[7200.0, 0.0, 3600.0, 3600.0], ssr=[5400.0, 0.0, 3600.0, np.nan],
                           tisr=[36000.0, 0.0, -3.6, 7200.0], fdir=[3600.0, 0.0, 7200.0, 1800.0])
    np.testing.assert_allclose(i["albedo"], [0.25, 0.0, 0.0, 0.0])
    np.testing.assert_allclose(i["influx_diffuse"], [1.0, 0.0, 0.0, 0.5])   # (ssrd - fdir)/3600, clipped
    np.testing.assert_allclose(i["influx_direct"], [1.0, 0.0, 2.0, 0.5])
    np.testing.assert_allclose(i["influx_toa"], [10.0, 0.0, 0.0, 2.0])
    raw = EO.get_data_influx([3600.0], [0.0], [-3.6], [7200.0], sanitize=False)
    np.testing.assert_allclose([raw["influx_diffuse"][0], raw["influx_toa"][0]], [-1.0, -0.001])


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_era5.npz")


def _golden():
    g = np.load(GOLDEN)
    raw = {k.split("|")[1]: g[k] for k in g.files if k.startswith("raw|")}
    return g, raw


def _close(got, want, what, tol=6e-7):
    """fp32 reference values (the reference computes these in float32) vs the float64 oracle
    (tol 6e-7: the reference's own rounding) or the fp32 kernels (tol 4e-6: both sides round,
    the kernels use approximate division / sqrt): relative, plus the same absolute for the
    difference-of-logs (shear exponent) and the angles."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol, equal_nan=True, err_msg=what)


def test_oracle_is_pinned_to_the_reference_source():
    """tests/golden/reference_era5.npz was produced by datasets/era5.py's own lines
    (make_golden.py::era5_cases); the oracle must reproduce them."""
    g, raw = _golden()
    for sanitize, tag in ((False, ""), (True, "_sanitized")):
        w = EO.get_data_wind(raw["u100"], raw["v100"], raw["u10"], raw["v10"], raw["fsr"], sanitize=sanitize)
        for k in ("wnd100m", "wnd_shear_exp", "wnd_azimuth"):
            _close(w[k], g[f"wind|{k}"], k)
        _close(w["roughness"], g[f"wind{tag}|roughness"], "roughness" + tag)
        i = EO.get_data_influx(raw["ssrd"], raw["ssr"], raw["tisr"], raw["fdir"], sanitize=sanitize)
        _close(i["albedo"], g["influx|albedo"], "albedo")
        for k in ("influx_toa", "influx_direct", "influx_diffuse"):
            _close(i[k], g[f"influx{tag}|{k}"], k + tag)
    t = pd.DatetimeIndex(g["time_ns"].astype("datetime64[ns]"))
    sol = O.solar_position(dict(time=t, lon=g["x"], lat=g["y"]), time_shift="-30min")
    np.testing.assert_allclose(sol["altitude"], g["influx|solar_altitude"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sol["azimuth"], g["influx|solar_azimuth"], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_gpu_matches_the_reference_golden_vectors():
    from atlite_b200 import era5

    g, raw = _golden()
    t = pd.DatetimeIndex(g["time_ns"].astype("datetime64[ns]"))
    ds = ab.Dataset({k: (("time", "y", "x"), v) for k, v in raw.items()},
                    coords=dict(time=t, x=g["x"], y=g["y"], lon=g["x"], lat=g["y"]))
    for sanitize, tag in ((False, ""), (True, "_sanitized")):
        w = era5.get_data_wind(ds, sanitize=sanitize)
        for k in ("wnd100m", "wnd_shear_exp", "wnd_azimuth"):
            _close(w.raw(k).cpu().numpy(), g[f"wind|{k}"], k, 4e-6)
        _close(w.raw("roughness").cpu().numpy(), g[f"wind{tag}|roughness"], "roughness" + tag, 4e-6)
        i = era5.get_data_influx(ds, sanitize=sanitize)
        _close(i.raw("albedo").cpu().numpy(), g["influx|albedo"], "albedo", 4e-6)
        for k in ("influx_toa", "influx_direct", "influx_diffuse"):
            _close(i.raw(k).cpu().numpy(), g[f"influx{tag}|{k}"], k + tag, 4e-6)
    np.testing.assert_allclose(i.raw("solar_altitude").cpu().numpy(), g["influx|solar_altitude"], rtol=0, atol=1e-10)
    d = np.angle(np.exp(1j * (i.raw("solar_azimuth").cpu().numpy() - g["influx|solar_azimuth"])))
    assert np.abs(d).max() < 2e-7


def _raw_dataset(nx, ny, nt, seed=0, x0=-10.0, y0=35.0):
    rng = np.random.default_rng(seed)
    x, y = syn.make_coords(nx, ny, x0, y0, 0.5, 0.5)
    t = syn.make_time(nt + 24 * 40)[24 * 40:]
    shp = (nt, ny, nx)
    f = lambda lo, hi: rng.uniform(lo, hi, shp).astype(np.float32)  # noqa: E731
    d = dict(u100=f(-15, 15), v100=f(-15, 15), u10=f(-8, 8), v10=f(-8, 8), fsr=f(-0.05, 2.0),
             ssrd=f(0, 3.0e6), tisr=f(0, 4.5e6), t2m=f(250, 305), stl4=f(270, 290), d2m=f(250, 295), ro=f(-1e-4, 1e-3))
    d["ssr"] = (d["ssrd"] * rng.uniform(0.6, 1.0, shp)).astype(np.float32)
    d["fdir"] = (d["ssrd"] * rng.uniform(0.0, 1.1, shp)).astype(np.float32)  # sometimes > ssrd: negative diffuse
    d["ssrd"][0, 0, :5] = 0.0        # night: albedo must be 0, not NaN
    d["ssr"][0, 1, 2] = np.nan
    d["u100"][0, 0, 0], d["v100"][0, 0, 0] = 0.0, -3.0   # due south
    d["u100"][0, 0, 1], d["v100"][0, 0, 1] = -2.0, 0.0   # due west
    d["fsr"][0, 2, 2] = np.nan
    return ab.Dataset({k: (("time", "y", "x"), v) for k, v in d.items()}, coords=dict(time=t, x=x, y=y, lon=x, lat=y)), d


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 20, 30), (33, 7, 5)])  # 128-bit path and odd sizes (scalar tail)
def test_gpu_wind_influx_match_oracle(shape):
    from atlite_b200 import era5

    ds, d = _raw_dataset(*shape)
    for sanitize in (True, False):
        got = era5.get_data_wind(ds, sanitize=sanitize)
        want = EO.get_data_wind(d["u100"], d["v100"], d["u10"], d["v10"], d["fsr"], sanitize=sanitize)
        for k, w in want.items():
            g = got.raw(k).cpu().numpy()
            assert g.dtype == np.float32 and np.array_equal(np.isnan(g), np.isnan(w)), k
            np.testing.assert_allclose(g, w, rtol=3e-6, atol=3e-6, err_msg=k, equal_nan=True)
        got = era5.get_data_influx(ds, sanitize=sanitize, solar_position_vars=False)
        want = EO.get_data_influx(d["ssrd"], d["ssr"], d["tisr"], d["fdir"], sanitize=sanitize)
        for k, w in want.items():
            g = got.raw(k).cpu().numpy()
            assert np.array_equal(np.isnan(g), np.isnan(w)), k
            np.testing.assert_allclose(g, w, rtol=2e-6, atol=1e-6, err_msg=k, equal_nan=True)
        if sanitize:
            assert (got.raw("influx_diffuse").cpu().numpy() >= 0).all()
    assert (got.raw("albedo").cpu().numpy()[0, 0, :5] == 0).all()


@pytest.mark.gpu
def test_gpu_solar_position_matches_the_pinned_oracle():
    from atlite_b200 import era5

    x, y = syn.make_coords(37, 23, -170.0, -80.0, 9.0, 7.0)   # both hemispheres, across the date line
    t = pd.date_range("2013-03-20 22:00", periods=30, freq="h").append(pd.date_range("2013-12-21 00:30", periods=5, freq="7h"))
    for shift in ("-30min", "0h"):
        alt, az = era5.solar_position(t, x, y, shift)
        want = O.solar_position(dict(time=t, lon=x, lat=y), time_shift=shift)
        np.testing.assert_allclose(alt.cpu().numpy(), want["altitude"], rtol=0, atol=1e-10)
        # the azimuth's arccos amplifies rounding near 0 / pi: compare on the circle
        d = np.angle(np.exp(1j * (az.cpu().numpy() - want["azimuth"])))
        assert np.abs(d).max() < 2e-7
        assert alt.dtype.is_floating_point and alt.element_size() == 8


@pytest.mark.gpu
def test_gpu_raw_download_to_conversion_end_to_end():
    """raw ERA5 variables -> era5.prepare (device) -> Cutout.wind / pv / temperature: the same
    numbers as the oracle pipeline fed with oracle-derived fields."""
    from atlite_b200 import era5
    import scipy.sparse as sp

    nx, ny, nt = 40, 12, 48
    ds, d = _raw_dataset(nx, ny, nt, seed=4)
    cut = ab.Cutout(data=era5.prepare(ds))
    # raw fields that are already on the device take the same path
    import torch

    ds_dev = ab.Dataset({k: (("time", "y", "x"), torch.from_numpy(v).cuda()) for k, v in d.items()},
                        coords=dict(ds.coords))
    again = era5.prepare(ds_dev, features=("wind", "influx"))
    for k in ("wnd100m", "wnd_azimuth", "albedo", "influx_diffuse", "solar_azimuth"):
        assert torch.equal(again.raw(k), cut.data.raw(k)) or torch.allclose(again.raw(k), cut.data.raw(k), equal_nan=True), k
    assert {"wnd100m", "roughness", "influx_toa", "albedo", "solar_altitude", "temperature", "runoff"} <= set(cut.data.data_vars)
    m = syn.make_shapes(nx, ny, 5)
    ow = EO.get_data_wind(d["u100"], d["v100"], d["u10"], d["v10"], d["fsr"])
    oi = EO.get_data_influx(d["ssrd"], d["ssr"], d["tisr"], d["fdir"])
    co = dict(time=ds.coords["time"], lon=ds.coords["x"], lat=ds.coords["y"])
    sol = O.solar_position(co, time_shift="-30min")
    od = dict(co, wnd100m=ow["wnd100m"], roughness=ow["roughness"], temperature=d["t2m"].astype(np.float64),
              solar_altitude=sol["altitude"], solar_azimuth=sol["azimuth"],
              **{k: oi[k] for k in ("influx_toa", "influx_direct", "influx_diffuse", "albedo")})
    cap = np.asarray(sp.csr_matrix(m).sum(-1)).ravel()
    from conftest import assert_parity

    turb = ab.get_windturbineconfig("Vestas_V112_3MW")
    got = np.asarray(cut.wind(turb, matrix=m, aggregate_time=None).values).T
    want = O.convert_and_aggregate(od, O.convert_wind, matrix=m, aggregate_time=None, turbine=turb)
    assert not np.isnan(got).any()  # the NaN / negative roughness cells were sanitised
    assert_parity(got, want, cap, what="raw -> wind")
    got = np.asarray(cut.pv("CSi", "latitude_optimal", matrix=m, aggregate_time=None).values).T
    want = O.convert_and_aggregate(od, O.convert_pv, matrix=m, aggregate_time=None,
                                   panel=ab.get_solarpanelconfig("CSi"), orientation=O.get_orientation("latitude_optimal"))
    assert_parity(got, want, cap, what="raw -> pv (stored solar position)")
