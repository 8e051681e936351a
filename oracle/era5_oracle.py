"""CPU oracle for the ERA5 prepare-time derivations (TEST INFRASTRUCTURE ONLY -- never
imported by atlite_b200).  NumPy float64 restatement of the arithmetic in
/root/reference/atlite/datasets/era5.py, each function citing the lines it follows.

PARITY PINNING: **pinned against the reference's own source.**  The reference computes
these inside ``get_data_*`` after a CDS download; tests/golden/make_golden.py::era5_cases
loads datasets/era5.py from /root/reference under the xarray/dask stand-in
(tests/golden/xr_shim.py), replaces only ``retrieve_data`` (the download) and
``_rename_and_clean_coords`` (coordinate renaming) and records the outputs of
get_data_wind / sanitize_wind / get_data_influx / sanitize_influx (incl. the stored solar
position) in tests/golden/reference_era5.npz.  tests/test_era5.py holds this oracle to
them (the reference computes in float32: 6e-7; solar position 1e-12), plus
hand-computable known answers.  Not pinned against genuine xarray/dask (absent here).
"""

from __future__ import annotations

import numpy as np


def get_data_wind(u100, v100, u10, v10, fsr, sanitize=True):
    """era5.py:120-135 (+ sanitize_wind :141-146)."""
    u100, v100, u10, v10, fsr = (np.asarray(a, dtype=np.float64) for a in (u100, v100, u10, v10, fsr))
    wnd100m = np.sqrt(u100**2 + v100**2)                                   # :121
    wnd10m = np.sqrt(u10**2 + v10**2)
    with np.errstate(divide="ignore", invalid="ignore"):
        shear = np.log(wnd10m / wnd100m) / np.log(10 / 100)               # :124-126
    az = np.arctan2(u100, v100)                                            # :129
    az = np.where(az >= 0, az, az + 2 * np.pi)                             # :130 (NaN keeps NaN + 2 pi = NaN)
    rough = fsr.copy()
    if sanitize:
        rough = np.where(rough >= 0.0, rough, 2e-4)                        # :145
    return dict(wnd100m=wnd100m, wnd_shear_exp=shear, wnd_azimuth=az, roughness=rough)


def get_data_influx(ssrd, ssr, tisr, fdir, sanitize=True):
    """era5.py:163-175 (+ sanitize_influx :195-201)."""
    ssrd, ssr, tisr, fdir = (np.asarray(a, dtype=np.float64) for a in (ssrd, ssr, tisr, fdir))
    with np.errstate(divide="ignore", invalid="ignore"):
        albedo = (ssrd - ssr) / np.where(ssrd != 0, ssrd, np.nan)          # :165
    albedo = np.where(np.isnan(albedo), 0.0, albedo)                       # :166 fillna(0)
    out = dict(albedo=albedo, influx_diffuse=ssrd - fdir, influx_direct=fdir.copy(), influx_toa=tisr.copy())  # :168
    for a in ("influx_direct", "influx_diffuse", "influx_toa"):
        out[a] = out[a] / (60.0 * 60.0)                                    # :174-175
        if sanitize:
            out[a] = np.where(out[a] < 0.0, 0.0, out[a])                   # :199-200 clip(min=0), NaN stays
    return out
