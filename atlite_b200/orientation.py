"""Panel orientation factories with the reference's callback protocol
``f(lon, lat, solar_position) -> dict(slope=..., azimuth=...)`` (radians),
cf. pv/orientation.py:13-88.  The fused PV kernel consumes the result as per-row
(latitude) tables, so callbacks may return scalars or arrays over ``y``.
"""

from __future__ import annotations

import numpy as np


def _values(a):
    return np.asarray(getattr(a, "values", a), dtype=np.float64)


def make_latitude_optimal():
    """Tilt rule of thumb of pv/orientation.py:50-67: 0.87|lat| up to 25 deg,
    0.76|lat| + radians(0.31) up to 50 deg (the reference adds 0.31 *degrees
    converted to radians*; reproduced as is), 40 deg beyond; facing the equator."""

    def latitude_optimal(lon, lat, solar_position):
        la = np.abs(_values(lat))
        slope = np.where(
            la <= np.radians(25),
            0.87 * la,
            np.where(la <= np.radians(50), 0.76 * la + np.radians(0.31), np.radians(40.0)),
        )
        azimuth = np.where(_values(lat) < 0, 0.0, np.pi)
        return dict(slope=slope, azimuth=azimuth)

    return latitude_optimal


def make_constant(slope, azimuth):
    slope = np.radians(slope)
    azimuth = np.radians(azimuth)

    def constant(lon, lat, solar_position):
        return dict(slope=slope, azimuth=azimuth)

    return constant


def make_latitude(azimuth=180):
    azimuth = np.radians(azimuth)

    def latitude(lon, lat, solar_position):
        return dict(slope=_values(lat), azimuth=azimuth)

    return latitude


_FACTORIES = {
    "latitude_optimal": make_latitude_optimal,
    "constant": make_constant,
    "latitude": make_latitude,
}


def get_orientation(name, **params):
    """'latitude_optimal' | 'latitude' | {'slope': deg, 'azimuth': deg[, 'name': ...]}."""
    if isinstance(name, dict):
        params = dict(name)
        name = params.pop("name", "constant")
    try:
        factory = _FACTORIES[name]
    except KeyError as e:
        raise AttributeError(f"unknown orientation 'make_{name}'") from e
    return factory(**params)
