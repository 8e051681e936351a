"""Turbine / panel resource tables for the wind and PV operators.

Provides the same lookups the reference exposes (resource.py:50-141, 227-372,
514-518): ``get_windturbineconfig``, ``get_solarpanelconfig``,
``windturbine_smooth`` and the ``windturbines`` / ``solarpanels`` registries.
The data tables are shipped as two JSON files exported from the reference's
YAML data sheets by ``tools/import_resources.py``; YAML files in the reference
format can still be passed as ``pathlib.Path``.
"""

from __future__ import annotations

import json
import logging
import os
from pathlib import Path

import numpy as np

logger = logging.getLogger(__name__)

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")


def _load(name):
    with open(os.path.join(_DIR, name)) as fh:
        return json.load(fh)


class _Registry(dict):
    """dict with attribute access, like the reference's ``arrowdict``."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:
            raise AttributeError(item) from e

    def __dir__(self):
        return list(self.keys())


_TURBINES = _load("turbines.json")
_PANELS = _load("panels.json")
_CSP = _load("csp.json")
# registries map name -> name (the reference maps name -> yaml path)
windturbines = _Registry({k: k for k in _TURBINES})
solarpanels = _Registry({k: k for k in _PANELS})
cspinstallations = _Registry({k: k for k in _CSP})


def _read_yaml(path):
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


def _max_v_is_zero_pow(turbine):
    return np.any(turbine["POW"][turbine["V"] == turbine["V"].max()] == 0)


def _validate_turbine(turbine, add_cutout_windspeed):
    """Checks of resource.py:304-372 with the same error messages' gist."""
    need = ("POW", "V", "P", "hub_height")
    if not all(k in turbine for k in need):
        raise ValueError(
            f"turbine config dict needs at least the following keys: {list(need)}\n"
            f"but are currently: {list(turbine.keys())}"
        )
    if not all(isinstance(turbine[p], (np.ndarray, list)) for p in ("POW", "V")):
        raise ValueError("turbine entries 'POW' and 'V' must be np.ndarray or list")
    turbine = dict(turbine)
    turbine["V"] = np.asarray(turbine["V"], dtype=float)
    turbine["POW"] = np.asarray(turbine["POW"], dtype=float)
    if len(turbine["POW"]) != len(turbine["V"]):
        raise ValueError("turbine wind speed and power arrays do not have equal length.")
    if not np.all(np.diff(turbine["V"]) >= 0):
        raise ValueError(
            "wind speed 'V' in the turbine config dict is expected to be increasing, "
            f"but is currently not in ascending order:\n{turbine['V']}"
        )
    if add_cutout_windspeed is True and not _max_v_is_zero_pow(turbine):
        turbine["V"] = np.append(turbine["V"], turbine["V"].max())
        turbine["POW"] = np.append(turbine["POW"], 0.0)
        logger.info(
            "adding a cut-out wind speed to the turbine power curve at V=%s m/s.",
            turbine["V"][-1],
        )
    if not _max_v_is_zero_pow(turbine):
        logger.warning(
            "The power curve does not have a cut-out wind speed, i.e. the power output "
            "corresponding to the\nhighest wind speed is not zero. You can either change "
            "the power curve manually or set\n'add_cutout_windspeed=True' in the "
            "Cutout.wind conversion method."
        )
    return turbine


def get_windturbineconfig(turbine, add_cutout_windspeed=True):
    """Turbine name | Path to a reference-format YAML | config dict -> validated dict
    with ``V``, ``POW`` (ndarrays), ``hub_height`` and ``P = max(POW)``."""
    if not isinstance(turbine, (str, Path, dict)):
        raise KeyError(f"`turbine` must be a str, pathlib.Path or dict, but is {type(turbine)}.")
    if isinstance(turbine, str) and turbine.startswith("oedb:"):
        raise NotImplementedError("OEDB download needs network access; pass a dict instead")
    if isinstance(turbine, str):
        name = turbine.replace(".yaml", "")
        if name not in _TURBINES:
            raise KeyError(name)
        d = _TURBINES[name]
        conf = dict(
            V=np.array(d["V"], dtype=float),
            POW=np.array(d["POW"], dtype=float),
            hub_height=d["hub_height"],
            P=float(np.max(d["POW"])),
        )
    elif isinstance(turbine, Path):
        d = _read_yaml(turbine)
        conf = dict(
            V=np.array(d["V"], dtype=float),
            POW=np.array(d["POW"], dtype=float),
            hub_height=d["HUB_HEIGHT"],
            P=float(np.max(d["POW"])),
        )
    else:
        conf = turbine
    return _validate_turbine(conf, add_cutout_windspeed)


def get_solarpanelconfig(panel):
    """Panel name | Path to a reference-format YAML -> coefficient dict."""
    assert isinstance(panel, (str, Path))
    if isinstance(panel, str):
        name = panel.replace(".yaml", "")
        if name not in _PANELS:
            raise KeyError(name)
        return dict(_PANELS[name])
    return _read_yaml(panel)


def windturbine_rated_capacity_per_unit(turbine):
    if isinstance(turbine, (str, Path)):
        turbine = get_windturbineconfig(turbine)
    return turbine["P"]


def windturbine_smooth(turbine, params=None):
    """Gaussian smoothing of the power curve (Andresen et al. 2015), producing the
    72-knot curve on linspace(0, 35, 72) the reference produces (resource.py:227-297)."""
    from scipy.signal import fftconvolve

    if params is None or params is True:
        params = {}
    eta = params.get("eta", 0.95)
    delta_v = params.get("Delta_v", 1.27)
    sigma = params.get("sigma", 2.29)

    v_reg = np.linspace(-50.0, 50.0, 1001)  # 0.1 m/s steps
    p_reg = np.interp(v_reg, turbine["V"], turbine["POW"])
    kern = np.exp(-((v_reg - delta_v) ** 2) / (2 * sigma * sigma)) / np.sqrt(
        2 * np.pi * sigma * sigma
    )
    conv = 0.1 * fftconvolve(p_reg, kern, mode="same")
    v_new = np.linspace(0.0, 35.0, 72)
    p_new = eta * np.interp(v_new, v_reg, conv)

    out = dict(turbine)
    out["V"], out["POW"] = v_new, p_new
    out["P"] = float(np.max(p_new))
    if any(out["POW"][np.where(out["V"] == 0.0)] > 1e-2):
        logger.warning(
            "Oversmoothing detected with parameters eta=%f, Delta_v=%f, sigma=%f. "
            "Turbine generates energy at 0 m/s wind speeds.",
            eta,
            delta_v,
            sigma,
        )
    return out


class EfficiencyTable:
    """Solar-field efficiency over (altitude, azimuth) in radians / p.u. -- the
    information of the reference's ``installation["efficiency"]`` DataArray
    (resource.py:190-220)."""

    dims = ("altitude", "azimuth")

    def __init__(self, altitude, azimuth, values):
        self.altitude = np.asarray(altitude, dtype=np.float64)
        self.azimuth = np.asarray(azimuth, dtype=np.float64)
        self.values = np.asarray(values, dtype=np.float64)
        if self.values.shape != (len(self.altitude), len(self.azimuth)):
            raise ValueError("efficiency table shape does not match its coordinates")
        if np.any(np.diff(self.altitude) <= 0) or np.any(np.diff(self.azimuth) <= 0):
            raise ValueError("efficiency table coordinates must be strictly increasing")

    @property
    def coords(self):
        return {"altitude": self.altitude, "azimuth": self.azimuth}

    @classmethod
    def from_any(cls, eff):
        if isinstance(eff, cls):
            return eff
        # xarray.DataArray with 'altitude' / 'azimuth' coordinates in rad
        e = eff.transpose("altitude", "azimuth")
        return cls(np.asarray(e.coords["altitude"]), np.asarray(e.coords["azimuth"]), np.asarray(e.values))


def get_cspinstallationconfig(installation):
    """CSP installation name | Path to a reference-format YAML -> config dict with
    'technology', 'r_irradiance' and 'efficiency' (an ``EfficiencyTable`` in rad / p.u.;
    the reference converts deg -> rad and % -> p.u. the same way, resource.py:200-220)."""
    assert isinstance(installation, (str, Path))
    if isinstance(installation, str):
        name = installation.replace(".yaml", "")
        if name not in _CSP:
            raise KeyError(name)
        d = _CSP[name]
        alt, az, tab = d["altitude_deg"], d["azimuth_deg"], d["efficiency_percent"]
        config = {"technology": d["technology"], "r_irradiance": d["r_irradiance"], **d.get("meta", {})}
    else:
        import pandas as pd

        d = _read_yaml(installation)
        df = pd.DataFrame(d["efficiency"]).set_index(["altitude", "azimuth"])["value"].unstack("azimuth")
        alt, az, tab = df.index.values, df.columns.values, df.values
        config = {k: v for k, v in d.items() if k != "efficiency"}
    config["path"] = installation
    config["efficiency"] = EfficiencyTable(np.radians(np.asarray(alt, float)), np.radians(np.asarray(az, float)),
                                           np.asarray(tab, float) / 1.0e2)
    return config
