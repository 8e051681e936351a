#!/usr/bin/env python
"""Export the reference's resource *data tables* (turbine power curves, panel
coefficients) into this repo's own compact JSON format.

Run once in the build container (needs /root/reference); the outputs
``atlite_b200/resources/{turbines,panels,csp}.json`` are committed, so nothing
reads /root/reference at run time.

Source data: /root/reference/atlite/resources/windturbine/*.yaml and
/root/reference/atlite/resources/solarpanel/*.yaml (CC-BY-4.0 data sheets).
The loader semantics that consume these tables are re-implemented in
``atlite_b200/resource.py`` (cf. reference atlite/resource.py:50-141).
"""
import glob
import json
import os
import sys

import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/atlite/resources"
OUT = os.path.join(os.path.dirname(__file__), "..", "atlite_b200", "resources")


def main():
    turbines = {}
    for f in sorted(glob.glob(os.path.join(REF, "windturbine", "*.yaml"))):
        name = os.path.basename(f)[: -len(".yaml")]
        d = yaml.safe_load(open(f))
        turbines[name] = {
            "hub_height": d["HUB_HEIGHT"],
            "V": [float(v) for v in d["V"]],
            "POW": [float(p) for p in d["POW"]],
            "meta": {k: d[k] for k in ("name", "manufacturer", "source") if k in d},
        }
    panels = {}
    for f in sorted(glob.glob(os.path.join(REF, "solarpanel", "*.yaml"))):
        name = os.path.basename(f)[: -len(".yaml")]
        panels[name] = yaml.safe_load(open(f))
    csp = {}
    for f in sorted(glob.glob(os.path.join(REF, "cspinstallation", "*.yaml"))):
        name = os.path.basename(f)[: -len(".yaml")]
        d = yaml.safe_load(open(f))
        eff = d["efficiency"]
        keys = sorted(eff["altitude"].keys())
        alt = sorted({float(eff["altitude"][k]) for k in keys})
        az = sorted({float(eff["azimuth"][k]) for k in keys})
        table = [[None] * len(az) for _ in alt]
        for k in keys:
            table[alt.index(float(eff["altitude"][k]))][az.index(float(eff["azimuth"][k]))] = float(eff["value"][k])
        csp[name] = {"technology": d.get("technology"), "r_irradiance": d["r_irradiance"],
                     "altitude_deg": alt, "azimuth_deg": az, "efficiency_percent": table,
                     "meta": {k: d[k] for k in ("name", "source") if k in d}}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "csp.json"), "w") as fh:
        json.dump(csp, fh, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "turbines.json"), "w") as fh:
        json.dump(turbines, fh, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "panels.json"), "w") as fh:
        json.dump(panels, fh, indent=1, sort_keys=True)
    print(f"{len(turbines)} turbines, {len(panels)} panels, {len(csp)} csp installations -> {os.path.abspath(OUT)}")


if __name__ == "__main__":
    main()
