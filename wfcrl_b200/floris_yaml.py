"""Reader of the FLORIS v3 input file the reference hands to its backend (``FlorisInterface(simul_file=...)``,
wfcrl/interface.py:462-479; the file is written from wfcrl/simulators/floris/inputs/template/case.yaml by
``create_floris_case``, wfcrl/simul_utils.py:34-48).

The CUDA path implements the model that template selects -- Gauss velocity deficit + Gauss deflection + Crespo-Hernandez
turbulence + SOSFS combination with secondary steering, yaw-added recovery and transverse velocities on a 3x3 rotor grid
(5x5 through the basic kernels) of ``nrel_5MW`` turbines.  This module turns such a file into the layout, the initial wind and the numeric overrides of
``WfConfig``; anything the kernels do not implement (another wake model, a disabled GCH term, another grid or turbine) is
refused loudly instead of being silently ignored.
"""
from __future__ import annotations

from typing import Any, Dict

SUPPORTED_MODELS = {"combination_model": "sosfs", "deflection_model": "gauss", "turbulence_model": "crespo_hernandez",
                    "velocity_model": "gauss"}
GCH_SWITCHES = ("enable_secondary_steering", "enable_yaw_added_recovery", "enable_transverse_velocities")


def _need(cond: bool, message: str):
    if not cond:
        raise ValueError(f"unsupported FLORIS input for the B200 backend: {message}")


def parse_floris_config(config: Dict[str, Any]) -> Dict[str, Any]:
    """``config``: the parsed YAML.  Returns ``{"xcoords", "ycoords", "wind_speed", "wind_direction", "overrides"}`` where
    ``overrides`` maps ``WfConfig`` field names to values (see include/wfcrl_b200.h)."""
    solver = config.get("solver", {})
    grid_points = int(solver.get("turbine_grid_points", 3))
    _need(solver.get("type", "turbine_grid") == "turbine_grid" and grid_points in (3, 5),
          "solver must be a turbine_grid with turbine_grid_points = 3 (tuned kernels) or 5 (basic kernels)")
    farm = config["farm"]
    xs, ys = [float(v) for v in farm["layout_x"]], [float(v) for v in farm["layout_y"]]
    _need(len(xs) == len(ys) and len(xs) >= 1, "layout_x and layout_y must have the same non-zero length")
    types = farm.get("turbine_type", ["nrel_5MW"])
    _need(all(t == "nrel_5MW" for t in types), f"turbine_type must be nrel_5MW for every turbine, got {types}")

    flow = config["flow_field"]
    speeds, directions = flow.get("wind_speeds", [8.0]), flow.get("wind_directions", [270.0])
    _need(len(speeds) == 1 and len(directions) == 1, "exactly one wind speed and one wind direction")
    _need(float(flow.get("reference_wind_height", -1)) in (-1.0, 90.0), "reference_wind_height must be the hub height (-1)")
    overrides = {"turbine_grid_points": grid_points, "air_density": float(flow.get("air_density", 1.225)),
                 "turbulence_intensity": float(flow.get("turbulence_intensity", 0.06)),
                 "wind_shear": float(flow.get("wind_shear", 0.12)), "wind_veer": float(flow.get("wind_veer", 0.0))}

    wake = config["wake"]
    for role, name in SUPPORTED_MODELS.items():
        got = wake["model_strings"].get(role)
        _need(got == name, f"{role} must be '{name}', got '{got}'")
    for switch in GCH_SWITCHES:
        _need(bool(wake.get(switch, False)), f"{switch} must be true (the kernels implement the full GCH model)")
    deflection = wake.get("wake_deflection_parameters", {}).get("gauss", {})
    velocity = wake.get("wake_velocity_parameters", {}).get("gauss", {})
    for key in ("alpha", "beta", "ka", "kb"):
        if key in deflection and key in velocity:
            _need(float(deflection[key]) == float(velocity[key]),
                  f"gauss deflection and velocity models must share {key} ({deflection[key]} vs {velocity[key]})")
        if key in deflection or key in velocity:
            overrides[key] = float(deflection.get(key, velocity.get(key)))
    for key in ("ad", "bd", "dm"):
        if key in deflection:
            overrides[key] = float(deflection[key])
    turbulence = wake.get("wake_turbulence_parameters", {}).get("crespo_hernandez", {})
    for key, field in (("initial", "ch_initial"), ("constant", "ch_constant"), ("ai", "ch_ai"), ("downstream", "ch_downstream")):
        if key in turbulence:
            overrides[field] = float(turbulence[key])
    return {"xcoords": xs, "ycoords": ys, "wind_speed": float(speeds[0]), "wind_direction": float(directions[0]),
            "overrides": overrides}


def load_floris_yaml(path: str) -> Dict[str, Any]:
    import yaml

    with open(path) as fp:
        return parse_floris_config(yaml.safe_load(fp))
