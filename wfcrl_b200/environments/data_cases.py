"""Farm cases for the Floris-backed environments.

Mirrors the information content of the reference's ``FarmCase`` / ``FlorisCase`` dataclasses
(wfcrl/environments/data_cases.py:26-102): number of turbines, coordinates, ``dt``, ``t_init``, ``max_iter``, the
``set_wind_*`` flags (both False for Floris cases, :86-87) and ``simul_params`` (direction 270, speed 8, :95-102).
Coordinates come from ``wfcrl_b200/data/layouts.json`` (see wfcrl_b200/layouts.py).  FAST.Farm cases are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from ..layouts import get_layout, named_layouts


@dataclass
class DefaultControl:
    yaw: tuple = (-40, 40, 5)
    pitch: tuple = (0, 45, 1)
    torque: tuple = (-2e4, 2e4, 1e3)


@dataclass
class FarmCase:
    num_turbines: int
    xcoords: List[float]
    ycoords: List[float]
    dt: int
    buffer_window: int = 300
    t_init: int = 300
    max_iter: int = 100
    set_wind_speed: bool = False
    set_wind_direction: bool = False
    wind_time_series: Any = None

    @property
    def interface_kwargs(self) -> Optional[Dict]:
        return None

    def dict(self):
        return self.interface_kwargs

    def __repr__(self):
        text = f"Wind farm simulation on {getattr(self, 'simulator', '?')}: "
        text += f"{self.num_turbines} turbines - {self.max_iter} timesteps\n"
        for arg, val in (self.interface_kwargs or {}).items():
            text += f"{arg}: {val}\n"
        return text


@dataclass(repr=False)
class FlorisCase(FarmCase):
    simulator: str = field(default="Floris", init=False)

    @property
    def simul_params(self) -> Dict:
        return {
            "xcoords": self.xcoords,
            "ycoords": self.ycoords,
            "direction": 270,
            "speed": 8,
            "wind_time_series": self.wind_time_series,
        }

    @property
    def interface_kwargs(self) -> Dict:
        return self.simul_params


def floris_case(layout_key: str) -> FlorisCase:
    """A fresh FlorisCase for a layout key such as ``"HornsRev1_"`` (a new object per call: ``make`` mutates it)."""
    c = get_layout(layout_key)
    return FlorisCase(num_turbines=c["num_turbines"], xcoords=list(c["xcoords"]), ycoords=list(c["ycoords"]),
                      dt=c["dt"], buffer_window=c["buffer_window"], t_init=c["t_init"])


def registered_layouts() -> List[str]:
    return named_layouts()
