"""Farm layouts of the reference's Floris cases (DATA extracted by tools/extract_layouts.py from
wfcrl/environments/data_cases.py:105-533) plus the procedural single-row farms (data_cases.py:501-519,
registration.py:23) and one documented alias."""
from __future__ import annotations

import json
import os
import re
from functools import lru_cache
from typing import Dict, List, Tuple

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "layouts.json")
_ROW_PATTERN = re.compile(r"Turb(\d+)_Row(\d+)_?$")


@lru_cache(maxsize=None)
def _load() -> Dict:
    with open(_DATA) as fp:
        return json.load(fp)


def named_layouts() -> List[str]:
    """Layout keys in the reference's registration order (registration.py:22-23)."""
    data = _load()
    names = list(data["named"].keys())
    names.extend(f"Turb{n}_Row1_" for n in range(1, data["row"]["max_turbines"] + 1))
    return names


# Extension: BASELINE.json names "Turb16_TCRWP_Floris", which the reference does not register
# (its only TCRWP key is Turb_TCRWP_ with 32 turbines).  The alias = the first 16 turbines of that layout.
ALIASES = {"Turb16_TCRWP_": ("Turb_TCRWP_", 16)}


def get_layout(name: str) -> Dict:
    """Return {num_turbines, xcoords, ycoords, dt, t_init, buffer_window} for a layout key such as ``HornsRev1_``."""
    if not name.endswith("_"):
        name += "_"
    data = _load()
    if name in data["named"]:
        return dict(data["named"][name])
    if name in ALIASES:
        base, n = ALIASES[name]
        case = dict(data["named"][base])
        case["xcoords"] = case["xcoords"][:n]
        case["ycoords"] = case["ycoords"][:n]
        case["num_turbines"] = n
        return case
    match = _ROW_PATTERN.match(name)
    if match and int(match.group(2)) == 1:
        n = int(match.group(1))
        row = data["row"]
        return {
            "num_turbines": n,
            "xcoords": [i * row["spacing"] for i in range(n)],
            "ycoords": [0.0 for _ in range(n)],
            "dt": row["dt"], "t_init": row["t_init"], "buffer_window": row["buffer_window"],
        }
    raise KeyError(f"unknown layout {name!r}")


def layout_xy(name: str) -> Tuple[List[float], List[float]]:
    case = get_layout(name)
    return case["xcoords"], case["ycoords"]
