"""Extract the farm layouts (pure DATA: turbine coordinates, dt, t_init) of the reference's Floris cases into
``wfcrl_b200/data/layouts.json``.

Run in the build container only (needs /root/reference):  python tools/extract_layouts.py
Source of the data: /root/reference/wfcrl/environments/data_cases.py:105-533 (``named_cases_dictionary`` and the
procedural single-row generator ``FarmRowFloris`` :501-519).  The reference module is executed stand-alone (it only
imports ``dataclasses``/``typing``), nothing from it is copied as code.
"""
import json
import os
import runpy
import sys

REF = "/root/reference/wfcrl/environments/data_cases.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wfcrl_b200", "data", "layouts.json")


def main():
    if not os.path.exists(REF):
        sys.exit("reference tree not present; layouts.json is committed, nothing to do")
    ns = runpy.run_path(REF)
    cases = {}
    for key, (_ff_case, fl_case) in ns["named_cases_dictionary"].items():
        cases[key] = {
            "num_turbines": int(fl_case.num_turbines),
            "xcoords": [float(v) for v in fl_case.xcoords],
            "ycoords": [float(v) for v in fl_case.ycoords],
            "dt": int(fl_case.dt),
            "t_init": int(fl_case.t_init),
            "buffer_window": int(fl_case.buffer_window),
        }
    row = ns["FarmRowFloris"]
    out = {
        "_source": "ifpen/wfcrl-env wfcrl/environments/data_cases.py (Floris cases)",
        "named": cases,
        "row": {
            "spacing": float(row.get_xcoords(2)[1]),
            "dt": int(row.dt), "t_init": int(row.t_init), "buffer_window": int(row.buffer_window),
            "max_turbines": 12,  # registration.py:23 registers Turb{1..12}_Row1_
        },
    }
    with open(OUT, "w") as fp:
        json.dump(out, fp, indent=1)
    print("wrote", OUT, {k: v["num_turbines"] for k, v in cases.items()})


if __name__ == "__main__":
    main()
