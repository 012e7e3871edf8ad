"""Regenerate tests/golden/kat1_ablaincourt.json from the reference's own stored notebook output.

Run in the build container only (needs /root/reference):  python tools/make_golden.py
Source: /root/reference/examples/demo.ipynb, the cell `observation = env.reset(); print(observation)` (notebook JSON lines
137-138): the OrderedDict printed for `Ablaincourt_Floris` holds the sampled free-stream wind and the 7 local wind speeds /
directions FLORIS 3.5 computed.  This is the only known-answer vector the reference holds for the hot path.
"""
import json
import os
import re
import sys

NB = "/root/reference/examples/demo.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "kat1_ablaincourt.json")


def _floats(text):
    return [float(v) for v in re.findall(r"-?\d+\.\d*(?:e-?\d+)?|-?\d+\.(?!\d)", text)]


def main():
    if not os.path.exists(NB):
        sys.exit("reference tree not present; the committed fixture stands")
    nb = json.load(open(NB))
    for cell in nb["cells"]:
        src = "".join(cell.get("source", []))
        if cell["cell_type"] == "code" and "env.reset()" in src and "print(observation)" in src:
            text = "".join("".join(o.get("text", [])) for o in cell.get("outputs", []))
            break
    else:
        sys.exit("reset cell not found")
    parts = dict(re.findall(r"\('(\w+)', array\(\[(.*?)\]\)\)", text, flags=re.S))
    free = _floats(parts["freewind_measurements"])
    kat = {
        "source": "ifpen/wfcrl-env examples/demo.ipynb:137-138 (stored output of env.reset() for Ablaincourt_Floris)",
        "env_id": "Ablaincourt_Floris",
        "wind_speed": free[0],
        "wind_direction": free[1],
        "yaw": _floats(parts["yaw"]),
        "local_wind_speed": _floats(parts["wind_speed"]),
        "local_wind_direction": _floats(parts["wind_direction"]),
        "print_resolution": 1e-8,
    }
    assert len(kat["local_wind_speed"]) == 7 and len(kat["local_wind_direction"]) == 7 and len(kat["yaw"]) == 7
    with open(OUT, "w") as fp:
        json.dump(kat, fp, indent=1)
    print("wrote", OUT)
    print(kat)


if __name__ == "__main__":
    main()
