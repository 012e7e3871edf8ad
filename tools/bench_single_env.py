"""BASELINE.json configs[0] / BASELINE.md section 3 protocol: ONE Gymnasium env (Turb6_Row2_Floris), reset at (8 m/s, 270 deg),
5 warm-up steps, 100 timed steps with U(-5,5) float32 yaw actions (seed 0).  Times the drop-in `envs.make` path of this
repo (one kernel launch + host round trip per step) next to the numpy oracle port of the reference step."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import env_oracle
from wfcrl_b200 import environments as envs
from wfcrl_b200.layouts import get_layout


def protocol(env, T, steps=100, warmup=5):
    rng = np.random.default_rng(0)
    env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    for _ in range(warmup):
        env.step({"yaw": rng.uniform(-5, 5, T).astype(np.float32)})
    t0 = time.perf_counter()
    total = 0.0
    for _ in range(steps):
        _o, r, _t, _tr, _i = env.step({"yaw": rng.uniform(-5, 5, T).astype(np.float32)})
        total += float(r[0])
    return steps / (time.perf_counter() - t0), total


rows = {}
for env_id, key in (("Turb6_Row2_Floris", "Turb6_Row2_"), ("Ablaincourt_Floris", "Ablaincourt_"), ("HornsRev1_Floris", "HornsRev1_")):
    case = get_layout(key)
    T = case["num_turbines"]
    ours = envs.make(env_id, log=False, max_num_steps=500)
    rate, tot = protocol(ours, T)
    ref = env_oracle.EnvOracle(case["xcoords"], case["ycoords"], max_num_steps=500)
    rrate, rtot = protocol(ref, T, steps=20 if T > 40 else 100)
    rows[env_id] = {"turbines": T, "ours_steps_per_s": rate, "oracle_port_steps_per_s": rrate, "speedup": rate / rrate,
                    "return_ours": tot}
    print(env_id, rows[env_id], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/single_env.json", "w"), indent=1)
