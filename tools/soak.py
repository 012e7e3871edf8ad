"""Soak run: thousands of env steps with in-loop resets over every kernel flavour; every output must stay finite and
inside its physical range, the non-finite guard counter must stay at zero."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from wfcrl_b200 import environments as envs

rows = []
for env_id, B, precision, steps, max_steps in (("HornsRev1_Floris", 8192, "f32", 2000, 500), ("Dec_Ablaincourt_Floris", 16384, "f32", 3000, 100),
                                                ("Turb32_Row5_Floris", 4096, "f64", 600, 150), ("HornsRev2_Floris", 2048, "f32", 1000, 200)):
    env = envs.make_vec(env_id, B, precision=precision, max_num_steps=max_steps)
    obs = env.reset(seed=123)
    T = env.num_turbines
    gen = torch.Generator(device="cuda").manual_seed(5)
    bad, lo_p, hi_p, lo_ws, t0 = 0, 1e9, -1e9, 1e9, time.perf_counter()
    for k in range(steps):
        a = torch.rand(B, T, device="cuda", generator=gen) * 14 - 7       # beyond the +-5 action bound on purpose
        if k % 97 == 0:
            a[:] = 5.0                                                      # saturating pushes towards the yaw bound
        out = env.step(a)
        reward = out[1] if not isinstance(out[1], dict) else out[1]["turbine_1"]
        o = env.backend.out
        if k % 10 == 0:
            bad += int((~torch.isfinite(reward)).sum()) + int((~torch.isfinite(o["power"])).sum())
            lo_p, hi_p = min(lo_p, float(o["power"].min())), max(hi_p, float(o["power"].max()))
            lo_ws = min(lo_ws, float(o["wind_speed"].min()))
            assert float(o["yaw"].abs().max()) <= 40.0
    torch.cuda.synchronize()
    stats = env.episode_statistics()
    nonfinite = int(env.backend.get_state("nonfinite").sum())
    rows.append({"env": env_id, "envs": B, "precision": precision, "steps": steps, "episodes": stats["episodes"],
                 "length_mean": stats["length_mean"], "nonfinite_rewards": nonfinite, "nonfinite_sampled": bad,
                 "power_min_MW": lo_p, "power_max_MW": hi_p, "local_ws_min": lo_ws,
                 "env_steps_per_s": B * steps / (time.perf_counter() - t0)})
    print(rows[-1], flush=True)
    assert nonfinite == 0 and bad == 0 and lo_p >= 0.0 and hi_p <= 5.01 and stats["length_mean"] == max_steps - 1
    env.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/soak.json", "w"), indent=1)
