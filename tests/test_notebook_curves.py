"""KAT-3: the reference's stored notebook figures and printed episode totals (examples/demo.ipynb cells 13/17 and 24/26).

Two Ablaincourt_Floris episodes computed by FLORIS 3.5 with YAWED turbines (single-agent env down to -10 deg, decentralised
env down to the -40 deg bound).  The fixture (tools/make_golden_curves.py) holds the farm-power plateau levels digitised from
the PNGs, the printed totals and the two FITTED unknowns per episode (wind speed, direction: the unseeded reset's draw is
not stored).  2 fitted numbers against 14 / 17 levels + an 11-digit total: all of them are reproduced at once only if the
yawed wake solve, the power law, the load proxies, the reward and the (multi-agent) transition logic agree with the
reference.  Tolerances: levels within a third of a pixel of the figure, totals within 1e-5 relative.
"""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, env_oracle
from tests._util import layout, notebook_multi_agent_episode, notebook_single_agent_episode, plateau_means

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat3_notebook_curves.json")))["episodes"]
RUNNERS = {"single_agent": notebook_single_agent_episode, "multi_agent": notebook_multi_agent_episode}


def _check(kind, total, power, level_px=1.0 / 3.0, total_rtol=1e-5):
    g = GOLD[kind]
    levels = np.array(g["plateau_levels_MW"])
    got = plateau_means(power, g["plateau_length"], len(levels))
    px = g["calibration"]["mw_per_pixel"]
    assert len(power) == 69
    assert np.max(np.abs(got - levels)) <= level_px * px, (np.max(np.abs(got - levels)) / px, "pixels")
    # the shape of the curve alone (increments between consecutive plateaus), independent of the axis offset
    assert np.max(np.abs(np.diff(got) - np.diff(levels))) <= 2 * level_px * px
    assert abs(total - g["printed_total_reward"]) <= total_rtol * g["printed_total_reward"]


@pytest.mark.parametrize("kind", ["single_agent", "multi_agent"])
def test_oracle_reproduces_notebook_power_curve_and_total(kind):
    lx, ly = layout("Ablaincourt_")
    cls = env_oracle.EnvOracle if kind == "single_agent" else env_oracle.MAEnvOracle
    env = cls(lx, ly, solver=c_oracle.solve, max_num_steps=70)
    g = GOLD[kind]
    total, power = RUNNERS[kind](env, {"wind_speed": g["fitted_wind_speed"], "wind_direction": g["fitted_wind_direction"]})
    _check(kind, total, power)
    assert abs(total - g["oracle_total_reward_at_fit"]) < 1e-9 * total   # the fixture records what the oracle gave


def test_fit_is_sharp():
    """The fit is not a free lunch: half a degree / one percent away the levels miss by many pixels."""
    lx, ly = layout("Ablaincourt_")
    g = GOLD["single_agent"]
    levels, px = np.array(g["plateau_levels_MW"]), g["calibration"]["mw_per_pixel"]
    for dws, dwd in ((0.0, 0.5), (0.0, -0.5), (0.08, 0.0)):
        env = env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=70)
        _t, power = notebook_single_agent_episode(env, {"wind_speed": g["fitted_wind_speed"] + dws,
                                                        "wind_direction": g["fitted_wind_direction"] + dwd})
        got = plateau_means(power, 5, len(levels))
        assert np.max(np.abs(np.diff(got) - np.diff(levels))) > 5 * px or np.max(np.abs(got - levels)) > 20 * px


@pytest.mark.gpu
@pytest.mark.parametrize("kind,env_id", [("single_agent", "Ablaincourt_Floris"), ("multi_agent", "Dec_Ablaincourt_Floris")])
def test_cuda_envs_reproduce_notebook_power_curve_and_total(cuda_device, kind, env_id):
    """The drop-in `envs.make(...)` objects on the CUDA path (FP64 kernel), same episodes, same tolerances."""
    from wfcrl_b200 import environments as envs

    g = GOLD[kind]
    env = envs.make(env_id, max_num_steps=70)
    total, power = RUNNERS[kind](env, {"wind_speed": g["fitted_wind_speed"], "wind_direction": g["fitted_wind_direction"]})
    _check(kind, total, power)
    assert abs(total - g["oracle_total_reward_at_fit"]) < 1e-9 * total


@pytest.mark.gpu
def test_cuda_fp32_batch_reproduces_notebook_single_agent_curve(cuda_device):
    """The FP32 throughput kernel through the batched env: same figure, same tolerances (FP32 bar: 1e-4 relative)."""
    import torch

    from wfcrl_b200 import environments as envs

    g = GOLD["single_agent"]
    env = envs.make_vec("Ablaincourt_Floris", 2, precision="f32", max_num_steps=70, auto_reset=False)
    env.reset(options={"wind_speed": g["fitted_wind_speed"], "wind_direction": g["fitted_wind_direction"]})
    total, power = 0.0, []
    for i in range(69):
        a = torch.zeros(2, 7, device="cuda")
        if i % 5 == 0:
            a[:, int(i / 5 % 7)] = -5.0
        _obs, reward, _term, trunc, info = env.step(a)
        total += float(reward[0])
        power.append(float(info["power"][0].sum()))
    assert bool(trunc.all())
    _check("single_agent", total, np.array(power), total_rtol=1e-4)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision,total_rtol", [("f64", 1e-5), ("f32", 1e-4)])
def test_cuda_batched_multi_agent_env_reproduces_notebook_curve(cuda_device, precision, total_rtol):
    """The decentralised episode (yaw down to the -40 deg bound) through the BATCHED multi-agent env: one call per agent
    cycle, the stale per-agent actuation constraint evaluated inside the kernel."""
    import torch

    from wfcrl_b200 import environments as envs

    g = GOLD["multi_agent"]
    env = envs.make_vec("Dec_Ablaincourt_Floris", 3, precision=precision, max_num_steps=70, auto_reset=False)
    env.reset(options={"wind_speed": g["fitted_wind_speed"], "wind_direction": g["fitted_wind_direction"]})
    total, power = 0.0, []
    for cycle in range(69):
        a = torch.zeros(3, 7, device="cuda")
        for j in range(7):
            if cycle % (4 * (j + 1)) == 0:
                a[:, j] = -5.0
        _obs, reward, _term, trunc, info = env.step(a)
        total += float(reward["turbine_1"][1])
        power.append(float(sum(info[agent]["power"][1] for agent in env.possible_agents)))
    assert bool(trunc["turbine_7"].all())
    _check("multi_agent", total, np.array(power), total_rtol=total_rtol)
    env.close()
