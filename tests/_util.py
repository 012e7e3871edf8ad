"""Shared helpers for the parity tests (inputs follow SURVEY.md section 8d)."""
import numpy as np

from wfcrl_b200.layouts import get_layout

CONFIG_LAYOUTS = ["Turb6_Row2_", "Ablaincourt_", "Turb_TCRWP_", "Turb16_TCRWP_", "Turb32_Row5_", "HornsRev1_"]


def sample_winds(n, seed=0, tie_every=0):
    """ws = clip(8*Weibull(8), 3, 28), wd = N(270, 20) % 360: the reset distribution of wfcrl/mdp.py:242-258."""
    rng = np.random.default_rng(seed)
    ws = np.clip(8 * rng.weibull(8, n), 3, 28)
    wd = np.clip(rng.normal(270, 20, n) % 360, 0, 360)
    if tie_every:
        wd[::tie_every] = 270.0  # exact x-ties for the row layouts
    return ws, wd


def host_trig(wd):
    dev = (((np.asarray(wd) % 360.0) - 270.0) % 360.0 + 360.0) % 360.0
    return np.cos(np.radians(dev)), np.sin(np.radians(dev))


def layout(name):
    c = get_layout(name)
    return np.asarray(c["xcoords"], dtype=np.float64), np.asarray(c["ycoords"], dtype=np.float64)


def rel_err(a, ref, floor):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor))


# ---- the two episodes of the reference's examples/demo.ipynb (cells 12-13 and 23-24), for any env with the reference API
def notebook_single_agent_episode(env, options):
    """`step_policy(i)`: turbine int(i/5 % T) moves -5 deg when i % 5 == 0.  Returns (total reward, farm power per step)."""
    env.reset(options=options)
    total, i, done, power = 0.0, 0, False, []
    while not done:
        action = {"yaw": np.zeros(env.num_turbines)}
        if i % 5 == 0:
            action["yaw"][int(i / 5 % env.num_turbines)] = -5.0
        _obs, reward, term, trunc, info = env.step(action)
        total += float(np.asarray(reward).reshape(-1)[0])
        power.append(float(np.sum(info["power"])))
        i += 1
        done = term or trunc
    return total, np.array(power)


def notebook_multi_agent_episode(env, options):
    """`multi_agent_step_routine` with `step_policy(i, j)`: agent j moves -5 deg when its step count i % (4 (j+1)) == 0.
    Returns (total reward of the first agent -- all agents get the same --, farm power per physical step)."""
    env.reset(options=options)
    totals = {a: 0.0 for a in env.possible_agents}
    done = {a: False for a in env.possible_agents}
    steps = {a: 0 for a in env.possible_agents}
    first, power = env.possible_agents[0], []
    for agent in env.agent_iter():
        _obs, reward, term, trunc, info = env.last()
        done[agent] = done[agent] or term or trunc
        totals[agent] += float(np.asarray(reward, dtype=np.float64).reshape(-1)[0])
        if agent == first and "power" in info:
            power.append(float(sum(env.infos[a]["power"] for a in env.possible_agents)))
        if done[agent]:
            action = None
        else:
            j = env.agent_name_mapping[agent]
            action = {"yaw": np.array([-5.0])} if steps[agent] % (4 * (j + 1)) == 0 else {"yaw": np.zeros(1)}
            steps[agent] += 1
        env.step(action)
    assert len(set(round(v, 9) for v in totals.values())) == 1
    return totals[first], np.array(power)


def plateau_inner(k, plateau_len):
    """Iterations strictly inside constant-power plateau k of a notebook figure (the first and last iteration of a
    plateau sit on the corners of the drawn polyline and are not digitised)."""
    return [i for i in range(plateau_len * k + 1, plateau_len * (k + 1) - 1) if i <= 67]


def plateau_means(power, plateau_len, n):
    return np.array([np.asarray(power)[plateau_inner(k, plateau_len)].mean() for k in range(n)])
