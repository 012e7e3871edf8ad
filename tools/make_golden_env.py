"""Pin the env semantics to the REFERENCE'S OWN CODE run here (build container only; needs /root/reference).

The reference's env layer -- wfcrl/{interface,mdp,simple_env,multiagent_env,rewards,wrappers}.py and
wfcrl/environments/{registration,data_cases}.py -- is pure Python and sits in /root/reference.  Its third-party imports
are absent from this container, so they are replaced in ``sys.modules`` by shims:

  * ``floris.tools.FlorisInterface``  -> the numpy oracle of the FLORIS 3.5 solve (oracle/floris_oracle.py), exposing the
    attributes the reference reads (interface.py:479, 551-567, 623, 632-647, 666);
  * ``gymnasium`` / ``pettingzoo``    -> the stand-ins of wfcrl_b200/spaces.py and wfcrl_b200/aec.py (gymnasium 0.29.1 /
    pettingzoo 1.24.3 semantics) plus thin ``Env`` / ``Wrapper`` / ``BaseWrapper`` classes;
  * ``mpi4py``, ``openfast_toolbox``  -> inert (FAST.Farm backend, out of scope).

``wfcrl.envs.make(...)`` is then driven UNMODIFIED through the scenarios below and every step's observation, reward,
flags and info is written to tests/golden/env_ref_*.json.  tests/test_env_golden*.py replay the same actions through
oracle/env_oracle.py, the single-env CUDA drop-ins and the batched kernels.

    python tools/make_golden_env.py
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import yaml

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------------------------------
# shims
# ----------------------------------------------------------------------------------------------------------------------
def install_shims():
    from oracle import floris_oracle
    from wfcrl_b200 import aec as _aec
    from wfcrl_b200 import spaces as _spaces

    assert not _spaces.HAVE_GYMNASIUM and not _aec.HAVE_PETTINGZOO, "real libraries present: use them directly"

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    # --- gymnasium ---------------------------------------------------------------------------------------------------
    class Env:
        metadata = {}

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):  # gymnasium.Wrapper: forwards attribute access and step/reset to the wrapped env
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

    sp = module("gymnasium.spaces", Box=_spaces.Box, Dict=_spaces.Dict, MultiDiscrete=_spaces.MultiDiscrete)
    reg = module("gymnasium.envs.registration", register=lambda **kw: None)
    envs_mod = module("gymnasium.envs", registration=reg)
    module("gymnasium", spaces=sp, Env=Env, Wrapper=Wrapper, envs=envs_mod)

    # --- pettingzoo --------------------------------------------------------------------------------------------------
    class BaseWrapper(_aec.AECEnv):  # pettingzoo.utils.wrappers.BaseWrapper 1.24.3
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") and name != "_cumulative_rewards":
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def observation_space(self, agent):
            return self.env.observation_space(agent)

        def action_space(self, agent):
            return self.env.action_space(agent)

        def observe(self, agent):
            return self.env.observe(agent)

        def state(self):
            return self.env.state()

        def step(self, action):
            self.env.step(action)

        def reset(self, seed=None, options=None):
            self.env.reset(seed=seed, options=options)

        def close(self):
            self.env.close()

    wr = module("pettingzoo.utils.wrappers", BaseWrapper=BaseWrapper)
    ut = module("pettingzoo.utils", agent_selector=_aec.agent_selector, wrappers=wr)
    module("pettingzoo", AECEnv=_aec.AECEnv, utils=ut)

    # --- mpi4py / openfast_toolbox (FAST.Farm side: never executed here) ---------------------------------------------
    class _Comm:
        pass

    module("mpi4py", MPI=types.SimpleNamespace(Comm=_Comm, COMM_WORLD=_Comm()))
    module("mpi4py.MPI", Comm=_Comm, COMM_WORLD=_Comm())
    noop = lambda *a, **k: None  # noqa: E731
    module("openfast_toolbox")
    module("openfast_toolbox.fastfarm", fastFarmBoxExtent=noop, fastFarmTurbSimExtent=noop, writeFastFarm=noop)
    module("openfast_toolbox.io")
    module("openfast_toolbox.io.fast_input_file", FASTInputFile=object)

    # --- floris: the oracle behind the attributes the reference touches ----------------------------------------------
    class _FlorisInterface:
        def __init__(self, configuration):
            with open(configuration) as fp:
                cfg = yaml.safe_load(fp)
            # the template must still be the model the oracle restates (case.yaml:14-16,27-60,84-89)
            ff, wk = cfg["flow_field"], cfg["wake"]
            assert cfg["solver"]["turbine_grid_points"] == floris_oracle.CASE["grid_points"]
            assert ff["air_density"] == floris_oracle.CASE["air_density"] and ff["wind_shear"] == floris_oracle.CASE["wind_shear"]
            assert ff["turbulence_intensity"] == floris_oracle.CASE["turbulence_intensity"] and ff["wind_veer"] == 0.0
            assert wk["model_strings"] == {"combination_model": "sosfs", "deflection_model": "gauss",
                                           "turbulence_model": "crespo_hernandez", "velocity_model": "gauss"}
            gd, gv = wk["wake_deflection_parameters"]["gauss"], wk["wake_velocity_parameters"]["gauss"]
            for key in ("ad", "alpha", "bd", "beta", "dm", "ka", "kb"):
                assert gd[key] == floris_oracle.CASE[key], key
            for key in ("alpha", "beta", "ka", "kb"):
                assert gv[key] == floris_oracle.CASE[key], key
            ch = wk["wake_turbulence_parameters"]["crespo_hernandez"]
            assert (ch["initial"], ch["constant"], ch["ai"], ch["downstream"]) == (0.1, 0.5, 0.8, -0.32)
            assert wk["enable_secondary_steering"] and wk["enable_yaw_added_recovery"] and wk["enable_transverse_velocities"]
            assert cfg["farm"]["turbine_type"] == ["nrel_5MW"]
            self.layout_x = [float(v) for v in cfg["farm"]["layout_x"]]
            self.layout_y = [float(v) for v in cfg["farm"]["layout_y"]]
            T = len(self.layout_x)
            flow = types.SimpleNamespace(wind_speeds=np.array([float(ff["wind_speeds"][0])]),
                                         wind_directions=np.array([float(ff["wind_directions"][0])]),
                                         u=None, v=None, w=None, turbulence_intensity_field=None)
            farm = types.SimpleNamespace(yaw_angles=np.zeros((1, 1, T)))
            self.floris = types.SimpleNamespace(flow_field=flow, farm=farm)
            self._power = None

        def reinitialize(self, wind_speeds=None, wind_directions=None):
            if wind_speeds is not None:
                self.floris.flow_field.wind_speeds = np.array(wind_speeds, dtype=np.float64)
            if wind_directions is not None:
                self.floris.flow_field.wind_directions = np.array(wind_directions, dtype=np.float64)

        def calculate_wake(self, yaw_angles=None):
            T = len(self.layout_x)
            yaw = np.zeros((1, 1, T)) if yaw_angles is None else np.asarray(yaw_angles, dtype=np.float64)
            ff = self.floris.flow_field
            sol = floris_oracle.solve(self.layout_x, self.layout_y, ff.wind_speeds[0], ff.wind_directions[0],
                                      yaw.reshape(T))
            ff.u, ff.v, ff.w = sol.u[None, None], sol.v[None, None], sol.w[None, None]
            ff.turbulence_intensity_field = sol.ti.reshape(1, 1, T, 1, 1)
            self.floris.farm.yaw_angles = yaw.copy()
            self._power = sol.power_W.reshape(1, 1, T)

        def get_turbine_powers(self):
            return self._power.copy()

    tools = module("floris.tools", FlorisInterface=_FlorisInterface)
    module("floris", tools=tools)


# ----------------------------------------------------------------------------------------------------------------------
# recording
# ----------------------------------------------------------------------------------------------------------------------
def _arr(a):
    a = np.asarray(a)
    return {"dtype": str(a.dtype), "shape": list(a.shape), "data": [float(v) for v in a.reshape(-1)]}


def _obs(obs):
    return {k: _arr(v) for k, v in obs.items()}


def record_single(env_id, make_kwargs, reset_kwargs, actions, np_seed=None):
    """Drive a centralised env (reference simple_env.py) through `actions`; one record per step."""
    from wfcrl import environments as envs

    if np_seed is not None:
        np.random.seed(np_seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = envs.make(env_id, **make_kwargs)
        obs = env.reset(**reset_kwargs)
    rec = {"reset_observation": _obs(obs), "steps": []}
    for a in actions:
        obs, reward, terminated, truncated, info = env.step({"yaw": np.array(a, copy=True)})
        rec["steps"].append({"action": _arr(a), "observation": _obs(obs), "reward": _arr(reward), "terminated": bool(terminated),
                             "truncated": bool(truncated), "power": _arr(info["power"]), "load": _arr(info["load"])})
        if truncated:
            break
    if hasattr(env, "history"):
        rec["history_lengths"] = {k: len(v) for k, v in env.history.items()}
    return rec


def record_multi(env_id, make_kwargs, reset_kwargs, policy, max_cycles):
    """Drive the AEC env (reference multiagent_env.py) with `policy(agent_index, step_of_agent) -> yaw action or None`
    through agent_iter(), recording what last() returns before every step, dead steps included."""
    from wfcrl import environments as envs

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = envs.make(env_id, **make_kwargs)
        env.reset(**reset_kwargs)
    rec = {"agents": list(env.possible_agents), "events": []}
    counts = {a: 0 for a in env.possible_agents}
    for agent in env.agent_iter(max_iter=max_cycles * len(env.possible_agents)):
        obs, reward, term, trunc, info = env.last()
        dead = bool(term or trunc)
        ev = {"agent": agent, "observation": _obs(obs), "cumulative_reward": _arr(reward), "terminated": bool(term),
              "truncated": bool(trunc), "info": {k: _arr(v) for k, v in info.items()}, "dead_step": dead}
        if dead:
            action = None
        else:
            j = env.agent_name_mapping[agent]
            action = policy(j, counts[agent])
            counts[agent] += 1
        ev["action"] = None if action is None else _arr(action)
        rec["events"].append(ev)
        env.step(None if action is None else {"yaw": np.array(action, copy=True)})
    rec["agents_left"] = list(env.agents)
    return rec


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; the committed fixtures stand")
    install_shims()
    sys.path.insert(0, REF)
    from wfcrl import rewards

    os.makedirs(OUT_DIR, exist_ok=True)
    workdir = tempfile.mkdtemp(prefix="wfcrl_golden_")  # FlorisInterface.from_case writes __simul__/ under the cwd
    os.chdir(workdir)
    out = {"_source": "generated by tools/make_golden_env.py from the UNMODIFIED /root/reference/wfcrl package "
                      "(interface.py, mdp.py, simple_env.py, multiagent_env.py, rewards.py, wrappers.py, "
                      "environments/) with FLORIS replaced by oracle/floris_oracle.py and gymnasium / pettingzoo by "
                      "stand-ins; floats are exact float64/float32 values", "scenarios": {}}
    sc = out["scenarios"]
    rng = np.random.default_rng(20261017)

    def uni(n, T, lo=-5, hi=5):
        return rng.uniform(lo, hi, (n, T)).astype(np.float32)

    # 1. random continuous actions, fixed wind (BASELINE configs[0] protocol), truncation at max_num_steps - 1
    sc["single_continuous_turb6"] = dict(
        env_id="Turb6_Row2_Floris", make_kwargs={"max_num_steps": 30}, reset_kwargs={"options": {"wind_speed": 8.0, "wind_direction": 270.0}})
    sc["single_continuous_turb6"]["record"] = record_single("Turb6_Row2_Floris", {"max_num_steps": 30},
                                                           {"options": {"wind_speed": 8.0, "wind_direction": 270.0}}, uni(40, 6))
    # 2. seeded reset (sampled wind), out-of-range actions (clipped to +-5), long enough to hit the actuation constraint
    sc["single_seeded_ablaincourt"] = dict(env_id="Ablaincourt_Floris", make_kwargs={"max_num_steps": 40}, reset_kwargs={"seed": 7})
    sc["single_seeded_ablaincourt"]["record"] = record_single("Ablaincourt_Floris", {"max_num_steps": 40}, {"seed": 7},
                                                             uni(45, 7, -9, 9))
    # 3. saturation at the yaw bounds: always +5 on even turbines, -5 on odd ones, narrow bounds so they are reached
    sat = np.tile(np.where(np.arange(6) % 2 == 0, 5.0, -5.0).astype(np.float32), (25, 1))
    sc["single_saturation_turb6"] = dict(env_id="Turb6_Row2_Floris", make_kwargs={"max_num_steps": 25, "controls": {"yaw": (-12, 8, 5)}},
                                         reset_kwargs={"options": {"wind_speed": 9.5, "wind_direction": 262.0}})
    sc["single_saturation_turb6"]["record"] = record_single(
        "Turb6_Row2_Floris", {"max_num_steps": 25, "controls": {"yaw": (-12, 8, 5)}},
        {"options": {"wind_speed": 9.5, "wind_direction": 262.0}}, sat)
    # 4. discrete control {0, 1, 2} -> (a - 1) * step
    disc = rng.integers(0, 3, (30, 3)).astype(np.float32)
    sc["single_discrete_turb3"] = dict(env_id="Turb3_Row1_Floris", make_kwargs={"max_num_steps": 28, "continuous_control": False, "controls": {"yaw": (-20, 20, 2)}},
                                       reset_kwargs={"options": {"wind_speed": 7.3, "wind_direction": 271.5}})
    sc["single_discrete_turb3"]["record"] = record_single(
        "Turb3_Row1_Floris", {"max_num_steps": 28, "continuous_control": False, "controls": {"yaw": (-20, 20, 2)}},
        {"options": {"wind_speed": 7.3, "wind_direction": 271.5}}, disc)
    # 5. / 6. reward shapers
    sc["single_reference_pct_turb3"] = dict(env_id="Turb3_Row1_Floris", make_kwargs={"max_num_steps": 15, "reward_shaper": "ReferencePercentage(1.7)"},
                                            reset_kwargs={"options": {"wind_speed": 8.0, "wind_direction": 270.0}})
    sc["single_reference_pct_turb3"]["record"] = record_single(
        "Turb3_Row1_Floris", {"max_num_steps": 15, "reward_shaper": rewards.ReferencePercentage(1.7)},
        {"options": {"wind_speed": 8.0, "wind_direction": 270.0}}, uni(14, 3))
    sc["single_step_pct_turb3"] = dict(env_id="Turb3_Row1_Floris", make_kwargs={"max_num_steps": 15, "reward_shaper": "StepPercentage(0.9)"},
                                       reset_kwargs={"options": {"wind_speed": 10.0, "wind_direction": 268.0}})
    sc["single_step_pct_turb3"]["record"] = record_single(
        "Turb3_Row1_Floris", {"max_num_steps": 15, "reward_shaper": rewards.StepPercentage(0.9)},
        {"options": {"wind_speed": 10.0, "wind_direction": 268.0}}, uni(14, 3))
    # 8. load_coef and log=False
    sc["single_load_coef_turb6"] = dict(env_id="Turb6_Row2_Floris", make_kwargs={"max_num_steps": 12, "load_coef": 0.7, "log": False},
                                        reset_kwargs={"options": {"wind_speed": 6.2, "wind_direction": 281.0}})
    sc["single_load_coef_turb6"]["record"] = record_single(
        "Turb6_Row2_Floris", {"max_num_steps": 12, "load_coef": 0.7, "log": False},
        {"options": {"wind_speed": 6.2, "wind_direction": 281.0}}, uni(11, 6))

    # 9. AEC cycle with the stale per-agent constraint and the dead steps at the end
    pol_rng = np.random.default_rng(99)
    table = pol_rng.uniform(-7, 7, (7, 64)).astype(np.float32)
    sc["multi_ablaincourt"] = dict(env_id="Dec_Ablaincourt_Floris", make_kwargs={"max_num_steps": 22},
                                   reset_kwargs={"options": {"wind_speed": 8.3, "wind_direction": 275.0}},
                                   policy_table=_arr(table))
    sc["multi_ablaincourt"]["record"] = record_multi(
        "Dec_Ablaincourt_Floris", {"max_num_steps": 22}, {"options": {"wind_speed": 8.3, "wind_direction": 275.0}},
        lambda j, k: table[j, k:k + 1], max_cycles=40)
    # 10. AEC with the StepPercentage shaper and discrete control
    dtable = pol_rng.integers(0, 3, (3, 64)).astype(np.float32)
    sc["multi_discrete_step_pct_turb3"] = dict(
        env_id="Dec_Turb3_Row1_Floris", make_kwargs={"max_num_steps": 12, "continuous_control": False, "reward_shaper": "StepPercentage()"},
        reset_kwargs={"seed": 3}, policy_table=_arr(dtable))
    sc["multi_discrete_step_pct_turb3"]["record"] = record_multi(
        "Dec_Turb3_Row1_Floris", {"max_num_steps": 12, "continuous_control": False, "reward_shaper": rewards.StepPercentage()},
        {"seed": 3}, lambda j, k: dtable[j, k:k + 1], max_cycles=30)

    # 7. time-series mode: the wind moves before every solve, reward normalised by the PREVIOUS state's wind.
    #    LAST on purpose: registration.py:94 writes `wind_time_series` into the registry's shared FarmCase object, so in the
    #    reference every later make() of the same layout in the same process silently stays in time-series mode.
    t = np.arange(40)
    series = np.stack([8.0 + 1.5 * np.sin(t / 5.0), 270.0 + 8.0 * np.cos(t / 7.0)], 1)
    sc["single_time_series_turb6"] = dict(env_id="Turb6_Row2_Floris", reset_kwargs={}, np_seed=11)
    # (the reference only works with a CSV path here: an ndarray trips `if self.wind_time_series and ...`, interface.py:589)
    csv_path = os.path.join(workdir, "wind_series.csv")
    with open(csv_path, "w") as fp:
        fp.write("speed,direction\n" + "\n".join(f"{a!r},{b!r}" for a, b in series.tolist()) + "\n")
    import pandas as pd

    # the doubles the reference actually sees (pandas' default float parser is not exactly round-trip)
    sc["single_time_series_turb6"]["make_kwargs"] = {"max_num_steps": 20, "wind_time_series": pd.read_csv(csv_path).values.tolist()}
    sc["single_time_series_turb6"]["record"] = record_single(
        "Turb6_Row2_Floris", {"max_num_steps": 20, "wind_time_series": csv_path}, {}, uni(19, 6), np_seed=11)
    for name, scen in sc.items():
        path = os.path.join(OUT_DIR, f"env_ref_{name}.json")
        with open(path, "w") as fp:
            json.dump({"_source": out["_source"], "name": name, **scen}, fp, separators=(",", ":"))
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
