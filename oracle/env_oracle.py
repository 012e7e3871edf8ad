"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the WFCRL env semantics wrapped around the wake solve.

Single-env, pure Python/numpy restatement of what the reference does per ``*_Floris`` env step, in the reference's
own order of operations and numpy dtypes (float32 state/actions, float64 measures):

* ``InterfaceOracle``   <- ``FlorisInterface``            wfcrl/interface.py:444-671
* ``MDPOracle``         <- ``WindFarmMDP``                 wfcrl/mdp.py:19-319
* ``EnvOracle``         <- ``WindFarmEnv``                 wfcrl/simple_env.py:13-99
* ``MAEnvOracle``       <- ``MAWindFarmEnv`` (AEC)         wfcrl/multiagent_env.py:15-257 (+ the PettingZoo 1.24.3 AECEnv
                            bookkeeping it inherits: ``last``, ``agent_iter``, ``_was_dead_step``, ``_clear_rewards``,
                            ``_accumulate_rewards``, ``agent_selector``)
* reward shapers        <- wfcrl/rewards.py:16-46

gymnasium / pettingzoo are not importable here, so spaces are represented by their float32 low/high arrays only.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import this module.  Parity status: the
env logic below is pinned by the reference notebook's behavioural outputs (spaces ``examples/demo.ipynb:98-99``; 69
history rows for ``max_num_steps=70`` and the yaw trajectory ``:312-316``); rewards and the multi-agent transition are
pinned by the notebook's printed episode totals and power figures at figure resolution (KAT-3, see floris_oracle.py).
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import numpy as np

from . import floris_oracle

ACTUATORS_RATE = {"yaw": 0.3, "pitch": 8}       # mdp.py:52
DEFAULT_BOUNDS = {"wind_speed": [3, 28], "wind_direction": [0, 360], "yaw": [-40, 40]}  # mdp.py:45-51
DEFAULT_YAW_CONTROL = (-40, 40, 5)               # data_cases.py:21


# ---- reward shapers (rewards.py) --------------------------------------------------------------------------------
class DoNothing:
    def __call__(self, reward):
        return reward

    def reset(self):
        pass


class ReferencePercentage:
    def __init__(self, reference):
        self.reference = reference

    def __call__(self, reward):
        return (reward - self.reference) / self.reference

    def reset(self):
        pass


class StepPercentage:
    def __init__(self, reference=0.0):
        self.reference = reference

    def __call__(self, reward):
        shaped = 0.0 if self.reference == 0 else (reward - self.reference) / self.reference
        self.reference = reward
        return shaped

    def reset(self, reference=0.0):
        self.reference = reference


# ---- FlorisInterface ------------------------------------------------------------------------------------------------
class InterfaceOracle:
    CONTROL_SET = ["yaw"]
    measure_map = {"yaw": 0, "wind_speed": 1, "wind_direction": 2, "load": [3, 4, 5, 6], "freewind_measurements": None}

    def __init__(self, layout_x, layout_y, max_iter, wind_speed=8.0, wind_direction=270.0, wind_time_series=None,
                 solver=None, rng_start=None):
        self.layout_x, self.layout_y = list(layout_x), list(layout_y)
        self.num_turbines = len(self.layout_x)
        self.max_iter = max_iter
        self.solver = solver or floris_oracle.solve
        self.wind_time_series = wind_time_series
        self._rng_start = rng_start
        self.wind_speed, self.wind_dir = 8.0, 270.0   # FlorisCase.simul_params (data_cases.py:99-100)
        self._gen = self._make_gen(wind_speed, wind_direction, wind_time_series)
        self.init(*next(self._gen))

    def _make_gen(self, ws, wd, series):  # interface.py:503-524
        if series is None:
            def gen():
                while True:
                    yield ws, wd
        else:
            series = np.asarray(series)
            start = np.random.randint(0, series.shape[0]) if self._rng_start is None else self._rng_start
            series = np.r_[series[start:], series[:start]]

            def gen():
                for row in series:
                    yield row
        return gen()

    def update_wind(self, ws, wd):  # interface.py:663-671
        self.wind_speed, self.wind_dir = ws, wd % 360

    def init(self, wind_speed=None, wind_direction=None):  # interface.py:588-613
        if self.wind_time_series is not None:
            wind_speed = wind_direction = None
        self._gen = self._make_gen(wind_speed, wind_direction, self.wind_time_series)
        self.update_wind(*next(self._gen))
        self._num_iter = 0
        self._yaw_cmd = np.zeros(self.num_turbines)
        self.current_measures = np.zeros((self.num_turbines, 7)) * np.nan
        self._sol = None

    def update_command(self, yaw=None):  # interface.py:557-586
        if yaw is not None:
            self._yaw_cmd[:] = yaw.astype(np.double)
        self.update_wind(*next(self._gen))
        sol = self.solver(self.layout_x, self.layout_y, self.wind_speed, self.wind_dir, self._yaw_cmd)
        self._sol = sol
        self.current_measures[:, 0] = self._yaw_cmd
        self.current_measures[:, [1, 2]] = np.array([sol.ws_local, sol.wd_local]).T
        self.current_measures[:, [3, 4, 5, 6]] = np.array([sol.ti, sol.std_u, sol.std_v, sol.std_w]).T * 1e7
        self._num_iter += 1
        return self._num_iter == self.max_iter

    def avg_powers(self):
        return self._sol.power_W.flatten()

    def avg_wind(self):
        return np.array([self.wind_speed, self.wind_dir]).squeeze()

    def get_measure(self, measure):  # interface.py:650-655
        if measure not in self.measure_map:
            return None
        if measure == "freewind_measurements":
            return self.avg_wind()
        return self.current_measures[:, self.measure_map[measure]]


# ---- WindFarmMDP ----------------------------------------------------------------------------------------------------
def _clip_to_space(element, low, high):
    for name, value in element.items():
        element[name] = np.clip(value, low[name], high[name])
    return element


class MDPOracle:
    POSSIBLE_STATE_ATTRIBUTES = ["freewind_measurements", "wind_speed", "wind_direction", "yaw", "pitch", "torque"]

    def __init__(self, interface, controls=None, continuous_control=True, start_iter=0):
        self.interface = interface
        self.num_turbines = T = interface.num_turbines
        self.controls = dict(controls or {"yaw": DEFAULT_YAW_CONTROL})
        for name, b in list(self.controls.items()):  # mdp.py:174-211
            if name not in ("yaw",):
                raise ValueError(f"Cannot control `{name}`")
            if not (2 <= len(b) <= 3):
                raise TypeError("bounds")
            if not b[0] < b[1]:
                raise ValueError("lower_bound < upper_bound")
            if len(b) == 2:
                self.controls[name] = tuple(b) + (1,)
        self.continuous_control = continuous_control
        self.start_iter = start_iter
        self.measures = [o for o in self.POSSIBLE_STATE_ATTRIBUTES
                         if o not in self.controls and o in interface.measure_map]
        self.state_attributes = list(self.controls.keys()) + self.measures
        ones = np.ones(T, dtype=np.float32)
        # gymnasium.spaces.Box stores float32 bounds (mdp.py:108-153)
        self.action_low = {n: np.full(T, -b[2], dtype=np.float32) for n, b in self.controls.items()}
        self.action_high = {n: np.full(T, b[2], dtype=np.float32) for n, b in self.controls.items()}
        self.low, self.high = {}, {}
        for attr in self.state_attributes:
            if attr == "freewind_measurements":
                lo = np.array([DEFAULT_BOUNDS["wind_speed"][0], DEFAULT_BOUNDS["wind_direction"][0]], dtype=np.float32)
                hi = np.array([DEFAULT_BOUNDS["wind_speed"][1], DEFAULT_BOUNDS["wind_direction"][1]], dtype=np.float32)
            elif attr in self.controls:
                lo, hi = ones * self.controls[attr][0], ones * self.controls[attr][1]
            else:
                lo, hi = ones * DEFAULT_BOUNDS[attr][0], ones * DEFAULT_BOUNDS[attr][1]
            self.low[attr], self.high[attr] = lo.astype(np.float32), hi.astype(np.float32)
        self.start_state = None
        self._acc = {c: np.zeros(T, dtype=np.float32) for c in self.controls}

    def get_accumulated_actions(self):
        return self._acc.copy()  # shallow: the arrays are shared (mdp.py:165-166)

    def reset(self, seed=None, options=None):  # mdp.py:233-271
        rng = np.random.default_rng(seed)
        if options is not None and "wind_speed" in options:
            ws = options["wind_speed"]
        else:
            ws = np.clip(8 * rng.weibull(8), self.low["freewind_measurements"][0], self.high["freewind_measurements"][0])
        if options is not None and "wind_direction" in options:
            wd = options["wind_direction"]
        else:
            wd = np.clip(rng.normal(270, 20) % 360, self.low["freewind_measurements"][1],
                         self.high["freewind_measurements"][1])
        self.interface.init(ws, wd)
        for _ in range(self.start_iter + 1):
            self.interface.update_command()
        start = OrderedDict({a: self.interface.get_measure(a) for a in self.state_attributes})
        self.start_state = _clip_to_space(start, self.low, self.high)
        self._acc = {c: np.zeros(self.num_turbines, dtype=np.float32) for c in self.controls}
        return self.start_state

    def transition(self, state, joint_action):  # mdp.py:291-319
        state = _clip_to_space(OrderedDict((k, v.astype(np.float32)) for k, v in state.items()), self.low, self.high)
        nxt = copy.deepcopy(state)
        for control, a in joint_action.items():
            assert control in self.controls
            a = np.array(a, np.float32)
            if self.continuous_control:
                a = np.clip(a, self.action_low[control], self.action_high[control])
            else:
                a = (a - 1) * self.controls[control][-1]
            nxt[control] = np.clip(state[control] + a, self.low[control], self.high[control])
            self._acc[control] += np.abs(a)
        return nxt

    def step_interface(self, state):  # mdp.py:273-284
        done = self.interface.update_command(**{c: state[c] for c in self.controls})
        powers = self.interface.avg_powers()
        for m in self.measures:
            state[m] = self.interface.get_measure(m)
        loads = self.interface.get_measure("load")
        if loads is not None:
            loads /= 1e7
        return state, powers / 1e6, loads, done

    def take_action(self, state, joint_action):
        return self.step_interface(self.transition(state, joint_action))


# ---- WindFarmEnv (Gymnasium) ----------------------------------------------------------------------------------------
class EnvOracle:
    def __init__(self, layout_x, layout_y, controls=None, continuous_control=True, reward_shaper=None, start_iter=0,
                 max_num_steps=500, load_coef=0.1, dt=60, solver=None, wind_time_series=None, rng_start=None):
        iface = InterfaceOracle(layout_x, layout_y, max_iter=start_iter + max_num_steps, solver=solver,
                                wind_time_series=wind_time_series, rng_start=rng_start)
        self.mdp = MDPOracle(iface, controls, continuous_control, start_iter)
        self.num_turbines = self.mdp.num_turbines
        self.reward_shaper = reward_shaper or DoNothing()
        self.load_coef = load_coef
        self.dt = dt
        self._state = None
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        self.num_moves = 0

    def reset(self, seed=None, options=None):  # simple_env.py:49-56 (returns the observation ONLY)
        self.mdp.reset(seed, options)
        self._state = self.mdp.start_state
        self.reward_shaper.reset()
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        self.num_moves = 0
        return copy.deepcopy(self._state)

    def step(self, actions):  # simple_env.py:58-96
        assert self._state is not None
        self.num_moves += 1
        for control in actions:
            if control not in ACTUATORS_RATE:
                continue
            actuating_time = self.accumulated_actions[control] / ACTUATORS_RATE[control]
            actuating_frac = actuating_time / self.num_moves / self.dt
            actions[control][actuating_frac >= 0.1] = 0.0
        next_state, powers, loads, truncated = self.mdp.take_action(self._state, actions)
        normalized = powers * 1e3 / (self._state["freewind_measurements"][0] ** 3)
        load_penalty = np.mean(np.abs(loads)) if loads is not None else 0
        reward = np.array([self.reward_shaper(normalized.mean() - self.load_coef * load_penalty)])
        self._state = next_state
        info = {"power": powers, "load": loads}
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        return copy.deepcopy(self._state), reward, False, truncated, info


# ---- MAWindFarmEnv (PettingZoo AEC) ------------------------------------------------------------------------------------
class _AgentSelector:  # pettingzoo.utils.agent_selector
    def __init__(self, order):
        self.order = list(order)
        self._i = 0
        self.selected = None

    def next(self):
        self._i = (self._i + 1) % len(self.order)
        self.selected = self.order[self._i - 1]
        return self.selected

    def is_last(self):
        return self.selected == self.order[-1]


class MAEnvOracle:
    def __init__(self, layout_x, layout_y, controls=None, continuous_control=True, reward_shaper=None, start_iter=0,
                 max_num_steps=500, load_coef=0.1, dt=60, solver=None):
        iface = InterfaceOracle(layout_x, layout_y, max_iter=start_iter + max_num_steps, solver=solver)
        self.mdp = MDPOracle(iface, controls, continuous_control, start_iter)
        self.num_turbines = self.mdp.num_turbines
        self.reward_shaper = reward_shaper or DoNothing()
        self.load_coef = load_coef
        self.dt = dt
        self._state = None
        self.possible_agents = ["turbine_" + str(r + 1) for r in range(self.num_turbines)]
        self.agent_name_mapping = dict(zip(self.possible_agents, range(self.num_turbines)))

    def observe(self, agent):  # multiagent_env.py:102-115
        out = OrderedDict()
        for key, part in self._state.items():
            if key != "freewind_measurements":
                out[key] = part[self.agent_name_mapping[agent]]
        return out

    def reset(self, seed=None, options=None):  # multiagent_env.py:117-157
        self.mdp.reset(seed, options)
        self._state = self.mdp.start_state
        self.reward_shaper.reset()
        self.agents = self.possible_agents[:]
        self._num_steps = {a: 0 for a in self.agents}
        self.rewards = {a: np.array([0.0]) for a in self.agents}
        self._cumulative_rewards = {a: np.array([0.0]) for a in self.agents}
        self.terminations = {a: False for a in self.agents}
        self.truncations = {a: False for a in self.agents}
        self.infos = {a: {} for a in self.agents}
        self.actions = {a: None for a in self.agents}
        acc = self.mdp.get_accumulated_actions()
        self.accumulated_actions = {a: {c: acc[c][i] for c in acc} for i, a in enumerate(self.agents)}
        self.num_moves = 0
        self._sel = _AgentSelector(self.agents)
        self.agent_selection = self._sel.next()

    def last(self):  # AECEnv.last
        a = self.agent_selection
        return self.observe(a), self._cumulative_rewards[a], self.terminations[a], self.truncations[a], self.infos[a]

    def agent_iter(self, max_iter=2 ** 63):
        n = 0
        while self.agents and n < max_iter:
            n += 1
            yield self.agent_selection

    def _was_dead_step(self, action):  # AECEnv._was_dead_step (only the all-dead-at-once case occurs here)
        assert action is None
        agent = self.agent_selection
        for d in (self.rewards, self._cumulative_rewards, self.terminations, self.truncations, self.infos):
            del d[agent]
        self.agents.remove(agent)
        dead = [a for a in self.agents if self.terminations[a] or self.truncations[a]]
        if dead:
            self.agent_selection = dead[0]

    def step(self, action):  # multiagent_env.py:159-254
        agent = self.agent_selection
        if self.truncations[agent] or self.terminations[agent]:
            self._was_dead_step(action)
            return
        self._num_steps[agent] += 1
        for control in action:
            if control not in self.mdp.controls:
                raise ValueError(f"Control `{control}` for agent {agent} is not activated.")
        if any(c not in action for c in self.mdp.controls):
            raise ValueError(f"Action {action} for agent {agent} is incomplete.")
        acc = self.accumulated_actions[agent]
        for control in action:
            if control not in ACTUATORS_RATE:
                continue
            actuating_time = acc[control] / ACTUATORS_RATE[control]
            actuating_frac = actuating_time / self._num_steps[agent] / self.dt
            if actuating_frac >= 0.1:
                action[control][:] = 0.0
        self._cumulative_rewards[agent] = 0
        self.actions[agent] = action
        if self._sel.is_last():
            joint = {c: np.zeros(self.num_turbines, dtype=np.float32) for c in self.mdp.controls}
            for j, (_a, act) in enumerate(self.actions.items()):
                for c in act:
                    joint[c][j] = np.asarray(act[c]).reshape(-1)[0]  # reference: joint_action[control][j] = action[control][:]
            next_state, powers, loads, truncated = self.mdp.take_action(self._state, joint)
            normalized = powers * 1e3 / (self._state["freewind_measurements"][0] ** 3)
            load_penalty = np.mean(np.abs(loads)) if loads is not None else 0
            reward = np.array([self.reward_shaper(normalized.mean() - self.load_coef * load_penalty)])
            self._state = next_state
            for a in self.agents:
                i = self.agent_name_mapping[a]
                self.infos[a]["load"] = loads[i]
                self.rewards[a] = reward
                self.truncations[a] = truncated
                self.terminations[a] = False
                self.infos[a]["power"] = powers[i]
            self.num_moves += 1
        else:
            for a in self.rewards:  # _clear_rewards
                self.rewards[a] = 0
        accumulator = self.mdp.get_accumulated_actions()
        for control in action:
            self.accumulated_actions[agent][control] = accumulator[control][self.agent_name_mapping[agent]]
        self.agent_selection = self._sel.next()
        for a, r in self.rewards.items():  # _accumulate_rewards
            self._cumulative_rewards[a] += r


# ----------------------------------------------------------------------------------------------------------------------
# Checker of the library's device-side reset sampler (wf_reset_sampled).  NOT a reference algorithm: the reference draws
# from numpy's Generator (mdp.py:235-258), whose stream cannot be reproduced on a GPU; the library draws the same
# DISTRIBUTION from Philox4x32-10 keyed by (seed; global env id, episode index).  This restatement pins the words and
# the transforms so the tests can check the draws value by value and against the reference's distribution.
# ----------------------------------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011): counter = 4 words, key = 2 words -> 4 words."""
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def _u53(a, b):
    return (((a >> 5) << 26) | (b >> 6)) / 9007199254740992.0


def sampled_reset_wind(seed: int, global_env_id: int, episode: int, ti_range=None):
    """(wind_speed, wind_direction[, ti]) the library draws for this (seed, env, episode)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    g = (global_env_id & 0xFFFFFFFF, (global_env_id >> 32) & 0xFFFFFFFF)
    r0 = philox4x32_10((g[0], g[1], episode, 0), key)
    r1 = philox4x32_10((g[0], g[1], episode, 1), key)
    ws = float(np.clip(8.0 * (-np.log1p(-_u53(r0[0], r0[1]))) ** 0.125, 3.0, 28.0))
    z = np.sqrt(-2.0 * np.log1p(-_u53(r0[2], r0[3]))) * np.cos(2.0 * np.pi * _u53(r1[0], r1[1]))
    wd = float(np.clip((270.0 + 20.0 * z) % 360.0, 0.0, 360.0))
    if ti_range is None:
        return ws, wd
    return ws, wd, ti_range[0] + (ti_range[1] - ti_range[0]) * _u53(r1[2], r1[3])
