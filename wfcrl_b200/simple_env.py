"""Centralised Gymnasium-style environment (single env, drop-in path).

Public surface of the reference's ``WindFarmEnv`` (wfcrl/simple_env.py:13-99): ``reset(seed, options) -> observation``
(the observation only, like the reference), ``step(actions) -> (obs, reward[1], terminated=False, truncated, info)`` with
``info = {"power": MW (T,), "load": (T, 4)}``, attributes ``mdp``, ``action_space``, ``observation_space``,
``num_turbines``, ``max_num_steps``, ``controls``, ``num_moves``, ``accumulated_actions``.
For thousands of envs at once use ``wfcrl_b200.vector_env.VecWindFarmEnv`` (same semantics fused into one kernel).
"""
from __future__ import annotations

import copy
from typing import Dict

import numpy as np

from .environments.data_cases import FarmCase
from .interface import BaseInterface
from .mdp import WindFarmMDP
from .rewards import DoNothingReward, RewardShaper

try:  # pragma: no cover
    import gymnasium as _gym

    _EnvBase = _gym.Env
except Exception:
    class _EnvBase:  # minimal gymnasium.Env stand-in
        metadata = {}

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass


class WindFarmEnv(_EnvBase):
    metadata = {"name": "centralized-windfarm"}

    def __init__(self, interface: BaseInterface, farm_case: FarmCase, controls: dict, continuous_control: bool = True,
                 reward_shaper: RewardShaper = None, start_iter: int = 0, max_num_steps: int = 500,
                 load_coef: float = 0.1):
        self.mdp = WindFarmMDP(interface=interface, farm_case=farm_case, controls=controls,
                               continuous_control=continuous_control, start_iter=start_iter,
                               horizon=start_iter + max_num_steps)
        self.continuous_control = continuous_control
        self.action_space = self.mdp.action_space
        self.observation_space = self.mdp.state_space
        self._state = self.mdp.start_state
        self.num_turbines = self.mdp.num_turbines
        self.max_num_steps = max_num_steps
        self.reward_shaper = reward_shaper if reward_shaper is not None else DoNothingReward()
        self.controls = controls
        self.dt = farm_case.dt
        self.farm_case = farm_case
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        self.num_moves = 0
        self.load_coef = load_coef

    def reset(self, seed=None, options=None):
        self.mdp.reset(seed, options)
        self._state = self.mdp.start_state
        self.reward_shaper.reset()
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        self.num_moves = 0
        return copy.deepcopy(self._state)

    def _apply_actuation_constraint(self, actions: Dict):
        """An actuator that has been moving for >= 10 % of the elapsed time is frozen for this step (in place)."""
        for control in actions:
            rate = self.mdp.ACTUATORS_RATE.get(control)
            if rate is None:
                continue
            busy_time = self.accumulated_actions[control] / rate
            busy_frac = busy_time / self.num_moves / self.farm_case.dt
            actions[control][busy_frac >= 0.1] = 0.0

    def step(self, actions: Dict):
        assert self._state is not None, "Call reset before `step`"
        self.num_moves += 1
        self._apply_actuation_constraint(actions)
        next_state, powers, loads, truncated = self.mdp.take_action(self._state, actions)
        # normalised by the free-stream speed of the state the action was taken in
        reference_speed = self._state["freewind_measurements"][0]
        reward = (powers * 1e3 / (reference_speed ** 3)).mean()
        info = {"power": powers}
        if loads is not None:
            reward = reward - self.load_coef * np.mean(np.abs(loads))
            info["load"] = loads
        reward = np.array([self.reward_shaper(reward)])
        self._state = next_state
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        return copy.deepcopy(self._state), reward, False, truncated, info

    def close(self):
        pass
